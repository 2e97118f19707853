// Micro-benchmark: tcgen05.ld (TMEM read) and MUFU.LG2 throughput per SM sub-partition, alone and together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
                 "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
                 "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]),
                 "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),
                 "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ float lg2a(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// mode 0: LDTM only (4 x16 per iteration); 1: MUFU only (32 per iteration); 2: same warp, loads of the next
// iteration in flight under the MUFUs of this one; 3: even warps of a sub-partition load, odd warps MUFU;
// 4: LDTM only with x32 (2 per iteration); 5: same-warp serial (ld, wait, mufu)
__global__ void probe(int mode, int iters, long long* cyc, float* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  float a[16], b[16], c[16], d[16];
  for (int i = 0; i < 16; ++i) { a[i] = 1.5f + i + threadIdx.x; b[i] = 2.5f + i; c[i] = 3.5f + i; d[i] = 4.5f + i; }
  __syncthreads();
  const long long t0 = clock64();
  const bool loader = (mode == 0) || (mode == 4) || (mode == 2) || (mode == 5) || (mode == 3 && ((warp >> 2) & 1) == 0);
  const bool mufu = (mode == 1) || (mode == 2) || (mode == 5) || (mode == 3 && ((warp >> 2) & 1) == 1);
  if (mode == 4) {
    float v[32];
    for (int it = 0; it < iters; ++it) {
      tmem_ld32(base + (uint32_t)((it & 1) * 64), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += v[0] + v[31];
      tmem_ld32(base + (uint32_t)((it & 1) * 64 + 32), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += v[0] + v[31];
    }
  } else if (mode == 2) {
    float e[16], f[16], g[16], h[16];
    tmem_ld16(base, a); tmem_ld16(base + 16, b); tmem_ld16(base + 32, c); tmem_ld16(base + 48, d);
    for (int it = 0; it < iters; it += 2) {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld16(base + 64, e); tmem_ld16(base + 80, f); tmem_ld16(base + 96, g); tmem_ld16(base + 112, h);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += lg2a(a[i]) + lg2a(b[i]) + lg2a(c[i]) + lg2a(d[i]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld16(base, a); tmem_ld16(base + 16, b); tmem_ld16(base + 32, c); tmem_ld16(base + 48, d);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += lg2a(e[i]) + lg2a(f[i]) + lg2a(g[i]) + lg2a(h[i]);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  } else {
    for (int it = 0; it < iters; ++it) {
      if (loader) {
        const uint32_t o = (uint32_t)((it & 3) * 64);
        tmem_ld16(base + o, a); tmem_ld16(base + o + 16, b); tmem_ld16(base + o + 32, c); tmem_ld16(base + o + 48, d);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!mufu) acc += a[0] + b[5] + c[7] + d[15];
      }
      if (mufu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += lg2a(a[i]) + lg2a(b[i]) + lg2a(c[i]) + lg2a(d[i]);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(512));
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  const char* names[] = {"ldtm x16 only", "mufu only", "same warp, ld under mufu", "even warps ld / odd warps mufu", "ldtm x32 only", "same warp serial"};
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 6; ++mode) {
      if (mode == 3 && warps < 8) continue;
      probe<<<148, warps * 32, 0>>>(mode, iters, cyc, sink);
      probe<<<148, warps * 32, 0>>>(mode, iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("warps/CTA %2d  mode %d (%-32s): %8.1f cycles per iteration of thread 0 (4 KB... per warp: 4 x16 loads = 8 KB and/or 32 MUFU)\n",
             warps, mode, names[mode], (double)h[0] / iters);
    }
  }
  return 0;
}
