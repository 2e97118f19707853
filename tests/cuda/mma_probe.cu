// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (M = 128, K = 8 per instruction) as stft_tc.cu uses it, per operand layout:
// shared-memory operands K-major without swizzle (the kernel's layout) or with the 128-byte swizzle, N = 128 or 256, A from shared
// memory or from TMEM -- alone and with the other warps of the CTA streaming 128-bit shared-memory stores (the epilogue's staging).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// mode bits: 1 = 128-byte swizzle, 2 = N 256, 4 = A from TMEM, 8 = other warps store to shared memory meanwhile
__global__ void __launch_bounds__(544, 1) probe(int mode, int rounds, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);                 // 128 x 32 floats
  float* sB = reinterpret_cast<float*>(smem + 16384);         // 256 x 32 floats
  float* sS = reinterpret_cast<float*>(smem + 16384 + 32768); // 64 KB scratch for the store traffic
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f + (i & 7);
  if (tid == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tmem_base;
  const bool swz = mode & 1, n256 = mode & 2, a_tmem = mode & 4, traffic = mode & 8;
  if (warp == 16) {
    if (lane == 0) {
      const uint32_t N = n256 ? 256u : 128u;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t layout = swz ? 2u : 0u;
      const uint32_t kstep = swz ? 32u : 256u;          // bytes to the next K = 8 slice
      const uint32_t lbo = swz ? 16u : 128u, sbo = 1024u;
      const long long t0 = clock64();
      for (int r = 0; r < rounds; ++r) {
        const uint32_t d = tb + (uint32_t)((r & 1) * (n256 ? 256 : 128));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t db = make_desc(smem_u32(sB) + ks * kstep, lbo, sbo, layout);
          const uint32_t acc = ks > 0;
          if (a_tmem) {
            const uint32_t ta = tb + 480u + (uint32_t)(ks * 8);     // A operand: 32 columns at the top of TMEM
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                         ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
          } else {
            const uint64_t da = make_desc(smem_u32(sA) + ks * kstep, lbo, sbo, layout);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                         ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      }
      const long long t1 = clock64();
      cyc[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (traffic) {
    // 16 warps: 128-bit stores, conflict-free, until the MMA thread is done
    const uint32_t a = smem_u32(sS) + (uint32_t)(warp * 4096 + lane * 16);
    long long n = 0;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(a + (uint32_t)(i * 512)), "f"((float)i) : "memory");
      ++n;
    }
    if (lane == 0 && warp == 0) cyc[148 + blockIdx.x] = n * 8 * 16;   // store instructions of the CTA (all warps run alike)
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 2 * 148 * 8);
  const int smem = 16384 + 32768 + 65536;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int rounds = 4000;
  for (int mode = 0; mode < 16; ++mode) {
    cudaMemset(cyc, 0, 2 * 148 * 8);
    probe<<<148, 544, smem>>>(mode, rounds, cyc);
    probe<<<148, 544, smem>>>(mode, rounds, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d error %s\n", mode, cudaGetErrorString(e)); return 1; }
    long long h[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / (rounds * 4.0);
    printf("mode %2d  %-10s N=%3d  A from %-4s %-14s: %7.1f cycles per K=8 instruction", mode, (mode & 1) ? "swizzle128" : "no swizzle",
           (mode & 2) ? 256 : 128, (mode & 4) ? "TMEM" : "smem", (mode & 8) ? "+ STS traffic" : "", per);
    if (mode & 8) printf("   (%.1f B/clk of stores alongside)", (double)h[148] * 512.0 / (double)h[0]);
    printf("\n");
  }
  return 0;
}
