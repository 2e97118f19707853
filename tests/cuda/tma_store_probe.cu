// Micro-benchmark: bytes per clock and SM that leave shared memory for global memory through the store paths stft_tc.cu could use,
// every SM running at once.  Four issuing threads per CTA (one per warp), each keeping up to DEPTH store groups in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store_probe tma_store_probe.cu -lcuda ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: 2-D boxes 32 floats x 32 rows (4 KB, 128-byte swizzle), rows 4 KB apart in global memory -- the spectrogram pattern
// mode 1: 3-D boxes 32 x 32 x 2 (8 KB): both 128-byte halves of 32 rows
// mode 2: mode 1 onto a region that stays in L2
// mode 3: 1-D bulk stores of 8 KB contiguous
// mode 4: 2-D boxes 64 floats x 32 rows without swizzle (8 KB, 256-byte rows)
// mode 5: 2-D boxes 256 floats x 8 rows without swizzle (8 KB, 1 KB rows)
// mode 6: st.global.v4 from registers, a warp writes 512 contiguous bytes per instruction (16 warps)
// mode 7 / 8: mode 1 with an L2 evict_first / evict_last cache hint on the stores
__global__ void __launch_bounds__(512, 1) probe(const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap m3,
                                                const __grid_constant__ CUtensorMap m4, const __grid_constant__ CUtensorMap m5, float* out,
                                                unsigned long long rows_total, int mode, int iters, int depth, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 4 * 2 * 8192 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  const unsigned long long rows_per_cta = rows_total / gridDim.x;          // rows of 4 KB (1,024 floats)
  if (mode != 6) {
    if (warp < 4 && lane == 0) {
      unsigned long long pol = 0;
      if (mode == 7) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      if (mode == 8) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
      const uint32_t src = smem_u32(smem) + (uint32_t)(warp * 16384);
      for (int it = 0; it < iters; ++it) {
        // the four issuers of a CTA write the four 32-row quarters of a 128-row tile; 16 chunks of 64 floats across, then the next tile
        const int tile = it / 16, ch = it % 16;
        unsigned long long row = (unsigned long long)blockIdx.x * rows_per_cta + (unsigned long long)((mode == 2 ? 0 : tile) * 128 + warp * 32);
        row %= (rows_total - 32);
        const uint32_t s = src + (uint32_t)((it & 1) * 8192);
        if (mode == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&m2), "r"(s), "r"(ch * 64), "r"((int)row) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&m2), "r"(s + 4096), "r"(ch * 64 + 32), "r"((int)row) : "memory");
        } else if (mode == 1 || mode == 2) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&m3), "r"(s), "r"(0), "r"((int)row), "r"(2 * ch) : "memory");
        } else if (mode >= 7) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(&m3), "r"(s), "r"(0), "r"((int)row), "r"(2 * ch), "l"(pol) : "memory");
        } else if (mode == 3) {
          float* g = out + ((unsigned long long)blockIdx.x * rows_per_cta * 1024ull + (unsigned long long)(it * 4 + warp) * 2048ull) % (rows_total * 1024ull - 2048ull);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s), "r"(8192) : "memory");
        } else if (mode == 4) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&m4), "r"(s), "r"(ch * 64), "r"((int)row) : "memory");
        } else {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&m5), "r"(s), "r"((ch & 3) * 256), "r"((int)(row + (ch >> 2) * 8)) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    // 16 warps x iters/4 stores of 512 B per warp instruction: same bytes per CTA as the other modes (iters x 4 x 8 KB)
    float4 v = make_float4((float)tid, 1.f, 2.f, 3.f);
    float* base = out + ((unsigned long long)blockIdx.x * rows_per_cta * 1024ull);
    const unsigned long long span = rows_per_cta * 1024ull;
    const int n = iters * 4 * 8192 / (16 * 512);
    for (int it = 0; it < n; ++it) {
      const unsigned long long off = ((unsigned long long)(it * 16 + warp) * 128ull + lane * 4ull) % (span - 128ull);
      *reinterpret_cast<float4*>(base + off) = v;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no driver entry point\n"); return 1; }
  encode_fn enc = (encode_fn)p;
  const unsigned long long rows = 148ull * 128ull * 24ull;          // 1.86 GB of 4 KB rows
  float* out; long long* cyc;
  cudaMalloc(&out, rows * 4096ull); cudaMalloc(&cyc, 148 * 8);
  CUtensorMap m2, m3, m4, m5;
  {
    const cuuint64_t d[2] = {1024, rows}, s[1] = {4096}; const cuuint32_t b[2] = {32, 32}, e[2] = {1, 1};
    if (enc(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("m2 failed\n"); return 1; }
  }
  {
    const cuuint64_t d[3] = {32, rows, 32}, s[2] = {4096, 128}; const cuuint32_t b[3] = {32, 32, 2}, e[3] = {1, 1, 1};
    if (enc(&m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("m3 failed\n"); return 1; }
  }
  {
    const cuuint64_t d[2] = {1024, rows}, s[1] = {4096}; const cuuint32_t b[2] = {64, 32}, e[2] = {1, 1};
    if (enc(&m4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("m4 failed\n"); return 1; }
  }
  {
    const cuuint64_t d[2] = {1024, rows}, s[1] = {4096}; const cuuint32_t b[2] = {256, 8}, e[2] = {1, 1};
    if (enc(&m5, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("m5 failed\n"); return 1; }
  }
  const int smem = 4 * 2 * 8192;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"2-D 32x32 swizzle128 (2 x 4 KB)", "3-D 32x32x2 swizzle128 (8 KB)", "3-D, target stays in L2", "1-D bulk 8 KB contiguous",
                         "2-D 64x32 no swizzle (8 KB)", "2-D 256x8 no swizzle (8 KB)", "st.global.v4, 16 warps", "3-D, L2 evict_first hint", "3-D, L2 evict_last hint"};
  const int iters = 16 * 24;                                             // 24 tiles of 128 rows per CTA
  for (int mode = 0; mode < 9; ++mode)
    for (int depth : {1, 2, 4}) {
      if (mode == 6 && depth != 1) continue;
      probe<<<148, 512, smem>>>(m2, m3, m4, m5, out, rows, mode, iters, depth, cyc);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      probe<<<148, 512, smem>>>(m2, m3, m4, m5, out, rows, mode, iters, depth, cyc);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d error %s\n", mode, cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, a, b);
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * 4 * 8192;
      printf("mode %d %-34s depth %d: %6.1f B/clk/SM (CTA 0), %6.2f TB/s whole chip\n", mode, names[mode], depth, bytes / (double)h[0], bytes * 148 / (ms * 1e-3) / 1e12);
    }
  return 0;
}
