// Standalone check of the hand-written tcgen05 TF32 path used by stft_tc.cu: shared-memory matrix
// descriptors (K-major, no swizzle), instruction descriptor, TMEM allocation, commit -> mbarrier,
// tcgen05.ld 32x32b.  D[128x128] = A[128x16] * B[128x16]^T, compared with the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tc_gemm_check tc_gemm_check.cu && ./tc_gemm_check
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__global__ void __launch_bounds__(128) gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  __shared__ __align__(128) float sA[128 * 16];
  __shared__ __align__(128) float sB[128 * 16];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  // canonical K-major layout: (r/8)*512 + (k/4)*128 + (r%8)*16 + (k%4)*4 bytes
  for (int i = tid; i < 128 * 16; i += 128) {
    const int r = i / 16, k = i % 16;
    const int off = (r / 8) * 128 + (k / 4) * 32 + (r % 8) * 4 + (k % 4);   // in floats
    sA[off] = A[i];
    sB[off] = B[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");      // generic-proxy smem writes -> async proxy (tensor core)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t da = make_desc(smem_u32(sA) + ks * 256, 128, 512);
      const uint64_t db = make_desc(smem_u32(sB) + ks * 256, 128, 512);
      const uint32_t acc = ks > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  // wait for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 16; ++j) D[tid * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tb));
}

static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> A(128 * 16), B(128 * 16), D(128 * 128), R(128 * 128);
  srand(1);
  for (auto& v : A) v = tf32((rand() / (float)RAND_MAX) * 2 - 1);
  for (auto& v : B) v = tf32((rand() / (float)RAND_MAX) * 2 - 1);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) { double s = 0; for (int k = 0; k < 16; ++k) s += (double)A[m * 16 + k] * B[n * 16 + k]; R[m * 128 + n] = (float)s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  gemm_kernel<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (size_t i = 0; i < D.size(); ++i) maxerr = fmax(maxerr, fabs((double)D[i] - R[i]));
  printf("max abs err %.3e  D[0]=%f R[0]=%f D[129]=%f R[129]=%f D[16383]=%f R[16383]=%f\n", maxerr, D[0], R[0], D[129], R[129], D[16383], R[16383]);
  return maxerr < 1e-5 ? 0 : 1;
}
