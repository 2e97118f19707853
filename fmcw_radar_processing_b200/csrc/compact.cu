// Compaction of the detected frames' slow-time magnitude rows into one contiguous signal
// (the growing concatenation at RP:257-260): exclusive scan of the detection flags, then a gather.
#include "fmcw_internal.cuh"

namespace fmcw {

// One CTA walks the flags in tiles of 1024 (n_frames * 4 B is tiny next to the frame data).
__global__ void __launch_bounds__(1024) scan_flags_kernel(const int32_t* __restrict__ det, uint64_t n_frames,
                                                          uint32_t* __restrict__ det_list,
                                                          unsigned long long* __restrict__ n_det) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < n_frames; base += 1024) {
    const uint64_t f = base + tid;
    const uint32_t flag = (f < n_frames && det[f] != 0) ? 1u : 0u;
    const uint32_t ballot = __ballot_sync(0xffffffffu, flag);
    const uint32_t below = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
      const uint32_t v = s_warp[lane];
      uint32_t x = v;
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, x, m);
        if (lane >= m) x += o;
      }
      s_warp[lane] = x - v;   // exclusive prefix of the warp totals
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t pos = carry + s_warp[warp] + below;
    if (flag) det_list[pos] = (uint32_t)f;
    __syncthreads();
    if (tid == 1023) s_carry = pos + flag;   // inclusive total of this tile
    __syncthreads();
  }
  if (tid == 0) *n_det = s_carry;
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const sig_t* __restrict__ slow_mag,
                                                          const uint32_t* __restrict__ det_list,
                                                          const unsigned long long* __restrict__ n_det, uint32_t PN,
                                                          uint64_t n_frames, sig_t* __restrict__ xc) {
  const unsigned long long L = *n_det * PN;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < L;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long k = i / PN;
    const uint32_t c = (uint32_t)(i - k * PN);
    xc[i] = slow_mag[(uint64_t)det_list[k] * PN + c];
  }
}

cudaError_t launch_compact(const CompactParams& p, cudaStream_t st) {
  scan_flags_kernel<<<1, 1024, 0, st>>>(p.detected, p.n_frames, p.det_list, p.n_det);
  const uint64_t total = p.n_frames * p.PN;
  uint64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks == 0) blocks = 1;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(p.slow_mag, p.det_list, p.n_det, p.PN, p.n_frames, p.xc);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) f32_to_sig_kernel(const float* __restrict__ src, sig_t* __restrict__ dst, unsigned long long n) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    dst[i] = (sig_t)src[i];
}

cudaError_t launch_f32_to_sig(const float* src, sig_t* dst, unsigned long long n, cudaStream_t st) {
  if (!n) return cudaSuccess;
  unsigned long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32_to_sig_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}

}  // namespace fmcw
