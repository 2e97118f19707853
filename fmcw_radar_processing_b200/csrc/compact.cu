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

// Recordings of up to COMPACT_FUSED_MAX frames: one launch.  A CTA owns 32 frames; it counts the detections before
// its first frame itself (n_frames * 4 B of flags are L2 resident), ranks its own frames with a ballot and copies
// their rows; the CTA of the last frames publishes the total.
constexpr uint64_t COMPACT_FUSED_MAX = 16384;
constexpr int COMPACT_FPB = 32;

__global__ void __launch_bounds__(256) compact_fused_kernel(const int32_t* __restrict__ det, uint64_t n_frames, uint32_t PN,
                                                            const sig_t* __restrict__ slow_mag, sig_t* __restrict__ xc,
                                                            unsigned long long* __restrict__ n_det) {
  __shared__ unsigned s_cnt[8];
  __shared__ unsigned s_before, s_ballot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t f0 = (uint64_t)blockIdx.x * COMPACT_FPB;
  unsigned cnt = 0;
  for (uint64_t i = tid; i < f0; i += 256) cnt += det[i] != 0 ? 1u : 0u;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
  if (lane == 0) s_cnt[warp] = cnt;
  if (warp == 0) {
    const uint64_t f = f0 + lane;
    const unsigned b = __ballot_sync(0xffffffffu, f < n_frames && det[f] != 0);
    if (lane == 0) s_ballot = b;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; ++w) t += s_cnt[w];
    s_before = t;
    if (f0 + COMPACT_FPB >= n_frames) *n_det = (unsigned long long)t + __popc(s_ballot);
  }
  __syncthreads();
  const unsigned before = s_before, ballot = s_ballot;
  for (uint32_t i = tid; i < COMPACT_FPB * PN; i += 256) {
    const uint32_t j = i / PN, c = i - j * PN;
    if (ballot & (1u << j)) {
      const unsigned k = before + __popc(ballot & ((1u << j) - 1u));
      xc[(uint64_t)k * PN + c] = slow_mag[(f0 + j) * PN + c];
    }
  }
}

// Long recordings: two launches, both over the whole grid.  count: one thread per frame, per-warp (32 frames) and per-CTA
// (1,024 frames) detection counts.  gather: a CTA owns 32 frames as above; its offset = the CTA counts before its 1,024-frame
// block + the warp counts before it inside the block (at most 1,954 + 31 small loads for 2 M frames).
__global__ void __launch_bounds__(1024) compact_count_kernel(const int32_t* __restrict__ det, uint64_t n_frames,
                                                             uint32_t* __restrict__ cnt32, uint32_t* __restrict__ cnt1024) {
  __shared__ uint32_t s_w[32];
  const uint64_t f = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = __popc(__ballot_sync(0xffffffffu, f < n_frames && det[f] != 0));
  if (lane == 0) { s_w[warp] = b; cnt32[(uint64_t)blockIdx.x * 32 + warp] = b; }
  __syncthreads();
  if (warp == 0) {
    uint32_t v = s_w[lane];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) cnt1024[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) compact_gather2_kernel(const int32_t* __restrict__ det, uint64_t n_frames, uint32_t PN,
                                                              const sig_t* __restrict__ slow_mag, sig_t* __restrict__ xc,
                                                              const uint32_t* __restrict__ cnt32, const uint32_t* __restrict__ cnt1024,
                                                              unsigned long long* __restrict__ n_det) {
  __shared__ unsigned long long s_cnt[8];
  __shared__ unsigned long long s_before;
  __shared__ unsigned s_ballot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t g = blockIdx.x;                       // 32-frame group
  const uint64_t f0 = g * COMPACT_FPB, blk = g >> 5, win = g & 31;
  unsigned long long cnt = 0;
  for (uint64_t i = tid; i < blk; i += 256) cnt += cnt1024[i];
  if ((uint64_t)tid < win) cnt += cnt32[blk * 32 + tid];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
  if (lane == 0) s_cnt[warp] = cnt;
  if (warp == 0) {
    const uint64_t f = f0 + lane;
    const unsigned b = __ballot_sync(0xffffffffu, f < n_frames && det[f] != 0);
    if (lane == 0) s_ballot = b;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += s_cnt[w];
    s_before = t;
    if (f0 + COMPACT_FPB >= n_frames) *n_det = t + __popc(s_ballot);
  }
  __syncthreads();
  const unsigned long long before = s_before;
  const unsigned ballot = s_ballot;
  for (uint32_t i = tid; i < COMPACT_FPB * PN; i += 256) {
    const uint32_t j = i / PN, c = i - j * PN;
    if (ballot & (1u << j)) {
      const unsigned long long k = before + __popc(ballot & ((1u << j) - 1u));
      xc[k * PN + c] = slow_mag[(f0 + j) * PN + c];
    }
  }
}

cudaError_t launch_compact(const CompactParams& p, cudaStream_t st) {
  if (p.n_frames > 0 && p.n_frames <= COMPACT_FUSED_MAX) {
    const unsigned blocks = (unsigned)((p.n_frames + COMPACT_FPB - 1) / COMPACT_FPB);
    compact_fused_kernel<<<blocks, 256, 0, st>>>(p.detected, p.n_frames, p.PN, p.slow_mag, p.xc, p.n_det);
    return cudaGetLastError();
  }

  if (p.n_frames > 0 && p.counts != nullptr) {
    const unsigned b1024 = (unsigned)((p.n_frames + 1023) / 1024);
    uint32_t* cnt32 = p.counts;
    uint32_t* cnt1024 = p.counts + (size_t)b1024 * 32;
    compact_count_kernel<<<b1024, 1024, 0, st>>>(p.detected, p.n_frames, cnt32, cnt1024);
    const unsigned groups = (unsigned)((p.n_frames + COMPACT_FPB - 1) / COMPACT_FPB);
    compact_gather2_kernel<<<groups, 256, 0, st>>>(p.detected, p.n_frames, p.PN, p.slow_mag, p.xc, cnt32, cnt1024, p.n_det);
    return cudaGetLastError();
  }
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
