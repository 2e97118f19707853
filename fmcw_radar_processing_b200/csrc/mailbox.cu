// Peer-memory ("mailbox") hand-offs of the frame-sharded STFT: instead of two tiny NCCL collectives
// with a host-enqueued stream hop each, every rank stores its shard header and, later, its local maximum straight
// into the mailboxes of all ranks over NVLink (plain st.global to peer-mapped memory) and raises a per-rank,
// per-step flag; the consumer kernels spin on the flags in their own mailbox.  Nothing in the reference
// corresponds to this (it is single-process MATLAB).
//
// Mailbox layout (one per rank, zero-initialised, mapped on every rank):
//   u64 flag_heads[64] | u64 flag_max[64] | f64 max[64] | f64 heads[64][window_length]
// heads[r] = {L_local, first window_length-1 samples of shard r}: the layout stft_plan_kernel reads.
// Flags carry the step number (monotonic), so a flag of an earlier step never satisfies a wait.  The step number is a
// device-side counter of the handle (mailbox_post_heads_kernel advances it, the other kernels of the pass read it), so the
// whole pass -- post, plan, max search, post, collect -- replays as ONE CUDA graph with constant kernel arguments.  One buffer is
// enough: a rank posts the heads of step k+1 only after its STFT of step k, which waited for every rank's maximum
// of step k, which every rank posted after consuming all heads of step k (and likewise for the maxima).
#include "fmcw_internal.cuh"

namespace fmcw {

namespace {

__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(512) mailbox_post_heads_kernel(const sig_t* __restrict__ xc,
                                                                 const unsigned long long* __restrict__ d_ndet, uint32_t PN,
                                                                 uint32_t win, MailboxSet mb, uint32_t world, uint32_t rank,
                                                                 unsigned long long* d_step) {
  const unsigned long long L = *d_ndet * PN;
  const uint32_t i = threadIdx.x;
  const unsigned long long step = *d_step + 1;       // this pass (every thread reads before thread 0 advances the counter)
  if (i < win) {
    const double v = (i == 0) ? (double)L : ((i - 1 < L) ? xc[i - 1] : 0.0);
    for (uint32_t p = 0; p < world; ++p) mailbox_heads(mb.ptr[p])[(size_t)rank * win + i] = v;
    __threadfence_system();
  }
  __syncthreads();
  if (i < world) st_flag(mailbox_flag_heads(mb.ptr[i]) + rank, step);
  if (i == 0) *d_step = step;
}

__global__ void __launch_bounds__(64) mailbox_post_max_kernel(const double* __restrict__ local_max, MailboxSet mb,
                                                              uint32_t world, uint32_t rank, const unsigned long long* d_step) {
  const uint32_t p = threadIdx.x;
  if (p >= world) return;
  const unsigned long long step = *d_step;
  mailbox_max(mb.ptr[p])[rank] = *local_max;
  __threadfence_system();
  st_flag(mailbox_flag_max(mb.ptr[p]) + rank, step);
}

// global maximum = max of the world posted maxima; -1 (error -6 in *d_err) if a rank never posted
__global__ void __launch_bounds__(64) mailbox_collect_max_kernel(void* own, uint32_t world, const unsigned long long* d_step,
                                                                 double* __restrict__ gmax, int* d_err) {
  __shared__ double s_m[64];
  __shared__ int s_bad;
  const uint32_t p = threadIdx.x;
  const unsigned long long step = *d_step;
  if (p == 0) s_bad = 0;
  __syncthreads();
  const bool ok = mailbox_wait(mailbox_flag_max(own), p, world, step);
  if (!ok) s_bad = 1;
  s_m[p] = (p < world && ok) ? const_cast<const volatile double*>(mailbox_max(own))[p] : 0.0;   // peer-written: not the read-only path
  __syncthreads();
  if (p == 0) {
    double m = 0.0;
    for (uint32_t r = 0; r < world; ++r) m = s_m[r] > m ? s_m[r] : m;
    if (s_bad) { *d_err = -6; m = 1.0; }
    *gmax = m;
  }
}

cudaError_t launch_mailbox_post_heads(const sig_t* xc, const unsigned long long* d_ndet, uint32_t PN, uint32_t win,
                                      const MailboxSet& mb, uint32_t world, uint32_t rank, unsigned long long* d_step,
                                      cudaStream_t st) {
  mailbox_post_heads_kernel<<<1, 512, 0, st>>>(xc, d_ndet, PN, win, mb, world, rank, d_step);
  return cudaGetLastError();
}
cudaError_t launch_mailbox_post_max(const double* local_max, const MailboxSet& mb, uint32_t world, uint32_t rank,
                                    const unsigned long long* d_step, cudaStream_t st) {
  mailbox_post_max_kernel<<<1, 64, 0, st>>>(local_max, mb, world, rank, d_step);
  return cudaGetLastError();
}
cudaError_t launch_mailbox_collect_max(void* own, uint32_t world, const unsigned long long* d_step, double* gmax, int* d_err,
                                       cudaStream_t st) {
  mailbox_collect_max_kernel<<<1, 64, 0, st>>>(own, world, d_step, gmax, d_err);
  return cudaGetLastError();
}

}  // namespace fmcw
