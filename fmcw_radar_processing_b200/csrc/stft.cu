// Cross-frame STFT micro-Doppler spectrogram (radar_processing.m RP:270-299), restated so that it
// never materialises the (nfft/2+1) x ncol PSD matrix of the literal code:
//   * nfft = 2^nextpow2(L) (RP:273) only fixes the fine frequency grid F_j = j*fs/nfft;
//   * interp1 onto the 1024 log-spaced frequencies (RP:293-299) touches only the fine-grid bins that
//     bracket a query; the windowed DTFT is evaluated at exactly those bins (~1.1k-1.7k);
//   * the global normalisation max(P) (RP:282-283) is found exactly from per-column bounds.
// Everything that depends on L (nfft, bins, coefficients, chunks) is planned ON THE DEVICE, so the
// fused path frames -> compaction -> STFT runs without a host round trip.
#include <cstdlib>

#include "fmcw_internal.cuh"

namespace fmcw {

constexpr float K_DB = 6.020599913279624f;   // 20*log10(2): psd = 20*log10(P/max) (RP:283)
constexpr int NP_MAX = 320;                  // bin positions a chunk may hold in shared memory

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) stft_plan_kernel(StftTables t, StftGeom g, const unsigned long long* d_ndet,
                                                         uint32_t PN, unsigned long long L_total_host,
                                                         unsigned long long sample_offset,
                                                         unsigned long long L_local_host, unsigned long long L_avail_host,
                                                         int n_chunks_req, int coef_rows, const double* gathered_,
                                                         uint32_t world, uint32_t rank, sig_t* __restrict__ xc, int spec_mode,
                                                         const unsigned long long* wait_flags, const unsigned long long* wait_step_ptr) {
  __shared__ int s_scan[1024];
  __shared__ int s_wsum[32];
  __shared__ unsigned long long s_nfft;
  __shared__ int s_valid;
  __shared__ int s_hit;
  __shared__ unsigned long long s_lay[4];   // L_total, L_local, L_avail, sample_offset of this pass
  // the shard headers: an all-gathered buffer, or this rank's mailbox that PEER GPUs write while this kernel runs --
  // volatile loads after the flag's acquire, never the read-only (ld.global.nc) path
  const volatile double* gathered = gathered_;
  StftPlan* P = t.plan;
  if (wait_flags) {
    // mailbox path: the shard headers arrive by peer stores; wait for the flag of every rank (mailbox.cu)
    __shared__ int s_late;
    if (threadIdx.x == 0) s_late = 0;
    __syncthreads();
    if (!mailbox_wait(wait_flags, threadIdx.x, world, *wait_step_ptr)) s_late = 1;
    __syncthreads();
    if (s_late) { if (threadIdx.x == 0) { P->valid = -6; P->nb = 0; P->n_chunks = 0; } return; }
  }
  // spec_mode 1: planning ahead (side stream, concurrently with the frame chain) for an assumed layout;
  // spec_mode 2: the real layout is known: if it equals the assumed one the tables stand, otherwise plan again
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nq = (int)g.nq;
  const int win = (int)g.win, hop = (int)g.hop, nov = win - hop;

  if (tid == 0) {
    s_hit = 0;
    unsigned long long L = L_total_host, Lloc = L_local_host, Lav = L_avail_host;
    if (gathered) {
      // sharded run: global length, this shard's offset and the halo (the win-1 samples that follow this shard,
      // possibly spanning several short or empty shards) from the all-gathered shard headers
      const uint32_t stride = 1 + (win - 1), hw = win - 1;
      unsigned long long total = 0, soff = 0, mine = 0;
      for (uint32_t r = 0; r < world; ++r) {
        const unsigned long long Lr = (unsigned long long)gathered[r * stride];
        if (r < rank) soff += Lr;
        if (r == rank) mine = Lr;
        total += Lr;
      }
      uint32_t got = 0;
      for (uint32_t r = rank + 1; r < world && got < hw; ++r) {
        const unsigned long long Lr = (unsigned long long)gathered[r * stride];
        const uint32_t take = (uint32_t)(Lr < (unsigned long long)(hw - got) ? Lr : (hw - got));
        for (uint32_t i = 0; i < take; ++i) xc[mine + got + i] = gathered[r * stride + 1 + i];
        got += take;
      }
      L = total; Lloc = mine; Lav = mine + got; sample_offset = soff;
    } else if (d_ndet && spec_mode != 1) { L = *d_ndet * PN; Lloc = L; Lav = L; }
    if (spec_mode == 2 && P->spec_state == 1 && P->valid > 0 && P->L_total == L && P->sample_offset == sample_offset &&
        P->L_avail == Lav && P->L_local == Lloc) {
      P->spec_state = 2;
      s_hit = 1;
    }
    s_lay[0] = L; s_lay[1] = Lloc; s_lay[2] = Lav; s_lay[3] = sample_offset;
  }
  __syncthreads();
  if (s_hit) return;
  if (tid == 0) {
    const unsigned long long L = s_lay[0], Lloc = s_lay[1], Lav = s_lay[2];
    sample_offset = s_lay[3];
    P->L_total = L; P->sample_offset = sample_offset; P->L_avail = Lav; P->L_local = Lloc;
    P->n_hard = 0; P->n_refined = 0; P->lb_max = 0.f; P->pmax_raw = 0.0; P->task_counter = 0; P->ticket_r = 0; P->ticket_h = 0;
    P->spec_state = (spec_mode == 1) ? 1 : 0;
    int ok = (L >= (unsigned long long)win) ? 1 : 0;
    int lg = (L <= 1) ? 0 : 64 - __clzll((long long)(L - 1));
    unsigned long long nfft = 1ull << lg;
    P->log2nfft = lg; P->nfft = nfft;
    unsigned long long ncol = ok ? (L - nov) / hop : 0;
    P->ncol_total = ncol;
    unsigned long long cb = (sample_offset + hop - 1) / hop;
    unsigned long long ce = (sample_offset + Lloc + hop - 1) / hop;
    if (ce > ncol) ce = ncol;
    // every owned column needs win samples inside [offset, offset + L_avail)
    if (sample_offset + Lav >= (unsigned long long)win) {
      unsigned long long last = (sample_offset + Lav - win) / hop + 1;
      if (ce > last) ce = last;
    } else ce = cb;
    if (cb > ce) cb = ce;
    P->col_begin = cb; P->col_end = ce;
    P->nq = nq;
    s_nfft = nfft; s_valid = ok;
    P->valid = ok;
  }
  __syncthreads();
  if (!s_valid) { if (tid == 0) { P->nb = 0; P->n_chunks = 0; } return; }
  const unsigned long long nfft = s_nfft;
  const double df = g.fs / (double)nfft;
  const long long jmax = (long long)(nfft / 2) - 1;

  // query frequency q -> bracket bin j, weight a (matlab logspace + interp1 'linear','extrap')
  long long j = 0; double a = 0.0;
  if (tid < nq) {
    const double d1 = log10(df), d2 = log10((double)(nfft / 2) * df);
    double y = d1 + ((double)tid * (d2 - d1)) / (double)(nq - 1);
    if (tid == 0) y = d1;
    if (tid == nq - 1) y = d2;
    const double fq = pow(10.0, y);
    j = (long long)floor(fq / df);
    if (j < 0) j = 0;
    if (j > jmax) j = jmax;
    a = (fq - (double)j * df) / df;
  }
  // new bins contributed by query q: 2 / 1 / 0
  long long jprev = __shfl_up_sync(0xffffffffu, j, 1);
  __shared__ long long s_lastj[32];
  if (lane == 31) s_lastj[warp] = j;
  __syncthreads();
  if (lane == 0 && warp > 0) jprev = s_lastj[warp - 1];
  int c = 0;
  if (tid < nq) c = (tid == 0) ? 2 : (j == jprev ? 0 : (j == jprev + 1 ? 1 : 2));
  // inclusive block scan
  int x = c;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) { int o = __shfl_up_sync(0xffffffffu, x, m); if (lane >= m) x += o; }
  if (lane == 31) s_wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int v = s_wsum[lane], y2 = v;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) { int o = __shfl_up_sync(0xffffffffu, y2, m); if (lane >= m) y2 += o; }
    s_wsum[lane] = y2 - v;
  }
  __syncthreads();
  const int cum = x + s_wsum[warp];
  const int pq = cum - 2;
  s_scan[tid] = (tid < nq) ? pq : 0x7fffffff;
  __shared__ int s_nb;
  if (tid == nq - 1) s_nb = cum;
  __syncthreads();
  const int nb = s_nb;
  if (nb > t.nb_max) { if (tid == 0) { P->valid = -2; P->nb = nb; } return; }
  if (tid < nq) {
    t.bins[pq] = (int)j;
    t.bins[pq + 1] = (int)j + 1;
    t.qpos[tid] = pq;
    t.aq[tid] = (float)a;
    // qend[p] = first query whose lower position is >= p
    if (tid == 0) { for (int p = 0; p <= pq; ++p) t.qend[p] = 0; }
    else { const int pp = s_scan[tid - 1]; for (int p = pp + 1; p <= pq; ++p) t.qend[p] = tid; }
    if (tid == nq - 1) for (int p = pq + 1; p <= nb; ++p) t.qend[p] = nq;
  }
  __syncthreads();
  __threadfence_block();
  // per-position constants and DTFT coefficients (float64 sincospi of an exactly reduced argument)
  const int half = win / 2;
  const int odd = win & 1;
  const long long mod = (long long)(2 * nfft);
  // only the rows of bins 0/1 (stft_colstat_kernel) are tabulated here; the full table of the CUDA-core kernels
  // comes from stft_coef_kernel, the tensor-core operands from stft_tc_prepare_kernel
  const int n_rows = nb < 2 ? nb : 2;
  for (int i = tid; i < n_rows * half; i += blockDim.x) {
    const int p = i / half, m = i - p * half;
    const long long bin = t.bins[p];
    const long long twod = odd ? (2 * m + 2) : (2 * m + 1);       // 2*delta, delta = tap distance from the centre
    const long long r = (twod * bin) % mod;
    double sn, cs;
    sincospi((double)r / (double)nfft, &sn, &cs);
    t.coef[(size_t)p * 2 * half + m] = (float)cs;
    t.coef[(size_t)p * 2 * half + half + m] = (float)sn;
  }
  for (int p = tid; p < nb; p += blockDim.x) {
    const long long bin = t.bins[p];
    t.kcb[p] = (bin == 0 || bin == (long long)(nfft / 2)) ? 0.f : K_DB;
  }
  // chunks of queries (multiples of 32) of about equal cost: ~60 issue slots per bin, ~14 per query
  if (tid == 0) {
    constexpr long long CB = 60, CQ = 14;
    int nch = n_chunks_req < 1 ? 1 : (n_chunks_req > MAX_CHUNKS ? MAX_CHUNKS : n_chunks_req);
    const long long total = (long long)nb * CB + (long long)nq * CQ;
    int cnt = 0;
    P->chunk_q0[0] = 0; P->chunk_p0[0] = s_scan[0];
    int last_q = 0;
    for (int k = 1; k < nch; ++k) {
      const long long target = (k * total) / nch;
      int q = last_q + 32;
      while (q < nq && (long long)s_scan[q] * CB + (long long)q * CQ < target) q += 32;
      if (q >= nq) break;
      ++cnt;
      P->chunk_q0[cnt] = q; P->chunk_p0[cnt] = s_scan[q];
      last_q = q;
    }
    ++cnt;
    P->chunk_q0[cnt] = nq; P->chunk_p0[cnt] = nb;
    P->n_chunks = cnt;
    P->nb = nb;
    int bad = 0;
    for (int k = 0; k < cnt; ++k) {
      const int p1 = s_scan[P->chunk_q0[k + 1] - 1] + 1;
      if (p1 - P->chunk_p0[k] + 1 > NP_MAX) bad = 1;
    }
    if (bad) P->valid = -3;
  }
}

// full coefficient table [nb][2*half] and the window DC response per bin for the CUDA-core kernels
__global__ void __launch_bounds__(256) stft_coef_kernel(StftTables t, StftGeom g, int spec_mode) {
  const StftPlan* P = t.plan;
  if (P->valid <= 0 || (spec_mode == 2 && P->spec_state == 2)) return;
  const int nb = P->nb, win = (int)g.win, half = win / 2, odd = win & 1;
  const unsigned long long nfft = P->nfft;
  const long long mod = (long long)(2 * nfft);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb * half; i += gridDim.x * blockDim.x) {
    const int p = i / half, m = i - p * half;
    const long long bin = t.bins[p];
    const long long twod = odd ? (2 * m + 2) : (2 * m + 1);
    const long long r = (twod * bin) % mod;
    double sn, cs;
    sincospi((double)r / (double)nfft, &sn, &cs);
    t.coef[(size_t)p * 2 * half + m] = (float)cs;
    t.coef[(size_t)p * 2 * half + half + m] = (float)sn;
  }
  // window DC response sum_n w[n] cos((n - c) w_p) (the window is symmetric, the sine part vanishes): the multiplier
  // of the column mean in the generic kernel's DC split; one warp per bin position
  const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int p = gw; p < nb; p += nw) {
    const long long bin = t.bins[p];
    double acc = (odd && lane == 0) ? t.win_d[half] : 0.0;
    for (int m = lane; m < half; m += 32) {
      const long long twod = odd ? (2 * m + 2) : (2 * m + 1);
      const long long r = (twod * bin) % mod;
      const int lo = half - 1 - m, hi = odd ? (half + 1 + m) : (half + m);
      acc += (t.win_d[lo] + t.win_d[hi]) * cospi((double)r / (double)nfft);
    }
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
    if (lane == 0) t.wdc[p] = (float)acc;
  }
}

// ------------------------------------------------------------------------------------------------
// global maximum of c_j |S_t(w_j)|^2 over the whole fine grid
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float atomic_max_nonneg(float* addr, float v) {
  return __uint_as_float(atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v)));
}

// per column: S0 = sum y, S1 = |S(w_1)|^2; lower bound max(S0^2, c_1 |S1|^2)
__global__ void __launch_bounds__(256) stft_colstat_kernel(StftTables t, StftGeom g, const sig_t* __restrict__ x) {
  const StftPlan* P = t.plan;
  if (P->valid <= 0) return;
  __shared__ float s_w[1024];
  __shared__ float s_c1[1024];
  __shared__ float s_red[8];
  const int win = (int)g.win, half = win / 2, odd = win & 1;
  const int p1 = (t.bins[0] == 1) ? 0 : 1;
  const float c1 = (t.kcb[p1] > 0.f) ? 2.f : 1.f;
  for (int i = threadIdx.x; i < win; i += blockDim.x) s_w[i] = t.win[i];
  for (int i = threadIdx.x; i < 2 * half; i += blockDim.x) s_c1[i] = t.coef[(size_t)p1 * 2 * half + i];
  __syncthreads();
  const unsigned long long cb = P->col_begin, ce = P->col_end, off = P->sample_offset;
  float best = 0.f;
  for (unsigned long long col = cb + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; col < ce;
       col += (unsigned long long)gridDim.x * blockDim.x) {
    const sig_t* xs = x + (col * g.hop - off);
    float s0 = 0.f, re = 0.f, im = 0.f, sabs = 0.f;
    const int cidx = half;   // centre tap for odd windows
    for (int m = 0; m < half; ++m) {
      const int lo = odd ? (cidx - 1 - m) : (half - 1 - m), hi = odd ? (cidx + 1 + m) : (half + m);
      const float ylo = s_w[lo] * (float)xs[lo], yhi = s_w[hi] * (float)xs[hi];
      s0 += ylo + yhi;
      sabs += fabsf(ylo) + fabsf(yhi);
      re = fmaf(ylo + yhi, s_c1[m], re);
      im = fmaf(ylo - yhi, s_c1[half + m], im);
    }
    if (odd) { const float yc = s_w[cidx] * (float)xs[cidx]; s0 += yc; re += yc; sabs += fabsf(yc); }
    const float lb = fmaxf(s0 * s0, c1 * fmaf(re, re, im * im));
    best = fmaxf(best, lb);
    t.col_ub[col - cb] = 2.f * sabs * sabs;          // trivial upper bound of this column's maximum
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, m));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) best = fmaxf(best, s_red[w]);
    if (best > 0.f) atomic_max_nonneg(&t.plan->lb_max, best);
  }
}

// last step of the max search: the lower bound from stft_colstat_kernel and the exhaustive results
__device__ __forceinline__ void finalize_max(StftPlan* P, double* export_dst) {
  const double hard = __longlong_as_double((long long)atomicMax(reinterpret_cast<unsigned long long*>(&P->pmax_raw), 0ull));
  const double lb = (double)P->lb_max;
  const double v = lb > hard ? lb : hard;
  P->pmax_raw = v;
  P->task_counter = 0;
  if (export_dst) *export_dst = v;
}

// Columns whose trivial bound 2*(sum|y|)^2 exceeds the lower bound get a certificate: for y >= 0 the
// spectrum is non-increasing on [0, pi/(win-1)], and on [pi/(win-1), pi] a uniform grid plus the
// Lipschitz constant sum|n-c||y_n| bounds it.  Columns that fail go to the exhaustive list.
__global__ void __launch_bounds__(256) stft_refine_kernel(StftTables t, StftGeom g, const sig_t* __restrict__ x, double* export_dst) {
  StftPlan* P = t.plan;
  if (P->valid <= 0) { if (blockIdx.x == 0 && threadIdx.x == 0 && export_dst) *export_dst = 0.0; return; }
  __shared__ float s_w[1024];
  __shared__ float s_wsum;
  const int win = (int)g.win;
  for (int i = threadIdx.x; i < win; i += blockDim.x) s_w[i] = t.win[i];
  __syncthreads();
  if (threadIdx.x == 0) { float a = 0.f; for (int i = 0; i < win; ++i) a += s_w[i]; s_wsum = a; }
  __syncthreads();
  const unsigned long long cb = P->col_begin, ce = P->col_end, off = P->sample_offset;
  const float lb = P->lb_max;
  const int lane = threadIdx.x & 31;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long first = cb + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  // warp-uniform trip count
  for (unsigned long long col0 = first - lane; col0 < ce; col0 += stride) {
    const unsigned long long col = col0 + lane;
    bool cand = false, nonneg = true;
    float dl = 0.f;
    if (col < ce && t.col_ub[col - cb] > lb) {         // only columns whose trivial bound beats the lower bound
      const sig_t* xs = x + (col * g.hop - off);
      const float c0 = 0.5f * (float)(win - 1);
      float s0 = 0.f;
      for (int n = 0; n < win; ++n) {
        const float y = s_w[n] * (float)xs[n];
        nonneg = nonneg && (y >= 0.f);
        s0 += y;
        dl = fmaf(fabsf((float)n - c0), fabsf(y), dl);
      }
      cand = true;
      if (nonneg) {
        // O(win) certificate: x = xbar + r gives sup_{w >= pi/(win-1)} |S(w)| <= rho * sum(w x) + sum w |x - xbar|
        const float xbar = s0 / s_wsum;
        float rs = 0.f;
        for (int n = 0; n < win; ++n) rs = fmaf(s_w[n], fabsf((float)xs[n] - xbar), rs);
        const float bnd = fmaf(g.rho, s0, rs);
        if (2.f * bnd * bnd * 1.00002f <= lb) { cand = false; atomicAdd(&P->n_refined, 1u); }
      }
    }
    unsigned mask = __ballot_sync(0xffffffffu, cand);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const unsigned long long ccol = col0 + src;
      const bool c_nonneg = __shfl_sync(0xffffffffu, (int)nonneg, src) != 0;
      const float c_dl = __shfl_sync(0xffffffffu, dl, src);
      bool fail = !c_nonneg;
      if (!fail && win > 2) {
        const sig_t* xs = x + (ccol * g.hop - off);
        const float w_lo = 3.14159265358979f / (float)(win - 1);
        const int G = 32 * win;
        const float delta = (3.14159265358979f - w_lo) / (float)G;
        const float slack = c_dl * delta * 0.5f;
        float worst = 0.f;
        for (int gi = lane; gi < G; gi += 32) {
          const float w = w_lo + ((float)gi + 0.5f) * delta;
          float sn, cs;
          sincosf(w, &sn, &cs);
          // Horner in z = exp(-jw)
          float ar = 0.f, ai = 0.f;
          for (int n = win - 1; n >= 0; --n) {
            const float tr = fmaf(ar, cs, ai * sn), ti = fmaf(ai, cs, -ar * sn);
            ar = tr + s_w[n] * (float)xs[n];
            ai = ti;
          }
          const float mag = sqrtf(fmaf(ar, ar, ai * ai)) + slack;
          worst = fmaxf(worst, mag);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, m));
        fail = 2.f * worst * worst * 1.00002f > lb;
      }
      if (lane == 0) {
        atomicAdd(&P->n_refined, 1u);
        if (fail) {
          const unsigned slot = atomicAdd(&P->n_hard, 1u);
          if (slot < t.hard_cap) t.hard_list[slot] = (unsigned)(ccol - cb);
          else P->valid = -5;
        }
      }
    }
  }
  // the last CTA to finish closes the search when no column needs the exhaustive scan
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&P->ticket_r, 1u) == gridDim.x - 1) {
      __threadfence();
      if (atomicAdd(&P->n_hard, 0u) == 0u) finalize_max(P, export_dst);
    }
  }
}

__device__ __forceinline__ void atomic_max_double_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// Exhaustive scan of the fine grid for the (rare) columns that have no certificate; float64.
__global__ void __launch_bounds__(256) stft_hard_kernel(StftTables t, StftGeom g, const sig_t* __restrict__ x, double* export_dst) {
  StftPlan* P = t.plan;
  if (P->valid <= 0) return;
  const unsigned nh = P->n_hard < t.hard_cap ? P->n_hard : t.hard_cap;
  if (nh == 0) return;
  __shared__ double s_y[1024];
  __shared__ double s_red[8];
  const int win = (int)g.win;
  const unsigned long long nfft = P->nfft, nyq = nfft / 2;
  double best = 0.0;
  for (unsigned h = 0; h < nh; ++h) {
    const unsigned long long col = P->col_begin + t.hard_list[h];
    const sig_t* xs = x + (col * g.hop - P->sample_offset);
    __syncthreads();
    for (int i = threadIdx.x; i < win; i += blockDim.x) s_y[i] = (double)t.win[i] * (double)xs[i];
    __syncthreads();
    for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j <= nyq;
         j += (unsigned long long)gridDim.x * blockDim.x) {
      double sn, cs;
      sincospi(2.0 * (double)j / (double)nfft, &sn, &cs);
      double ar = 0.0, ai = 0.0;
      for (int n = win - 1; n >= 0; --n) {
        const double tr = ar * cs + ai * sn, ti = ai * cs - ar * sn;
        ar = tr + s_y[n];
        ai = ti;
      }
      const double p2 = (ar * ar + ai * ai) * ((j == 0 || j == nyq) ? 1.0 : 2.0);
      best = p2 > best ? p2 : best;
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) { const double o = __shfl_xor_sync(0xffffffffu, best, m); best = o > best ? o : best; }
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) best = s_red[w] > best ? s_red[w] : best;
    if (best > 0.0) atomic_max_double_nonneg(&P->pmax_raw, best);
    __threadfence();
    if (atomicAdd(&P->ticket_h, 1u) == gridDim.x - 1) { __threadfence(); finalize_max(P, export_dst); }
  }
}

__global__ void stft_set_max_kernel(StftTables t, double v) { t.plan->pmax_raw = v; t.plan->task_counter = 0; }

// ------------------------------------------------------------------------------------------------
// main kernel: one thread = CPT spectrogram columns, all of one chunk's bins
// ------------------------------------------------------------------------------------------------
// shared-memory access through 32-bit shared-window addresses (keeps the address arithmetic out of the
// uniform datapath, which otherwise re-derives the generic base every iteration)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Writes the staged [COLS_W][QF] tile of one warp as full rows of QF consecutive log-frequency bins per
// column (time-major layout: 4*QF contiguous bytes per column).
template <int QF, int COLS_W>
__device__ __forceinline__ void flush_stage(uint32_t a_stage, int ncols_valid, float* __restrict__ out_warp,
                                            unsigned long long row_stride, int qbase, int nvalid, int lane) {
  constexpr int CPI = 32 / QF;          // columns written per iteration
  __syncwarp();
  const int sub = lane / QF, ql = lane % QF;
  float* ptr = out_warp + (unsigned long long)sub * row_stride + qbase + ql;
  uint32_t a = a_stage + (uint32_t)((sub * (QF + 1) + ql) * 4);
  if (ql < nvalid) {
#pragma unroll 4
    for (int c = sub; c < ncols_valid; c += CPI) {
      *ptr = lds32(a);
      ptr += CPI * row_stride;
      a += CPI * (QF + 1) * 4;
    }
  }
  __syncwarp();
}

template <int HALF, int CPT, int QF, int LAYOUT, int MAIN_THREADS, int MINB>
__global__ void __launch_bounds__(MAIN_THREADS, MINB)
stft_main_kernel(StftTables t, StftGeom g, const sig_t* __restrict__ x, float* __restrict__ out,
                 unsigned long long capacity_cols, unsigned long long ld_cols, int* d_err) {
  StftPlan* P = t.plan;
  if (P->valid <= 0) { if (threadIdx.x == 0 && blockIdx.x == 0 && P->valid < 0) *d_err = P->valid; return; }
  constexpr int WIN = 2 * HALF;
  constexpr int COLS_W = 32 * CPT;                 // columns per warp
  constexpr int COLS_B = MAIN_THREADS * CPT;       // columns per CTA task
  extern __shared__ __align__(16) float s_main[];
  float* s_coef = s_main;                                      // [NP_MAX][WIN]
  float2* s_meta = reinterpret_cast<float2*>(s_coef + NP_MAX * WIN);   // [NP_MAX] {K*log2(c_p), #queries completed}
  float* s_aq = reinterpret_cast<float*>(s_meta + NP_MAX);     // [MAX_NQ]
  float* s_ws = s_aq + MAX_NQ;                                 // [WIN]
  int* s_task = reinterpret_cast<int*>(s_ws + WIN);            // [4]
  float* s_stage = reinterpret_cast<float*>(s_task + 4);       // [warps][COLS_W][QF+1] (time-major layout only)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long cb = P->col_begin, ce = P->col_end, off = P->sample_offset;
  const unsigned long long ncl = ce - cb;
  if (ncl > capacity_cols) { if (tid == 0 && blockIdx.x == 0) *d_err = -4; return; }
  const int nq = P->nq, n_chunks = P->n_chunks;
  if (tid < WIN) s_ws[tid] = (float)((double)t.win[tid] / sqrt(P->pmax_raw));
  const unsigned long long n_cblk = (ncl + COLS_B - 1) / COLS_B;
  const long long n_tasks = (long long)(n_cblk * (unsigned long long)n_chunks);
  const uint32_t a_coef = smem_u32(s_coef), a_meta = smem_u32(s_meta), a_aq = smem_u32(s_aq);
  const uint32_t a_stage = smem_u32(s_stage) + (uint32_t)((LAYOUT == 0) ? warp * COLS_W * (QF + 1) * 4 : 0);
  const uint32_t a_st_lane = a_stage + (uint32_t)(lane * (QF + 1) * 4);

  for (;;) {
    __syncthreads();   // previous task is done with the tables and with s_task
    if (tid == 0) s_task[0] = (int)atomicAdd(&P->task_counter, 1u);
    __syncthreads();
    const long long task = s_task[0];
    if (task >= n_tasks) break;
    const int ch = (int)(task % n_chunks);
    const unsigned long long cblk = (unsigned long long)(task / n_chunks);
    const int q0 = P->chunk_q0[ch], q1 = P->chunk_q0[ch + 1];
    const int p0 = P->chunk_p0[ch];
    const int p1 = t.qpos[q1 - 1] + 1;
    const int np = p1 - p0 + 1;
    for (int i = tid; i < np * WIN / 4; i += MAIN_THREADS)
      reinterpret_cast<float4*>(s_coef)[i] = reinterpret_cast<const float4*>(t.coef + (size_t)p0 * WIN)[i];
    for (int i = tid; i < np; i += MAIN_THREADS) {
      int cnt = 0;
      if (i > 0) {
        int qa = t.qend[p0 + i - 1], qb = t.qend[p0 + i];
        qa = qa < q0 ? q0 : qa;
        qb = qb > q1 ? q1 : qb;
        cnt = qb > qa ? qb - qa : 0;
      }
      s_meta[i] = make_float2(t.kcb[p0 + i], __int_as_float(cnt));
    }
    for (int i = tid; i < q1 - q0; i += MAIN_THREADS) s_aq[i] = t.aq[q0 + i];
    __syncthreads();

    // window * x, folded into even / odd parts: |S|^2 = (sum e_m cos)^2 + (sum o_m sin)^2
    float e[CPT][HALF], o[CPT][HALF];
    const unsigned long long warp_col0 = cb + cblk * COLS_B + (unsigned long long)warp * COLS_W;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      unsigned long long col = warp_col0 + c * 32 + lane;
      if (col >= ce) col = ce - 1;
      const sig_t* xs = x + (col * g.hop - off);
#pragma unroll
      for (int m = 0; m < HALF; ++m) {
        const float ylo = s_ws[HALF - 1 - m] * (float)__ldg(xs + HALF - 1 - m), yhi = s_ws[HALF + m] * (float)__ldg(xs + HALF + m);
        e[c][m] = ylo + yhi;
        o[c][m] = ylo - yhi;
      }
    }
    const int ncols_valid = (warp_col0 >= ce) ? 0 : (int)((ce - warp_col0) < (unsigned long long)COLS_W ? (ce - warp_col0) : COLS_W);
    float* out_warp = out + (warp_col0 - cb) * (unsigned long long)nq;   // time-major
    float prev[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) prev[c] = 0.f;
    int qrel = 0;               // queries of this chunk emitted so far
    uint32_t a_cf = a_coef;

#pragma unroll 1
    for (int ip = 0; ip < np; ++ip, a_cf += WIN * 4) {
      float cs[WIN];
#pragma unroll
      for (int v = 0; v < WIN / 4; ++v) {
        const float4 f = lds128(a_cf + 16 * v);
        cs[4 * v] = f.x; cs[4 * v + 1] = f.y; cs[4 * v + 2] = f.z; cs[4 * v + 3] = f.w;
      }
      const float2 meta = lds64(a_meta + 8 * ip);
      float re[CPT], im[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) { re[c] = e[c][0] * cs[0]; im[c] = o[c][0] * cs[HALF]; }
#pragma unroll
      for (int m = 1; m < HALF; ++m) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          re[c] = fmaf(e[c][m], cs[m], re[c]);
          im[c] = fmaf(o[c][m], cs[HALF + m], im[c]);
        }
      }
      float db[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) db[c] = fmaf(K_DB, lg2_approx(fmaf(re[c], re[c], im[c] * im[c])), meta.x);
      const int cnt = __float_as_int(meta.y);
      for (int k = 0; k < cnt; ++k, ++qrel) {
        const float a = lds32(a_aq + 4 * qrel);
        if (LAYOUT == 0) {
          const int slot = qrel & (QF - 1);
#pragma unroll
          for (int c = 0; c < CPT; ++c) sts32(a_st_lane + (uint32_t)((c * 32 * (QF + 1) + slot) * 4), fmaf(a, db[c] - prev[c], prev[c]));
          if (slot == QF - 1) flush_stage<QF, COLS_W>(a_stage, ncols_valid, out_warp, (unsigned long long)nq, q0 + qrel - (QF - 1), QF, lane);
        } else {
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const unsigned long long col = warp_col0 + c * 32 + lane;
            if (col < ce) out[(unsigned long long)(q0 + qrel) * ld_cols + (col - cb)] = fmaf(a, db[c] - prev[c], prev[c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CPT; ++c) prev[c] = db[c];
    }
    if (LAYOUT == 0) {
      const int rem = qrel & (QF - 1);
      if (rem) flush_stage<QF, COLS_W>(a_stage, ncols_valid, out_warp, (unsigned long long)nq, q0 + qrel - rem, rem, lane);
    }
  }
}

// Generic window length (any win >= 2, any hop, odd lengths): one thread per column, the folded taps of the 128
// columns of a CTA in shared memory (thread-minor), bins in register tiles of 8 whose coefficients are staged
// as [tap][cos x 8 | sin x 8] (four broadcast LDS.128 per tap -> 16 FMAs).  The column mean is removed in float64
// and re-enters through the tabulated window response (same DC split as the tensor-core kernel).  Time-major
// output is staged per warp like in stft_main_kernel.  This is the path of the C5 window / hop sweep.
template <int LAYOUT>
__global__ void __launch_bounds__(128) stft_generic_serial_kernel(StftTables t, StftGeom g, const sig_t* __restrict__ x,
                                                           float* __restrict__ out, unsigned long long capacity_cols,
                                                           unsigned long long ld_cols, int* d_err) {
  const StftPlan* P = t.plan;
  if (P->valid <= 0) { if (threadIdx.x == 0 && blockIdx.x == 0 && P->valid < 0) *d_err = P->valid; return; }
  constexpr int QF = 16, BT = 8;
  extern __shared__ __align__(16) float s_dyn[];
  const int win = (int)g.win, half = win / 2, odd = win & 1;
  float* s_eo = s_dyn;                                   // [2*half][128] even / odd parts, thread-minor
  float* s_cf = s_eo + (size_t)2 * half * 128;           // [half][16] coefficient tile: cos of 8 bins | sin of 8 bins
  float* s_stage = s_cf + (size_t)half * 16;             // [4 warps][32][QF+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long cb = P->col_begin, ce = P->col_end, off = P->sample_offset;
  const unsigned long long ncl = ce - cb;
  if (ncl > capacity_cols) { if (tid == 0 && blockIdx.x == 0) *d_err = -4; return; }
  const int nq = P->nq, nb = P->nb;
  const float inv = (float)(1.0 / sqrt(P->pmax_raw));
  const unsigned long long n_blk = (ncl + 127) / 128;
  const uint32_t a_stage = smem_u32(s_stage) + (uint32_t)(warp * 32 * (QF + 1) * 4);
  const uint32_t a_st_lane = a_stage + (uint32_t)(lane * (QF + 1) * 4);
  for (unsigned long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
    unsigned long long col = cb + blk * 128 + tid;
    const bool act = col < ce;
    if (!act) col = ce - 1;
    const sig_t* xs = x + (col * g.hop - off);
    double mean_d = 0.0;
    for (int n = 0; n < win; ++n) mean_d += xs[n];
    mean_d /= (double)win;
    const float mean = (float)mean_d;
    for (int m = 0; m < half; ++m) {
      const int lo = half - 1 - m, hi = odd ? (half + 1 + m) : (half + m);
      const float ylo = t.win[lo] * inv * (float)(xs[lo] - (double)mean), yhi = t.win[hi] * inv * (float)(xs[hi] - (double)mean);
      s_eo[m * 128 + tid] = ylo + yhi;
      s_eo[(half + m) * 128 + tid] = ylo - yhi;
    }
    const float yc = odd ? t.win[half] * inv * (float)(xs[half] - (double)mean) : 0.f;
    const float xb = mean * inv;
    const unsigned long long warp_col0 = cb + blk * 128 + (unsigned long long)warp * 32;
    const int ncols_valid = (warp_col0 >= ce) ? 0 : (int)((ce - warp_col0) < 32ull ? (ce - warp_col0) : 32ull);
    float* out_warp = out + (warp_col0 - cb) * (unsigned long long)nq;
    float prev = 0.f;
    int qcur = 0;
    for (int pb = 0; pb < nb; pb += BT) {
      __syncthreads();                                   // previous coefficient tile consumed
      for (int i = tid; i < half * 16; i += 128) {
        const int m = i >> 4, j = i & 15, p = pb + (j & 7);
        s_cf[i] = (p < nb) ? t.coef[(size_t)p * 2 * half + ((j < 8) ? m : half + m)] : 0.f;
      }
      __syncthreads();
      float re[BT], im[BT];
#pragma unroll
      for (int j = 0; j < BT; ++j) { re[j] = yc; im[j] = 0.f; }
      for (int m = 0; m < half; ++m) {
        const float e = s_eo[m * 128 + tid], o = s_eo[(half + m) * 128 + tid];
        const float4 c0 = *reinterpret_cast<const float4*>(s_cf + m * 16), c1 = *reinterpret_cast<const float4*>(s_cf + m * 16 + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(s_cf + m * 16 + 8), s1 = *reinterpret_cast<const float4*>(s_cf + m * 16 + 12);
        re[0] = fmaf(e, c0.x, re[0]); re[1] = fmaf(e, c0.y, re[1]); re[2] = fmaf(e, c0.z, re[2]); re[3] = fmaf(e, c0.w, re[3]);
        re[4] = fmaf(e, c1.x, re[4]); re[5] = fmaf(e, c1.y, re[5]); re[6] = fmaf(e, c1.z, re[6]); re[7] = fmaf(e, c1.w, re[7]);
        im[0] = fmaf(o, s0.x, im[0]); im[1] = fmaf(o, s0.y, im[1]); im[2] = fmaf(o, s0.z, im[2]); im[3] = fmaf(o, s0.w, im[3]);
        im[4] = fmaf(o, s1.x, im[4]); im[5] = fmaf(o, s1.y, im[5]); im[6] = fmaf(o, s1.z, im[6]); im[7] = fmaf(o, s1.w, im[7]);
      }
#pragma unroll
      for (int j = 0; j < BT; ++j) {
        const int p = pb + j;
        if (p < nb) {
          const float r = fmaf(xb, __ldg(t.wdc + p), re[j]);          // + mean * window DC response
          const float db = fmaf(K_DB, lg2_approx(fmaf(r, r, im[j] * im[j])), __ldg(t.kcb + p));
          if (p > 0) {
            const int qe = __ldg(t.qend + p);
            for (; qcur < qe; ++qcur) {
              const float v = fmaf(__ldg(t.aq + qcur), db - prev, prev);
              if (LAYOUT == 0) {
                const int slot = qcur & (QF - 1);
                sts32(a_st_lane + (uint32_t)(slot * 4), v);
                if (slot == QF - 1) flush_stage<QF, 32>(a_stage, ncols_valid, out_warp, (unsigned long long)nq, qcur - (QF - 1), QF, lane);
              } else if (act) {
                out[(unsigned long long)qcur * ld_cols + (col - cb)] = v;
              }
            }
          }
          prev = db;
        }
      }
    }
    if (LAYOUT == 0) {
      const int rem = qcur & (QF - 1);
      if (rem) flush_stage<QF, 32>(a_stage, ncols_valid, out_warp, (unsigned long long)nq, qcur - rem, rem, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// device-side hand-offs of the sharded path (no host round trip between the collectives)
// ------------------------------------------------------------------------------------------------
// Generic window length up to 256 (any hop, odd lengths), the C5 sweep: a CTA task = 128 columns x one query chunk,
// handed out by an atomic counter so that recordings with few columns still fill the GPU.  The folded taps of the
// 128 columns sit in shared memory ([tap][column]); bins are walked in passes of 32: warp w owns 8 bins, a lane 4
// columns (2 LDS.128 of taps + 4 broadcast LDS.128 of coefficients -> 64 FMAs per tap), the 32 x 128 dB values go
// through shared memory to the column-owning threads, which run the interp1 and stage time-major rows as the
// other CUDA-core kernels do.  Mean removal in float64 and the DC term as mean * W(w_bin) as everywhere else.
// ------------------------------------------------------------------------------------------------
constexpr int GEN_COLS = 128, GEN_PB = 32, GEN_CFS = 36, GEN_DBS = 33, GEN_QF = 16;
constexpr int GEN_MAX_WIN = 256;

static size_t generic_tiled_smem(uint32_t win) {
  const size_t half = win / 2;
  return (2 * half * GEN_COLS + 2 * half * GEN_CFS + GEN_COLS * GEN_DBS + 4 * 32 * (GEN_QF + 1) + 2 * GEN_COLS + 4) * sizeof(float);
}

template <int LAYOUT>
__global__ void __launch_bounds__(128) stft_generic_tiled_kernel(StftTables t, StftGeom g, const sig_t* __restrict__ x,
                                                                 float* __restrict__ out, unsigned long long capacity_cols,
                                                                 unsigned long long ld_cols, int* d_err) {
  StftPlan* P = t.plan;
  if (P->valid <= 0) { if (threadIdx.x == 0 && blockIdx.x == 0 && P->valid < 0) *d_err = P->valid; return; }
  constexpr int QF = GEN_QF;
  extern __shared__ __align__(16) float s_dyn[];
  const int win = (int)g.win, half = win / 2, odd = win & 1;
  float* s_eo = s_dyn;                                     // [2*half][128] even / odd parts, column-minor
  float* s_cf = s_eo + (size_t)2 * half * GEN_COLS;        // [2*half][GEN_CFS] cos (rows < half) / sin of the pass's 32 bins
  float* s_db = s_cf + (size_t)2 * half * GEN_CFS;         // [128][GEN_DBS] dB of the pass, column-major
  float* s_stage = s_db + GEN_COLS * GEN_DBS;              // [4 warps][32][QF+1]
  float* s_xb = s_stage + 4 * 32 * (QF + 1);               // [128] mean * inv
  float* s_yc = s_xb + GEN_COLS;                           // [128] centre tap of odd windows
  int* s_task = reinterpret_cast<int*>(s_yc + GEN_COLS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long cb = P->col_begin, ce = P->col_end, off = P->sample_offset;
  const unsigned long long ncl = ce - cb;
  if (ncl > capacity_cols) { if (tid == 0 && blockIdx.x == 0) *d_err = -4; return; }
  const int nq = P->nq, nb = P->nb, n_chunks = P->n_chunks;
  const float inv = (float)(1.0 / sqrt(P->pmax_raw));
  const unsigned long long n_blk = (ncl + GEN_COLS - 1) / GEN_COLS;
  const long long n_tasks = (long long)(n_blk * (unsigned long long)n_chunks);
  const uint32_t a_stage = smem_u32(s_stage) + (uint32_t)(warp * 32 * (QF + 1) * 4);
  const uint32_t a_st_lane = a_stage + (uint32_t)(lane * (QF + 1) * 4);
  for (;;) {
    __syncthreads();                                       // previous task is done with every shared array
    if (tid == 0) s_task[0] = (int)atomicAdd(&P->task_counter, 1u);
    __syncthreads();
    const long long task = s_task[0];
    if (task >= n_tasks) break;
    const int ch = (int)(task % n_chunks);
    const unsigned long long blk = (unsigned long long)(task / n_chunks);
    const int q0 = P->chunk_q0[ch], q1 = P->chunk_q0[ch + 1];
    const int p0 = P->chunk_p0[ch];
    const int np = t.qpos[q1 - 1] + 1 - p0 + 1;            // bin positions [p0, p0 + np)
    // ---- folded, windowed, mean-free taps of this thread's column ----
    unsigned long long col = cb + blk * GEN_COLS + tid;
    const bool act = col < ce;
    if (!act) col = ce - 1;
    {
      const sig_t* xs = x + (col * g.hop - off);
      double mean_d = 0.0;
      for (int n = 0; n < win; ++n) mean_d += xs[n];
      mean_d /= (double)win;
      const float mean = (float)mean_d;
      for (int m = 0; m < half; ++m) {
        const int lo = half - 1 - m, hi = odd ? (half + 1 + m) : (half + m);
        const float ylo = t.win[lo] * inv * (float)(xs[lo] - (double)mean), yhi = t.win[hi] * inv * (float)(xs[hi] - (double)mean);
        s_eo[m * GEN_COLS + tid] = ylo + yhi;
        s_eo[(half + m) * GEN_COLS + tid] = ylo - yhi;
      }
      s_yc[tid] = odd ? t.win[half] * inv * (float)(xs[half] - (double)mean) : 0.f;
      s_xb[tid] = mean * inv;
    }
    const unsigned long long warp_col0 = cb + blk * GEN_COLS + (unsigned long long)warp * 32;
    const int ncols_valid = (warp_col0 >= ce) ? 0 : (int)((ce - warp_col0) < 32ull ? (ce - warp_col0) : 32ull);
    float* out_warp = out + (warp_col0 - cb) * (unsigned long long)nq;
    float prev = 0.f;
    int qcur = q0;
    for (int pb = 0; pb < np; pb += GEN_PB) {
      __syncthreads();                                     // taps complete; previous pass is done with s_cf and s_db
      for (int i = tid; i < 2 * half * GEN_PB; i += 128) {
        const int pp = i / (2 * half), m = i - pp * 2 * half, p = p0 + pb + pp;
        s_cf[m * GEN_CFS + pp] = (pb + pp < np) ? t.coef[(size_t)p * 2 * half + m] : 0.f;
      }
      __syncthreads();
      {
        float re[4][8], im[4][8];
        const float4 yc4 = *reinterpret_cast<const float4*>(s_yc + 4 * lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          re[0][j] = yc4.x; re[1][j] = yc4.y; re[2][j] = yc4.z; re[3][j] = yc4.w;
          im[0][j] = 0.f; im[1][j] = 0.f; im[2][j] = 0.f; im[3][j] = 0.f;
        }
        const float* pe = s_eo + 4 * lane;
        const float* pc = s_cf + 8 * warp;
#pragma unroll 2
        for (int m = 0; m < half; ++m) {
          const float4 e4 = *reinterpret_cast<const float4*>(pe + m * GEN_COLS);
          const float4 o4 = *reinterpret_cast<const float4*>(pe + (half + m) * GEN_COLS);
          const float4 c0 = *reinterpret_cast<const float4*>(pc + m * GEN_CFS), c1 = *reinterpret_cast<const float4*>(pc + m * GEN_CFS + 4);
          const float4 s0 = *reinterpret_cast<const float4*>(pc + (half + m) * GEN_CFS), s1 = *reinterpret_cast<const float4*>(pc + (half + m) * GEN_CFS + 4);
          const float ev[4] = {e4.x, e4.y, e4.z, e4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w};
          const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { re[i][j] = fmaf(ev[i], cv[j], re[i][j]); im[i][j] = fmaf(ov[i], sv[j], im[i][j]); }
        }
        const float4 xb4 = *reinterpret_cast<const float4*>(s_xb + 4 * lane);
        const float xbv[4] = {xb4.x, xb4.y, xb4.z, xb4.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pr = pb + 8 * warp + j;
          const int p = p0 + (pr < np ? pr : np - 1);
          const float wd = __ldg(t.wdc + p), kc = __ldg(t.kcb + p);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float r = fmaf(xbv[i], wd, re[i][j]);          // + mean * window DC response
            s_db[(4 * lane + i) * GEN_DBS + 8 * warp + j] = fmaf(K_DB, lg2_approx(fmaf(r, r, im[i][j] * im[i][j])), kc);
          }
        }
      }
      __syncthreads();
      // ---- interp1 of this thread's column over the pass's bins ----
      const int nbp = (np - pb) < GEN_PB ? (np - pb) : GEN_PB;
      for (int b = 0; b < nbp; ++b) {
        const int p = p0 + pb + b;
        const float db = s_db[tid * GEN_DBS + b];
        if (pb + b > 0) {
          int qe = __ldg(t.qend + p);
          qe = qe > q1 ? q1 : qe;
          for (; qcur < qe; ++qcur) {
            const float v = fmaf(__ldg(t.aq + qcur), db - prev, prev);
            if (LAYOUT == 0) {
              const int slot = qcur & (QF - 1);
              sts32(a_st_lane + (uint32_t)(slot * 4), v);
              if (slot == QF - 1) flush_stage<QF, 32>(a_stage, ncols_valid, out_warp, (unsigned long long)nq, qcur - (QF - 1), QF, lane);
            } else if (act) {
              out[(unsigned long long)qcur * ld_cols + (col - cb)] = v;
            }
          }
        }
        prev = db;
      }
    }
    if (LAYOUT == 0) {
      const int rem = qcur & (QF - 1);
      if (rem) flush_stage<QF, 32>(a_stage, ncols_valid, out_warp, (unsigned long long)nq, qcur - rem, rem, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// msg = {L_local, first win-1 samples} in float64: the shard header that is all-gathered
__global__ void shard_pack_kernel(const sig_t* __restrict__ xc, const unsigned long long* __restrict__ d_ndet, uint32_t PN,
                                  uint32_t win, double* __restrict__ msg) {
  const unsigned long long L = *d_ndet * PN;
  const uint32_t i = threadIdx.x;
  if (i == 0) msg[0] = (double)L;
  if (i < win - 1) msg[1 + i] = (i < L) ? xc[i] : 0.0;
}

__global__ void stft_set_max_dev_kernel(StftTables t, const double* src) { t.plan->pmax_raw = *src; t.plan->task_counter = 0; }

cudaError_t launch_shard_pack(const sig_t* xc, const unsigned long long* d_ndet, uint32_t PN, uint32_t win, double* msg,
                              cudaStream_t st) {
  shard_pack_kernel<<<1, 1024, 0, st>>>(xc, d_ndet, PN, win, msg);
  return cudaGetLastError();
}
cudaError_t launch_stft_set_max_dev(const StftTables& t, const double* src, cudaStream_t st) {
  stft_set_max_dev_kernel<<<1, 1, 0, st>>>(t, src);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int sm_count() { return device_sm_count(); }

int stft_variant() {
  static const int variant = env_int("FMCW_STFT_VARIANT", -1);
  return variant;
}

cudaError_t launch_stft_plan(const StftTables& t, const StftGeom& g, const unsigned long long* d_ndet, uint32_t PN,
                             unsigned long long L_total_host, unsigned long long sample_offset,
                             unsigned long long L_local_host, unsigned long long L_avail_host, int n_chunks,
                             cudaStream_t st, const double* gathered, uint32_t world, uint32_t rank, sig_t* xc, int spec_mode,
                             const unsigned long long* wait_flags, const unsigned long long* wait_step) {
  const bool tc = (g.win == 20 && stft_variant() < 0);
  stft_plan_kernel<<<1, 1024, 0, st>>>(t, g, d_ndet, PN, L_total_host, sample_offset, L_local_host, L_avail_host, n_chunks,
                                       tc ? 2 : 0, gathered, world, rank, xc, spec_mode, wait_flags, wait_step);
  if (!tc) { stft_coef_kernel<<<128, 256, 0, st>>>(t, g, spec_mode); return cudaGetLastError(); }
  return launch_stft_tc_prepare(t, g, t.tcB, t.nb_max, st, spec_mode);
  return cudaGetLastError();
}

cudaError_t launch_stft_max(const StftTables& t, const StftGeom& g, const sig_t* x, cudaStream_t st, double* export_dst) {
  const int grid = sm_count() * 8;
  stft_colstat_kernel<<<grid, 256, 0, st>>>(t, g, x);
  stft_refine_kernel<<<grid, 256, 0, st>>>(t, g, x, export_dst);
  stft_hard_kernel<<<sm_count() * 4, 256, 0, st>>>(t, g, x, export_dst);
  return cudaGetLastError();
}

cudaError_t launch_stft_set_max(const StftTables& t, double pmax_raw, cudaStream_t st) {
  stft_set_max_kernel<<<1, 1, 0, st>>>(t, pmax_raw);
  return cudaGetLastError();
}

template <int HALF, int CPT, int QF, int THREADS, int MINB>
static cudaError_t launch_main_variant(const StftTables& t, const StftGeom& g, const sig_t* x, float* out,
                                       unsigned long long capacity_cols, unsigned long long ld_cols, int layout, int* d_err,
                                       cudaStream_t st, int sms) {
  const size_t base = (size_t)(NP_MAX * 2 * HALF + 2 * NP_MAX + MAX_NQ + 2 * HALF + 4) * sizeof(float);
  cudaError_t e;
  if (layout == 0) {
    const size_t smem = base + (size_t)(THREADS / 32) * 32 * CPT * (QF + 1) * sizeof(float);
    e = cudaFuncSetAttribute(stft_main_kernel<HALF, CPT, QF, 0, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    stft_main_kernel<HALF, CPT, QF, 0, THREADS, MINB><<<sms * MINB, THREADS, smem, st>>>(t, g, x, out, capacity_cols, ld_cols, d_err);
  } else {
    e = cudaFuncSetAttribute(stft_main_kernel<HALF, CPT, QF, 1, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base);
    if (e != cudaSuccess) return e;
    stft_main_kernel<HALF, CPT, QF, 1, THREADS, MINB><<<sms * MINB, THREADS, base, st>>>(t, g, x, out, capacity_cols, ld_cols, d_err);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Fine-grid PSD band (RP:283 restricted to rows first_bin + i*bin_step): what surf(T, F, psd) at RP:333 draws between its
// ylim.  Float64 DTFT like stft_precise_kernel (a CTA owns 128 columns, their windowed taps in shared memory, cos / sin of a
// tile of 8 rows by sincospi of the exactly reduced argument); one-sided doubling (RP:276) and the 1/max(P) of the plan.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stft_finegrid_kernel(StftTables t, StftGeom g, const sig_t* __restrict__ x,
                                                            float* __restrict__ psd, unsigned long long first_bin,
                                                            unsigned long long bin_step, unsigned n_rows,
                                                            unsigned long long capacity_cols, int* d_err) {
  const StftPlan* P = t.plan;
  if (P->valid <= 0) { if (threadIdx.x == 0 && blockIdx.x == 0 && P->valid < 0) *d_err = P->valid; return; }
  constexpr int BT = 8, NC = 128;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int win = (int)g.win;
  double* s_y = reinterpret_cast<double*>(s_raw);               // [win][NC] windowed, normalised taps
  double* s_cs = s_y + (size_t)win * NC;                        // [win][2*BT] cos | sin of (n - c) * w for 8 rows
  const int tid = threadIdx.x;
  const unsigned long long cb = P->col_begin, ce = P->col_end, off = P->sample_offset;
  const unsigned long long ncl = ce - cb;
  if (ncl > capacity_cols) { if (tid == 0 && blockIdx.x == 0) *d_err = -4; return; }
  const unsigned long long nfft = P->nfft;
  const double inv = 1.0 / sqrt(P->pmax_raw);
  const unsigned long long n_blk = (ncl + NC - 1) / NC;
  for (unsigned long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
    unsigned long long col = cb + blk * NC + tid;
    const bool act = col < ce;
    if (!act) col = ce - 1;
    __syncthreads();
    const sig_t* xs = x + (col * g.hop - off);
    for (int n = 0; n < win; ++n) s_y[(size_t)n * NC + tid] = t.win_d[n] * inv * xs[n];
    for (unsigned r0 = 0; r0 < n_rows; r0 += BT) {
      __syncthreads();
      for (int i = tid; i < win * BT; i += NC) {
        const int n = i / BT, j = i - n * BT;
        double sv = 0.0, cv = 0.0;
        if (r0 + j < n_rows) {
          const unsigned long long bin = first_bin + (unsigned long long)(r0 + j) * bin_step;
          const unsigned long long r = ((unsigned long long)n * bin) % nfft;      // exact reduction of n * bin / nfft turns
          sincospi(2.0 * (double)r / (double)nfft, &sv, &cv);
        }
        s_cs[(size_t)n * 2 * BT + j] = cv;
        s_cs[(size_t)n * 2 * BT + BT + j] = sv;
      }
      __syncthreads();
      double re[BT], im[BT];
#pragma unroll
      for (int j = 0; j < BT; ++j) { re[j] = 0.0; im[j] = 0.0; }
      for (int n = 0; n < win; ++n) {
        const double y = s_y[(size_t)n * NC + tid];
        const double* cs = s_cs + (size_t)n * 2 * BT;
#pragma unroll
        for (int j = 0; j < BT; ++j) { re[j] = fma(y, cs[j], re[j]); im[j] = fma(y, cs[BT + j], im[j]); }
      }
      if (act) {
        float* dst = psd + (col - cb) * (unsigned long long)n_rows + r0;
#pragma unroll
        for (int j = 0; j < BT; ++j) {
          if (r0 + j < n_rows) {
            const unsigned long long bin = first_bin + (unsigned long long)(r0 + j) * bin_step;
            const float cdb = (bin == 0 || bin == nfft / 2) ? 0.f : K_DB;
            dst[j] = fmaf(K_DB, lg2_approx((float)fma(re[j], re[j], im[j] * im[j])), cdb);
          }
        }
      }
    }
  }
}

cudaError_t launch_stft_finegrid(const StftTables& t, const StftGeom& g, const sig_t* x, float* psd, unsigned long long first_bin,
                                 unsigned long long bin_step, unsigned n_rows, unsigned long long capacity_cols, int* d_err,
                                 cudaStream_t st) {
  const size_t smem = ((size_t)g.win * 128 + (size_t)g.win * 16) * sizeof(double);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(stft_finegrid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  stft_finegrid_kernel<<<sm_count() * per_sm, 128, smem, st>>>(t, g, x, psd, first_bin, bin_step, n_rows, capacity_cols, d_err);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Float64 STFT (FMCW_OPT_STFT_PRECISION = 1): the reference arithmetic (RP:276-299 in double) for callers that need the
// 1e-4 relative tolerance on bins far below the -140 dB the TF32 x 2 tensor-core kernel guarantees.  Any window, any hop.
// A CTA owns NC columns; the folded, windowed taps of its columns sit in shared memory as float64; the cos / sin of a
// tile of 8 bins are evaluated once per CTA by sincospi of the exactly reduced argument; a thread owns a column and walks
// the sorted distinct bins, running the interp1 as the bins complete (same bookkeeping as the generic-window kernels).
// ------------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(128) stft_precise_kernel(StftTables t, StftGeom g, const sig_t* __restrict__ x,
                                                           float* __restrict__ out, unsigned long long capacity_cols,
                                                           unsigned long long ld_cols, int* d_err, int NC) {
  const StftPlan* P = t.plan;
  if (P->valid <= 0) { if (threadIdx.x == 0 && blockIdx.x == 0 && P->valid < 0) *d_err = P->valid; return; }
  constexpr int QF = 16, BT = 8;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int win = (int)g.win, half = win / 2, odd = win & 1;
  double* s_eo = reinterpret_cast<double*>(s_raw);              // [2*half + 1][NC]: even parts, odd parts, centre tap
  double* s_cf = s_eo + (size_t)(2 * half + 1) * NC;            // [half][2*BT] cos of 8 bins | sin of 8 bins
  float* s_stage = reinterpret_cast<float*>(s_cf + (size_t)half * 2 * BT);   // [warps][32][QF+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long cb = P->col_begin, ce = P->col_end, off = P->sample_offset;
  const unsigned long long ncl = ce - cb;
  if (ncl > capacity_cols) { if (tid == 0 && blockIdx.x == 0) *d_err = -4; return; }
  const int nq = P->nq, nb = P->nb;
  const unsigned long long nfft = P->nfft;
  const long long mod = (long long)(2 * nfft);
  const double inv = 1.0 / sqrt(P->pmax_raw);
  const unsigned long long n_blk = (ncl + NC - 1) / NC;
  const uint32_t a_stage = smem_u32(s_stage) + (uint32_t)(warp * 32 * (QF + 1) * 4);
  const uint32_t a_st_lane = a_stage + (uint32_t)(lane * (QF + 1) * 4);
  const bool owner = tid < NC;
  for (unsigned long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
    unsigned long long col = cb + blk * NC + tid;
    const bool act = owner && col < ce;
    if (!(col < ce)) col = ce - 1;
    __syncthreads();                                     // the previous block's taps are consumed
    if (owner) {
      const sig_t* xs = x + (col * g.hop - off);
      for (int m = 0; m < half; ++m) {
        const int lo = half - 1 - m, hi = odd ? (half + 1 + m) : (half + m);
        const double ylo = t.win_d[lo] * inv * xs[lo], yhi = t.win_d[hi] * inv * xs[hi];
        s_eo[(size_t)m * NC + tid] = ylo + yhi;
        s_eo[(size_t)(half + m) * NC + tid] = ylo - yhi;
      }
      s_eo[(size_t)(2 * half) * NC + tid] = odd ? t.win_d[half] * inv * xs[half] : 0.0;
    }
    const unsigned long long warp_col0 = cb + blk * NC + (unsigned long long)warp * 32;
    const int ncols_valid = (warp * 32 >= NC || warp_col0 >= ce) ? 0 : (int)((ce - warp_col0) < 32ull ? (ce - warp_col0) : 32ull);
    float* out_warp = out + (warp_col0 - cb) * (unsigned long long)nq;
    float prev = 0.f;
    int qcur = 0;
    for (int pb = 0; pb < nb; pb += BT) {
      __syncthreads();                                   // previous coefficient tile consumed (and the taps written)
      for (int i = tid; i < half * BT; i += blockDim.x) {
        const int m = i / BT, j = i - m * BT, p = pb + j;
        double sv = 0.0, cv = 0.0;
        if (p < nb) {
          // phase (m + d) * w with d = 1/2 (even windows) or 1 (odd windows, centre tap apart): 2*pi*(2m + 2d)*bin / (2 nfft)
          const long long r = ((long long)(2 * m + (odd ? 2 : 1)) * (long long)t.bins[p]) % mod;
          sincospi((double)r / (double)nfft, &sv, &cv);
        }
        s_cf[(size_t)m * 2 * BT + j] = cv;
        s_cf[(size_t)m * 2 * BT + BT + j] = sv;
      }
      __syncthreads();
      if (warp * 32 < NC) {
        double re[BT], im[BT];
        const double yc = s_eo[(size_t)(2 * half) * NC + (owner ? tid : 0)];
#pragma unroll
        for (int j = 0; j < BT; ++j) { re[j] = yc; im[j] = 0.0; }
        for (int m = 0; m < half; ++m) {
          const double e = s_eo[(size_t)m * NC + tid], o = s_eo[(size_t)(half + m) * NC + tid];
          const double* cf = s_cf + (size_t)m * 2 * BT;
#pragma unroll
          for (int j = 0; j < BT; ++j) { re[j] = fma(e, cf[j], re[j]); im[j] = fma(o, cf[BT + j], im[j]); }
        }
#pragma unroll
        for (int j = 0; j < BT; ++j) {
          const int p = pb + j;
          if (p < nb) {
            const float pw = (float)fma(re[j], re[j], im[j] * im[j]);
            const float db = fmaf(K_DB, lg2_approx(pw), __ldg(t.kcb + p));
            if (p > 0) {
              const int qe = __ldg(t.qend + p);
              for (; qcur < qe; ++qcur) {
                const float v = fmaf(__ldg(t.aq + qcur), db - prev, prev);
                if (LAYOUT == 0) {
                  const int slot = qcur & (QF - 1);
                  sts32(a_st_lane + (uint32_t)(slot * 4), v);
                  if (slot == QF - 1) flush_stage<QF, 32>(a_stage, ncols_valid, out_warp, (unsigned long long)nq, qcur - (QF - 1), QF, lane);
                } else if (act) {
                  out[(unsigned long long)qcur * ld_cols + (col - cb)] = v;
                }
              }
            }
            prev = db;
          }
        }
      }
    }
    if (LAYOUT == 0 && warp * 32 < NC) {
      const int rem = qcur & (QF - 1);
      if (rem) flush_stage<QF, 32>(a_stage, ncols_valid, out_warp, (unsigned long long)nq, qcur - rem, rem, lane);
    }
  }
}

static cudaError_t launch_stft_precise(const StftTables& t, const StftGeom& g, const sig_t* x, float* out,
                                       unsigned long long capacity_cols, unsigned long long ld_cols, int layout, int* d_err,
                                       cudaStream_t st, int sms) {
  const int half = (int)g.win / 2;
  const int NC = g.win <= 96 ? 128 : g.win <= 192 ? 64 : 32;
  const size_t smem = ((size_t)(2 * half + 1) * NC + (size_t)half * 16) * sizeof(double) + (size_t)4 * 32 * 17 * sizeof(float);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  cudaError_t e;
  if (layout == 0) {
    e = cudaFuncSetAttribute(stft_precise_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    stft_precise_kernel<0><<<sms * per_sm, 128, smem, st>>>(t, g, x, out, capacity_cols, ld_cols, d_err, NC);
  } else {
    e = cudaFuncSetAttribute(stft_precise_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    stft_precise_kernel<1><<<sms * per_sm, 128, smem, st>>>(t, g, x, out, capacity_cols, ld_cols, d_err, NC);
  }
  return cudaGetLastError();
}

cudaError_t launch_stft_main(const StftTables& t, const StftGeom& g, const sig_t* x, float* out,
                             unsigned long long capacity_cols, unsigned long long ld_cols, int layout, int* d_err,
                             cudaStream_t st, const double* gmax_dev, int precise) {
  const int sms = sm_count();
  if (precise) {
    if (gmax_dev) {
      cudaError_t e0 = launch_stft_set_max_dev(t, gmax_dev, st);
      if (e0 != cudaSuccess) return e0;
    }
    return launch_stft_precise(t, g, x, out, capacity_cols, ld_cols, layout, d_err, st, sms);
  }
  if (g.win == 20 && stft_variant() < 0)
    return launch_stft_tc_main(t, g, x, out, t.tcB, capacity_cols, ld_cols, layout, d_err, st, gmax_dev);
  if (gmax_dev) {
    cudaError_t e0 = launch_stft_set_max_dev(t, gmax_dev, st);
    if (e0 != cudaSuccess) return e0;
  }
  if (g.win == 20) {
    const int variant = stft_variant();
    switch (variant) {
      case 1: return launch_main_variant<10, 4, 16, 128, 3>(t, g, x, out, capacity_cols, ld_cols, layout, d_err, st, sms);
      case 2: return launch_main_variant<10, 2, 32, 256, 2>(t, g, x, out, capacity_cols, ld_cols, layout, d_err, st, sms);
      default: return launch_main_variant<10, 4, 16, 256, 2>(t, g, x, out, capacity_cols, ld_cols, layout, d_err, st, sms);
      case 4: return launch_main_variant<10, 1, 16, 256, 4>(t, g, x, out, capacity_cols, ld_cols, layout, d_err, st, sms);
      case 0: return launch_main_variant<10, 2, 16, 256, 3>(t, g, x, out, capacity_cols, ld_cols, layout, d_err, st, sms);
    }
  } else if (g.win <= (uint32_t)GEN_MAX_WIN) {
    const size_t smem = generic_tiled_smem(g.win);
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    cudaError_t e;
    if (layout == 0) {
      e = cudaFuncSetAttribute(stft_generic_tiled_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      stft_generic_tiled_kernel<0><<<sms * per_sm, 128, smem, st>>>(t, g, x, out, capacity_cols, ld_cols, d_err);
    } else {
      e = cudaFuncSetAttribute(stft_generic_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      stft_generic_tiled_kernel<1><<<sms * per_sm, 128, smem, st>>>(t, g, x, out, capacity_cols, ld_cols, d_err);
    }
  } else {
    const size_t smem = ((size_t)(2 * (g.win / 2)) * 128 + (size_t)(g.win / 2) * 16 + 4 * 32 * 17) * sizeof(float);
    cudaError_t e;
    if (layout == 0) {
      e = cudaFuncSetAttribute(stft_generic_serial_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      stft_generic_serial_kernel<0><<<sms * 4, 128, smem, st>>>(t, g, x, out, capacity_cols, ld_cols, d_err);
    } else {
      e = cudaFuncSetAttribute(stft_generic_serial_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      stft_generic_serial_kernel<1><<<sms * 4, 128, smem, st>>>(t, g, x, out, capacity_cols, ld_cols, d_err);
    }
  }
  return cudaGetLastError();
}

}  // namespace fmcw
