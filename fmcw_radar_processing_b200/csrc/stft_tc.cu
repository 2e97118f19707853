// STFT main kernel on the 5th-generation tensor cores (tcgen05 + TMEM), window_length = 20.
//
// The 20-tap windowed DTFT at the ~1.1k-1.7k planned bins (see stft.cu) is the contraction
//     Re[col][bin] = sum_k E[col][k] * C[bin][k],   Im[col][bin] = sum_k O[col][k] * S[bin][k]
// with E/O the even/odd folded, windowed taps of a spectrogram column.  On the fp32 pipe this costs 20 FMAs
// per 4-byte output and bounds the kernel at ~50 % FMA-pipe utilisation (profiles/ncu_full_r1b.txt).  Here:
//   * one CTA tile = 128 spectrogram columns (the M = 128 rows of the UMMA tile, one TMEM lane each);
//   * bins are walked in chunks of 128 (N = 128): Re and Im accumulators = 256 TMEM columns;
//   * operands are split hi + lo in TF32 (cvt.rna) and three products hi*hi + hi*lo + lo*hi are accumulated
//     in fp32 -- ~2^-22 relative, the float32 class of the CUDA-core kernel (single-pass TF32/BF16 would not
//     meet the 1e-3 dB tolerance);
//   * the column mean is removed before the split and returns as an extra tap (k = 10: mean * window DC
//     response W(w_bin), tabulated in float64), so the large DC term is never rounded together with the
//     small residual;
//   * K = 10 (+1) taps are padded to 16 = two K = 8 TF32 instructions; 12 tcgen05.mma per chunk;
//   * A tiles are written to shared memory by the epilogue threads (K-major, no swizzle, core matrices of
//     8 rows x 16 bytes); B tiles (planned once on the device) arrive by 1-D bulk copy (cp.async.bulk) on an
//     mbarrier; tcgen05.commit signals the epilogue, which reads the accumulators with tcgen05.ld, takes
//     |S|^2 -> lg2 -> dB, runs the interp1 onto the log-frequency axis and writes the spectrogram.
// Warp roles per CTA (192 threads, two CTAs per SM so that one CTA's MMAs overlap the other's epilogue):
// warps 0-3 epilogue (thread = column = TMEM lane), warp 4 MMA issuer, warp 5 bulk-copy producer.
#include <cstdlib>

#include "fmcw_internal.cuh"

namespace fmcw {

namespace {

constexpr float K_DB = 6.020599913279624f;
constexpr int TC_HALF = 10, TC_KP = 16;
constexpr int TC_M = 128, TC_N = 128;
constexpr int TC_MAT_BYTES = TC_N * TC_KP * 4;       // 8 KB per operand matrix
constexpr int TC_B_BYTES = 4 * TC_MAT_BYTES;         // Chi | Clo | Shi | Slo
constexpr int TC_THREADS = 192;
constexpr int TC_QF = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// canonical K-major, non-swizzled operand layout: (row/8)*SBO + (k/4)*128 + (row%8)*16 + (k%4)*4 bytes
__host__ __device__ __forceinline__ int tc_off_floats(int row, int k) {
  return (row >> 3) * (TC_KP * 8) + (k >> 2) * 32 + (row & 7) * 4 + (k & 3);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(128u >> 4) << 16;                 // leading byte offset: next 16-byte K chunk
  d |= (uint64_t)((TC_KP * 32u) >> 4) << 32;        // stride byte offset: next group of 8 rows
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell), no swizzle
  return d;
}
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// staged [32 columns][16 queries] tile of one warp -> 64-byte rows of the time-major spectrogram
__device__ __forceinline__ void flush_rows(uint32_t a_stage, int ncols_valid, float* __restrict__ out_warp,
                                           unsigned long long row_stride, int qbase, int nvalid, int lane) {
  __syncwarp();
  const int sub = lane >> 4, ql = lane & 15;
  float* ptr = out_warp + (unsigned long long)sub * row_stride + qbase + ql;
  uint32_t a = a_stage + (uint32_t)((sub * (TC_QF + 1) + ql) * 4);
  if (ql < nvalid) {
#pragma unroll 4
    for (int c = sub; c < ncols_valid; c += 2) {
      *ptr = lds32(a);
      ptr += 2 * row_stride;
      a += 2 * (TC_QF + 1) * 4;
    }
  }
  __syncwarp();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// B operands: per chunk of 128 bin positions the matrices Chi | Clo | Shi | Slo in the UMMA layout, and
// per position {K*log2(c_p), number of log-frequency queries that position completes}
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stft_tc_prepare_kernel(StftTables t, StftGeom g, float* __restrict__ tcB,
                                                              float2* __restrict__ tc_meta, int n_chunk_cap) {
  const StftPlan* P = t.plan;
  if (P->valid <= 0) return;
  const int nb = P->nb;
  const int n_chunks = (nb + TC_N - 1) / TC_N;
  if (n_chunks > n_chunk_cap) return;
  const unsigned long long nfft = P->nfft;
  const long long mod = (long long)(2 * nfft);
  const int total = n_chunks * TC_N * TC_KP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % TC_KP, pos = i / TC_KP;
    const int ch = pos / TC_N, n = pos % TC_N;
    double cv = 0.0, sv = 0.0;
    if (pos < nb) {
      const long long bin = t.bins[pos];
      if (k < TC_HALF) {
        const long long r = ((long long)(2 * k + 1) * bin) % mod;
        sincospi((double)r / (double)nfft, &sv, &cv);
      } else if (k == TC_HALF) {
        // window DC response at this bin: sum_m (w[9-m] + w[10+m]) cos((m+1/2) w)
        for (int m = 0; m < TC_HALF; ++m) {
          const long long r = ((long long)(2 * m + 1) * bin) % mod;
          cv += ((double)t.win[TC_HALF - 1 - m] + (double)t.win[TC_HALF + m]) * cospi((double)r / (double)nfft);
        }
      }
    }
    const float chi = tf32_rna((float)cv), shi = tf32_rna((float)sv);
    const float clo = tf32_rna((float)(cv - (double)chi)), slo = tf32_rna((float)(sv - (double)shi));
    float* blk = tcB + (size_t)ch * (TC_B_BYTES / 4);
    const int off = tc_off_floats(n, k);
    blk[off] = chi;
    blk[TC_MAT_BYTES / 4 + off] = clo;
    blk[2 * (TC_MAT_BYTES / 4) + off] = shi;
    blk[3 * (TC_MAT_BYTES / 4) + off] = slo;
    if (k == 0) {
      float kcb = 0.f;
      int cnt = 0;
      if (pos < nb) {
        kcb = t.kcb[pos];
        if (pos > 0) cnt = t.qend[pos] - t.qend[pos - 1];
      }
      tc_meta[pos] = make_float2(kcb, __int_as_float(cnt));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(TC_THREADS, 2)
stft_tc_kernel(StftTables t, StftGeom g, const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ tcB,
               const float2* __restrict__ tc_meta, unsigned long long capacity_cols, unsigned long long ld_cols, int* d_err) {
  StftPlan* P = t.plan;
  if (P->valid <= 0) { if (threadIdx.x == 0 && blockIdx.x == 0 && P->valid < 0) *d_err = P->valid; return; }
  extern __shared__ __align__(128) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);                                   // Ehi | Elo | Ohi | Olo
  float* sB = reinterpret_cast<float*>(smem + 4 * TC_MAT_BYTES);                // Chi | Clo | Shi | Slo
  float* s_aq = reinterpret_cast<float*>(smem + 4 * TC_MAT_BYTES + TC_B_BYTES); // [MAX_NQ]
  float* s_ws = s_aq + MAX_NQ;                                                  // [32]
  float* s_stage = s_ws + 32;                                                   // [4 warps][32][17]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + 4 * 32 * (TC_QF + 1));
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 8);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long cb = P->col_begin, ce = P->col_end, off = P->sample_offset;
  const unsigned long long ncl = ce - cb;
  if (ncl > capacity_cols) { if (tid == 0 && blockIdx.x == 0) *d_err = -4; return; }
  const int nq = P->nq, nb = P->nb;
  const int n_chunks = (nb + TC_N - 1) / TC_N;
  const unsigned long long n_tiles = (ncl + TC_M - 1) / TC_M;

  const uint32_t bar_a = smem_u32(&s_bar[0]), bar_bf = smem_u32(&s_bar[1]), bar_be = smem_u32(&s_bar[2]);
  const uint32_t bar_tf = smem_u32(&s_bar[3]), bar_te = smem_u32(&s_bar[4]);
  if (tid == 0) {
    mbar_init(bar_a, 128); mbar_init(bar_bf, 1); mbar_init(bar_be, 1); mbar_init(bar_tf, 1); mbar_init(bar_te, 128);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < nq; i += TC_THREADS) s_aq[i] = t.aq[i];
  if (tid < 2 * TC_HALF) s_ws[tid] = (float)((double)t.win[tid] / sqrt(P->pmax_raw));
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *s_tmem;

  if (warp < 4) {
    // ===================== epilogue warps: thread = spectrogram column = TMEM lane =====================
    const uint32_t a_stage = smem_u32(s_stage) + (uint32_t)(warp * 32 * (TC_QF + 1) * 4);
    const uint32_t a_st_lane = a_stage + (uint32_t)(lane * (TC_QF + 1) * 4);
    const uint32_t a_aq = smem_u32(s_aq);
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float inv = (float)(1.0 / sqrt(P->pmax_raw));
    uint32_t ph_tf = 0;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const unsigned long long tile_col0 = cb + tile * TC_M;
      unsigned long long col = tile_col0 + tid;
      if (col >= ce) col = ce - 1;
      // ---- A operands of this column: mean removed, windowed, folded, split hi/lo ----
      {
        const float* xs = x + (col * g.hop - off);
        float xv[2 * TC_HALF];
        float mean = 0.f;
#pragma unroll
        for (int n = 0; n < 2 * TC_HALF; ++n) { xv[n] = __ldg(xs + n); mean += xv[n]; }
        mean *= (1.0f / (2 * TC_HALF));
        float ev[TC_KP], ov[TC_KP];
#pragma unroll
        for (int m = 0; m < TC_HALF; ++m) {
          const float ylo = s_ws[TC_HALF - 1 - m] * (xv[TC_HALF - 1 - m] - mean), yhi = s_ws[TC_HALF + m] * (xv[TC_HALF + m] - mean);
          ev[m] = ylo + yhi;
          ov[m] = ylo - yhi;
        }
        ev[TC_HALF] = mean * inv;      // DC tap: multiplies the tabulated window response
        ov[TC_HALF] = 0.f;
#pragma unroll
        for (int m = TC_HALF + 1; m < TC_KP; ++m) { ev[m] = 0.f; ov[m] = 0.f; }
        float* rowp = sA + (tid >> 3) * (TC_KP * 8) + (tid & 7) * 4;
#pragma unroll
        for (int kc = 0; kc < TC_KP / 4; ++kc) {
          float4 eh, el, oh, ol;
          float* ehp = &eh.x; float* elp = &el.x; float* ohp = &oh.x; float* olp = &ol.x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float e = ev[4 * kc + j], o = ov[4 * kc + j];
            ehp[j] = tf32_rna(e); elp[j] = tf32_rna(e - ehp[j]);
            ohp[j] = tf32_rna(o); olp[j] = tf32_rna(o - ohp[j]);
          }
          *reinterpret_cast<float4*>(rowp + kc * 32) = eh;
          *reinterpret_cast<float4*>(rowp + (TC_MAT_BYTES / 4) + kc * 32) = el;
          *reinterpret_cast<float4*>(rowp + 2 * (TC_MAT_BYTES / 4) + kc * 32) = oh;
          *reinterpret_cast<float4*>(rowp + 3 * (TC_MAT_BYTES / 4) + kc * 32) = ol;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor-core (async proxy) reads
      mbar_arrive(bar_a);

      const unsigned long long warp_col0 = tile_col0 + (unsigned long long)warp * 32;
      const int ncols_valid = (warp_col0 >= ce) ? 0 : (int)((ce - warp_col0) < 32ull ? (ce - warp_col0) : 32ull);
      float* out_warp = out + (warp_col0 - cb) * (unsigned long long)nq;
      const bool col_ok = (tile_col0 + tid) < ce;
      float prev = 0.f;
      int qcur = 0;
      for (int ch = 0; ch < n_chunks; ++ch) {
        mbar_wait(bar_tf, ph_tf);
        ph_tf ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float2* meta = tc_meta + ch * TC_N;
#pragma unroll 1
        for (int g16 = 0; g16 < TC_N / 16; ++g16) {
          float re[16], im[16];
          tmem_ld16(t_lane + (uint32_t)(g16 * 16), re);
          tmem_ld16(t_lane + (uint32_t)(TC_N + g16 * 16), im);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (g16 == TC_N / 16 - 1) {           // accumulators are in registers: hand TMEM back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_te);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 mt = __ldg(meta + g16 * 16 + j);
            const float db = fmaf(K_DB, lg2_approx(fmaf(re[j], re[j], im[j] * im[j])), mt.x);
            const int cnt = __float_as_int(mt.y);
            for (int k = 0; k < cnt; ++k, ++qcur) {
              const float a = lds32(a_aq + 4 * qcur);
              const float val = fmaf(a, db - prev, prev);
              if (LAYOUT == 0) {
                const int slot = qcur & (TC_QF - 1);
                sts32(a_st_lane + (uint32_t)(slot * 4), val);
                if (slot == TC_QF - 1) flush_rows(a_stage, ncols_valid, out_warp, (unsigned long long)nq, qcur - (TC_QF - 1), TC_QF, lane);
              } else {
                if (col_ok) out[(unsigned long long)qcur * ld_cols + (tile_col0 + tid - cb)] = val;
              }
            }
            prev = db;
          }
        }
      }
      if (LAYOUT == 0) {
        const int rem = qcur & (TC_QF - 1);
        if (rem) flush_rows(a_stage, ncols_valid, out_warp, (unsigned long long)nq, qcur - rem, rem, lane);
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer: one thread drives the tensor core =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
      uint32_t ph_a = 0, ph_bf = 0, ph_te = 0;
      for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(bar_a, ph_a);
        ph_a ^= 1;
        for (int ch = 0; ch < n_chunks; ++ch) {
          mbar_wait(bar_bf, ph_bf);
          ph_bf ^= 1;
          mbar_wait(bar_te, ph_te ^ 1);      // TMEM free (passes immediately the first time)
          ph_te ^= 1;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int part = 0; part < 2; ++part) {          // 0: Re = E * C^T, 1: Im = O * S^T
            const uint32_t d = tmem_base + (uint32_t)(part * TC_N);
            const uint32_t a_hi = aA + (uint32_t)((2 * part) * TC_MAT_BYTES), a_lo = a_hi + TC_MAT_BYTES;
            const uint32_t b_hi = aB + (uint32_t)((2 * part) * TC_MAT_BYTES), b_lo = b_hi + TC_MAT_BYTES;
#pragma unroll
            for (int ks = 0; ks < TC_KP / 8; ++ks) {
              const uint32_t o = (uint32_t)(ks * 256);
              umma_tf32(d, make_desc(a_hi + o), make_desc(b_hi + o), idesc, ks > 0);
              umma_tf32(d, make_desc(a_hi + o), make_desc(b_lo + o), idesc, 1u);
              umma_tf32(d, make_desc(a_lo + o), make_desc(b_hi + o), idesc, 1u);
            }
          }
          umma_commit(bar_be);     // B tile consumed
          umma_commit(bar_tf);     // accumulators ready
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== producer: B tiles by 1-D bulk copy =====================
    if (lane == 0) {
      uint32_t ph_be = 0;
      const uint32_t aB = smem_u32(sB);
      for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int ch = 0; ch < n_chunks; ++ch) {
          mbar_wait(bar_be, ph_be ^ 1);
          ph_be ^= 1;
          mbar_expect_tx(bar_bf, TC_B_BYTES);
          bulk_g2s(aB, tcB + (size_t)ch * (TC_B_BYTES / 4), TC_B_BYTES, bar_bf);
        }
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
}

size_t stft_tc_table_bytes(int nb_max) {
  const int n_chunks = (nb_max + TC_N - 1) / TC_N;
  return (size_t)n_chunks * TC_B_BYTES;
}
size_t stft_tc_meta_bytes(int nb_max) {
  const int n_chunks = (nb_max + TC_N - 1) / TC_N;
  return (size_t)n_chunks * TC_N * sizeof(float2);
}

cudaError_t launch_stft_tc_prepare(const StftTables& t, const StftGeom& g, float* tcB, float2* tc_meta, int nb_max,
                                   cudaStream_t st) {
  const int n_chunk_cap = (nb_max + TC_N - 1) / TC_N;
  stft_tc_prepare_kernel<<<64, 256, 0, st>>>(t, g, tcB, tc_meta, n_chunk_cap);
  return cudaGetLastError();
}

cudaError_t launch_stft_tc_main(const StftTables& t, const StftGeom& g, const float* x, float* out, const float* tcB,
                                const float2* tc_meta, unsigned long long capacity_cols, unsigned long long ld_cols,
                                int layout, int* d_err, cudaStream_t st) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const size_t smem = 4 * TC_MAT_BYTES + TC_B_BYTES + (MAX_NQ + 32 + 4 * 32 * (TC_QF + 1)) * sizeof(float) + 8 * 8 + 16;
  cudaError_t e;
  if (layout == 0) {
    e = cudaFuncSetAttribute(stft_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    stft_tc_kernel<0><<<sms * 2, TC_THREADS, smem, st>>>(t, g, x, out, tcB, tc_meta, capacity_cols, ld_cols, d_err);
  } else {
    e = cudaFuncSetAttribute(stft_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    stft_tc_kernel<1><<<sms * 2, TC_THREADS, smem, st>>>(t, g, x, out, tcB, tc_meta, capacity_cols, ld_cols, d_err);
  }
  return cudaGetLastError();
}

}  // namespace fmcw
