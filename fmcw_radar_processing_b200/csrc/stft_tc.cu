// STFT main kernel on the 5th-generation tensor cores (tcgen05 + TMEM), window_length = 20.
//
// The 20-tap windowed DTFT at the two fine-grid bins bracketing every log-frequency query (see stft.cu) is the
// contraction
//     Re[col][bin] = sum_k E[col][k] * C[bin][k],   Im[col][bin] = sum_k O[col][k] * S[bin][k]
// with E/O the even/odd folded, windowed taps of a spectrogram column.  On the fp32 pipe this costs 20 FMAs
// per 4-byte output and bounds the kernel at ~50 % FMA-pipe utilisation (profiles/ncu_full_r1b.txt).  Here:
//   * one CTA tile = 128 spectrogram columns (the M = 128 rows of the UMMA tile, one TMEM lane each);
//   * a chunk = 64 consecutive queries = 128 bins laid out as (lower, upper) bracket pairs (N = 128); Re and Im
//     accumulators = 256 TMEM columns.  The layout is static, so the epilogue needs no index tables: a thread
//     reads 32 accumulator columns = 16 queries, and the interp1 is register arithmetic;
//   * operands are split hi + lo in TF32 (cvt.rna) and the three products hi*hi + hi*lo + lo*hi are laid side
//     by side along K (10 taps x 3 + 2 DC slots = 32 = four K = 8 instructions per part, 8 tcgen05.mma per
//     chunk), accumulated in fp32 -- ~2^-22 relative, the float32 class (single-pass TF32/BF16 would not meet
//     the 1e-3 dB tolerance);
//   * a TF32-exact estimate of the column mean is removed before the split and returns as the DC slots
//     (mean * window DC response W(w_bin), tabulated in float64 and split hi + lo), so the large DC term is
//     never rounded together with the small residual;
//   * the one-sided doubling (RP:276) is folded into the B operands (sqrt(2) on every bin but DC and Nyquist);
//   * A tiles are written to shared memory by a dedicated builder warp, one tile ahead (K-major, no swizzle, core
//     matrices of 8 rows x 16 bytes); B tiles (planned once on the device) arrive by 1-D bulk copy (cp.async.bulk) on an
//     mbarrier; tcgen05.commit signals the epilogue, which reads the accumulators with tcgen05.ld, takes
//     |S|^2 -> lg2, interpolates, scales to dB and stages the spectrogram rows in shared memory.
// One CTA per SM (736 threads): warps 0-15 epilogue (lane quarter = warp & 3; the four warps of a quarter each take 16 queries
// of every chunk), warp 16 MMA issuer, warp 17 bulk-copy producer, warp 18 A-tile builder, warps 19-22 TMA-store issuers (one
// per quarter: 32 columns x 64 queries leave as one 3-D cp.async.bulk.tensor store of two 128-byte-swizzled boxes).  TMEM (512
// columns), the A tiles, the B tiles and the output staging boxes are double buffered: the MMAs of chunk c+1 and the stores of
// chunk c-1 overlap the lg2 phase of chunk c.  Without a tensor map (run-time query count, frequency-major layout,
// FMCW_TC_TMA=0) the epilogue warps flush the staged rows themselves (64-bit shared-memory loads, coalesced global stores).
// What bounds the kernel, measured: profiles/stft_tc_bounds_r2.txt (FMCW_TC_DEBUG / FMCW_TC_PROF below are its instruments).
#include <cstdio>
#include <cstdlib>

#include <cuda.h>

#include "fmcw_internal.cuh"

namespace fmcw {

namespace {

constexpr float K_DB = 6.020599913279624f;            // 20 log10(2): dB per octave of the (power) ratio, RP:283
constexpr int TC_HALF = 10;                            // folded taps
constexpr int TC_M = 128, TC_N = 128, TC_K = 32;       // UMMA tile; K = 3 x 10 taps + 2 DC slots
constexpr int TC_QC = TC_N / 2;                        // queries per chunk
constexpr int TC_EW = 16;                              // epilogue warps
constexpr int TC_THREADS = (TC_EW + 7) * 32;          // + MMA issuer, bulk-copy producer, A-tile builder, four TMA-store issuers
constexpr int TC_A_MAT_BYTES = TC_M * TC_K * 4;        // 16 KB per A operand matrix (E or O)
constexpr int TC_A_BYTES = 2 * TC_A_MAT_BYTES;
constexpr int TC_B_MAT_BYTES = TC_N * TC_K * 4;        // 16 KB per B operand matrix (C or S)
constexpr int TC_B_BYTES = 2 * TC_B_MAT_BYTES;
constexpr int TC_TMEM_COLS = 4 * TC_N;                 // two stages of (Re | Im)
// output staging: [buffer][quarter][column][64 queries], row stride = 4 (mod 32) words: conflict-free 128-bit
// stores by column and conflict-free 64-bit loads by query
constexpr int TC_SROW = TC_QC + 4;
constexpr int TC_SQ = 32 * TC_SROW;                    // floats per quarter
constexpr int TC_SBUF = 4 * TC_SQ;                     // floats per buffer
constexpr int TC_OFF_B = 2 * TC_A_BYTES;
constexpr int TC_OFF_STG = TC_OFF_B + 2 * TC_B_BYTES;
constexpr int TC_OFF_AQ = TC_OFF_STG + 2 * TC_SBUF * 4;
constexpr int TC_OFF_BAR = TC_OFF_AQ + (MAX_NQ + 32) * 4;
constexpr int TC_SMEM = TC_OFF_BAR + 32 * 8 + 16;
// TMA-store epilogue: per buffer and quarter two boxes of [32 columns][32 queries] floats in the 128-byte swizzle
constexpr int TC_TBOX = 32 * 32 * 4;                   // 4 KB
constexpr int TC_TQ = 2 * TC_TBOX;                     // per quarter
constexpr int TC_TBUF = 4 * TC_TQ;                     // per buffer (32 KB; fits the 34.8 KB of the padded rows)
static_assert(TC_OFF_STG % 1024 == 0 && (TC_SBUF * 4) >= TC_TBUF, "swizzled boxes need 1024-byte alignment");
static_assert(MAX_NQ % TC_QC == 0, "query table is padded to whole chunks");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// canonical K-major, non-swizzled operand layout: (row/8)*SBO + (k/4)*128 + (row%8)*16 + (k%4)*4 bytes
__host__ __device__ __forceinline__ int tc_off_floats(int row, int k) {
  return (row >> 3) * (TC_K * 8) + (k >> 2) * 32 + (row & 7) * 4 + (k & 3);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(128u >> 4) << 16;                 // leading byte offset: next 16-byte K chunk
  d |= (uint64_t)((TC_K * 32u) >> 4) << 32;        // stride byte offset: next group of 8 rows
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell), no swizzle
  return d;
}
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(parity), "r"(1000000u) : "memory");
  }
}
// both barriers' try_waits are in flight together (one trip through the MIO queue instead of two)
__device__ __forceinline__ void mbar_wait2(uint32_t a, uint32_t pa, uint32_t b, uint32_t pb) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p, q;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %5;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4, %5;\nand.pred p, p, q;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(pa), "r"(b), "r"(pb), "r"(1000000u) : "memory");
  }
}
__device__ __forceinline__ bool mbar_test(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  return ok != 0;
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&a)[16], float (&b)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]),
                 "=f"(a[8]), "=f"(a[9]), "=f"(a[10]), "=f"(a[11]), "=f"(a[12]), "=f"(a[13]), "=f"(a[14]), "=f"(a[15]),
                 "=f"(b[0]), "=f"(b[1]), "=f"(b[2]), "=f"(b[3]), "=f"(b[4]), "=f"(b[5]), "=f"(b[6]), "=f"(b[7]),
                 "=f"(b[8]), "=f"(b[9]), "=f"(b[10]), "=f"(b[11]), "=f"(b[12]), "=f"(b[13]), "=f"(b[14]), "=f"(b[15])
               : "r"(taddr) : "memory");
}
// wait for this thread's outstanding tcgen05.ld and tie the destination registers to the wait, so that no use of
// them can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld(float (&a)[16], float (&b)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  asm volatile("" : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(a[4]), "+f"(a[5]), "+f"(a[6]), "+f"(a[7]),
                    "+f"(a[8]), "+f"(a[9]), "+f"(a[10]), "+f"(a[11]), "+f"(a[12]), "+f"(a[13]), "+f"(a[14]), "+f"(a[15]));
  asm volatile("" : "+f"(b[0]), "+f"(b[1]), "+f"(b[2]), "+f"(b[3]), "+f"(b[4]), "+f"(b[5]), "+f"(b[6]), "+f"(b[7]),
                    "+f"(b[8]), "+f"(b[9]), "+f"(b[10]), "+f"(b[11]), "+f"(b[12]), "+f"(b[13]), "+f"(b[14]), "+f"(b[15]));
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return r;
}

// volatile: keeps its place among the (volatile) shared-memory and global accesses of the epilogue loop, so the
// MUFU work stays spread over the loop instead of being bunched by the scheduler
__device__ __forceinline__ float lg2_pinned(float x) {
  float y;
  asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void stg64_if(float* p, float2 v, bool ok) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %3, 0;\n@q st.global.v2.f32 [%0], {%1, %2};\n}"
               ::"l"(p), "f"(v.x), "f"(v.y), "r"((int)ok) : "memory");
}
__device__ __forceinline__ void stg32_if(float* p, float v, bool ok) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q st.global.f32 [%0], %1;\n}" ::"l"(p), "f"(v), "r"((int)ok) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tmap), "r"(smem), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return r;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// B operands.  Chunk c, row n: query 64c + n/2, lower (n even) or upper (n odd) bracket bin.  Along K:
//   C: [Chi(10) | Clo(10) | Chi(10) | Whi | Wlo]     against A = [Ehi | Ehi | Elo | M | M]
//   S: [Shi(10) | Slo(10) | Shi(10) |  0  |  0 ]     against A = [Ohi | Ohi | Olo | 0 | 0]
// with cos/sin((k+1/2) w_bin) and the window DC response W(w_bin) in float64, times sqrt(2) for every bin that
// the one-sided spectrum doubles.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stft_tc_prepare_kernel(StftTables t, StftGeom g, float* __restrict__ tcB,
                                                              int spec_mode) {
  const StftPlan* P = t.plan;
  if (P->valid <= 0 || (spec_mode == 2 && P->spec_state == 2)) return;   // tables of the look-ahead plan stand
  const int nq = P->nq;
  const int n_chunks = (nq + TC_QC - 1) / TC_QC;
  const unsigned long long nfft = P->nfft;
  const long long mod = (long long)(2 * nfft);
  constexpr int NT = TC_HALF + 1;
  constexpr int mat = TC_N * TC_K;               // floats per operand matrix
  const int total = n_chunks * TC_N * NT;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % NT, idx = i / NT;
    const int ch = idx / TC_N, n = idx % TC_N;
    const int q = ch * TC_QC + (n >> 1);
    double cv = 0.0, sv = 0.0;
    if (q < nq) {
      const long long bin = (long long)t.bins[t.qpos[q]] + (n & 1);
      const double sc = (bin == 0 || (unsigned long long)bin == nfft / 2) ? 1.0 : 1.4142135623730951;
      if (k < TC_HALF) {
        const long long r = ((long long)(2 * k + 1) * bin) % mod;
        sincospi((double)r / (double)nfft, &sv, &cv);
        sv *= sc;
      } else {
        // window DC response at this bin: sum_m (w[9-m] + w[10+m]) cos((m+1/2) w)
        for (int m = 0; m < TC_HALF; ++m) {
          const long long r = ((long long)(2 * m + 1) * bin) % mod;
          cv += (t.win_d[TC_HALF - 1 - m] + t.win_d[TC_HALF + m]) * cospi((double)r / (double)nfft);
        }
      }
      cv *= sc;
    }
    const float chi = tf32_rna((float)cv), shi = tf32_rna((float)sv);
    const float clo = tf32_rna((float)(cv - (double)chi)), slo = tf32_rna((float)(sv - (double)shi));
    float* C = tcB + (size_t)ch * (2 * mat);
    float* S = C + mat;
    if (k < TC_HALF) {
      C[tc_off_floats(n, k)] = chi; C[tc_off_floats(n, TC_HALF + k)] = clo; C[tc_off_floats(n, 2 * TC_HALF + k)] = chi;
      S[tc_off_floats(n, k)] = shi; S[tc_off_floats(n, TC_HALF + k)] = slo; S[tc_off_floats(n, 2 * TC_HALF + k)] = shi;
    } else {
      C[tc_off_floats(n, 3 * TC_HALF)] = chi; C[tc_off_floats(n, 3 * TC_HALF + 1)] = clo;
      S[tc_off_floats(n, 3 * TC_HALF)] = 0.f; S[tc_off_floats(n, 3 * TC_HALF + 1)] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// main kernel: 16 epilogue warps + MMA issuer + bulk-copy producer
//   LAYOUT 0: time-major out[col][query]; LAYOUT 1: frequency-major out[query][ld_cols].
//   NQC: compile-time number of queries (1024) for the time-major fast path, 0 = run-time nq.
// ------------------------------------------------------------------------------------------------
template <int LAYOUT, int NQC, bool TMA, bool PROF = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
stft_tc_kernel(StftTables t, StftGeom g, const sig_t* __restrict__ x, float* __restrict__ out, const float* __restrict__ tcB,
               unsigned long long capacity_cols, unsigned long long ld_cols, int* d_err, int dbg_mode,
               const double* __restrict__ gmax_dev, const __grid_constant__ CUtensorMap tmap, int tma3d) {
  StftPlan* P = t.plan;
  if (P->valid <= 0) { if (threadIdx.x == 0 && blockIdx.x == 0 && P->valid < 0) *d_err = P->valid; return; }
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);                          // 2 x (E | O)
  float* sB = reinterpret_cast<float*>(smem + TC_OFF_B);               // 2 x (C | S)
  float* s_stg = reinterpret_cast<float*>(smem + TC_OFF_STG);          // 2 x [4 quarters][32 columns][TC_SROW]
  float* s_aq = reinterpret_cast<float*>(smem + TC_OFF_AQ);            // [MAX_NQ] interp1 weights (zero padded)
  float* s_ws = s_aq + MAX_NQ;                                         // [32] normalised window
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 32);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long cb = P->col_begin, ce = P->col_end, off = P->sample_offset;
  const unsigned long long ncl = ce - cb;
  if (ncl > capacity_cols) { if (tid == 0 && blockIdx.x == 0) *d_err = -4; return; }
  const int nq = (NQC > 0) ? NQC : P->nq;
  const int n_chunks = (nq + TC_QC - 1) / TC_QC;
  const unsigned long long n_tiles = (ncl + TC_M - 1) / TC_M;
  // normalisation max(P): the plan's own search, or the all-reduced value of a sharded run
  const double pmax = gmax_dev ? *gmax_dev : P->pmax_raw;
  if (gmax_dev && blockIdx.x == 0 && threadIdx.x == 0) P->pmax_raw = pmax;
  const double inv_d = 1.0 / sqrt(pmax);

  // barriers: [0,1] a_full, [2,3] a_empty, [4,5] b_full, [6,7] b_empty, [8,9] t_full, [10,11] t_empty,
  // [12 + 2*quarter + buffer] staging rows of a quarter complete (one arrival per warp)
  const uint32_t bar0 = smem_u32(&s_bar[0]);
  auto BAR = [&](int i) { return bar0 + (uint32_t)(i * 8); };
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(0 + i), 1); mbar_init(BAR(2 + i), 1);
      mbar_init(BAR(4 + i), 1); mbar_init(BAR(6 + i), 1);
      mbar_init(BAR(8 + i), 1); mbar_init(BAR(10 + i), TC_EW);
    }
    for (int i = 0; i < 8; ++i) mbar_init(BAR(12 + i), 4);
    for (int i = 0; i < 8; ++i) mbar_init(BAR(20 + i), 1);      // [20 + 2*quarter + buffer] TMA store of the buffer has read it
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == TC_EW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < MAX_NQ; i += TC_THREADS) s_aq[i] = (i < nq) ? t.aq[i] : 0.f;
  if (tid < 2 * TC_HALF) s_ws[tid] = (float)((double)t.win[tid] * inv_d);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *s_tmem;

  if (warp < TC_EW) {
    // ====== epilogue warps: lane quarter qd = warp & 3 (TMEM lanes 32*qd..), sub-warp sw = warp >> 2 ======
    const int qd = warp & 3, sw = warp >> 2;
    const int m = qd * 32 + lane;                         // row of the tile = spectrogram column = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(sw * 32);
    const uint32_t a_aq = smem_u32(s_aq) + (uint32_t)(sw * 16 * 4);            // weights of this warp's 16 queries
    const uint32_t a_stq = smem_u32(s_stg) + (uint32_t)(qd * TC_SQ * 4);       // this quarter's staging rows
    const uint32_t a_st_w = a_stq + (uint32_t)((lane * TC_SROW + sw * 16) * 4);        // store: own column, own 16 queries
    const uint32_t a_st_r = a_stq + (uint32_t)((sw * 8 * TC_SROW + 2 * lane) * 4);     // load: 8 columns, 2 queries per lane

    // the finished chunk still in the staging rows: its 8 columns x 64 queries leave as 256-byte rows while the
    // next chunk's lg2 phase runs
    float* pend_ptr = out;          // out + this warp's first column row + chunk's first query + 2*lane
    int pend_cols = 0, pend_q = 0;  // valid columns among the warp's 8 (0: nothing staged yet); first query of this lane's pair
    uint32_t pend_addr = a_st_r, pend_bar = 0, pend_par = 0;
    bool pending = false;
    float2 pend_v = make_float2(0.f, 0.f);
    auto flush_begin = [&]() {      // the other three warps of the quarter have staged their queries; first row in flight
      if (LAYOUT != 0 || TMA) return;
      if (pending) mbar_wait(pend_bar, pend_par);
      pend_v = lds64(pend_addr);
    };
    auto flush_col = [&](int j) {   // store row j (loaded one step earlier), load row j + 1
      if (LAYOUT != 0 || TMA) return;
      const bool ok = j < pend_cols;
      if (NQC > 0) {
        stg64_if(pend_ptr + (size_t)j * NQC, pend_v, ok);
      } else {
        float* p = pend_ptr + (size_t)j * nq;
        stg32_if(p, pend_v.x, ok && pend_q < nq);
        stg32_if(p + 1, pend_v.y, ok && pend_q + 1 < nq);
      }
      if (j < 7) pend_v = lds64(pend_addr + (uint32_t)((j + 1) * TC_SROW * 4));
    };

    // |S|^2 -> lg2 of the (lower, upper) pairs of 8 queries, interp1 in registers, dB; every pair of queries also
    // moves one staged row of the previous chunk to global memory (rows jf .. jf+3)
    const float2 kk = make_float2(K_DB, K_DB);
    auto half_chunk = [&](const float (&re)[16], const float (&im)[16], uint32_t a_w, float* o, int jf) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 ra = make_float2(re[4 * j], re[4 * j + 1]), ia = make_float2(im[4 * j], im[4 * j + 1]);
        const float2 rb = make_float2(re[4 * j + 2], re[4 * j + 3]), ib = make_float2(im[4 * j + 2], im[4 * j + 3]);
        const float2 pa = fma2(ra, ra, mul2(ia, ia)), pb = fma2(rb, rb, mul2(ib, ib));   // (lower, upper) of queries 2j, 2j+1
        flush_col(jf + j);
        const float2 l_lo = make_float2(lg2_pinned(pa.x), lg2_pinned(pb.x));
        const float2 l_hi = make_float2(lg2_pinned(pa.y), lg2_pinned(pb.y));
        const float2 a2 = lds64(a_w + (uint32_t)(j * 8));
        const float2 o2 = mul2(kk, fma2(a2, sub2(l_hi, l_lo), l_lo));
        o[2 * j] = o2.x; o[2 * j + 1] = o2.y;
      }
    };
    const bool skip_math = (dbg_mode & 1);      // FMCW_TC_DEBUG bits (timing experiments, results are wrong): 1 no lg2 phase / stores,
                                                // 2 A tiles built twice only, 8 no TMA stores
    float reA[16], imA[16], reB[16], imB[16];              // accumulator halves

    uint32_t pf[7] = {0, 0, 0, 0, 0, 0, 0}, pt = 0;       // PROF: cycles per phase of this warp (32-bit: short runs only)
    auto tick = [&](int i) { if (PROF) { const uint32_t n = (uint32_t)clock(); pf[i] += n - pt; pt = n; } };
    if (PROF) pt = (uint32_t)clock();
    unsigned long long it = 0;                            // local tile counter
    if constexpr (TMA) {
      // ---- fast path (time-major, 1,024 queries): the staged boxes leave through the TMA-store warp.  Per chunk a warp makes
      // three trips through the MIO queue it shares with the MUFUs of its neighbours (barrier pair, TMEM load, staging stores);
      // the chunk loop is unrolled by two so that the TMEM stage / staging buffer and every barrier address are immediates.
      static_assert(LAYOUT == 0 && NQC > 0 && (NQC / TC_QC) % 4 == 0, "parities below assume a multiple of four chunks per tile");
      constexpr int NCH = NQC / TC_QC;
      uint32_t a_box[4];                                  // this lane's four 16-byte pieces of its staged row (128-byte swizzle)
      {
        const uint32_t box0 = smem_u32(s_stg) + (uint32_t)(qd * TC_TQ + (sw >> 1) * TC_TBOX + lane * 128);
#pragma unroll
        for (int i = 0; i < 4; ++i) a_box[i] = box0 + (uint32_t)(((((sw & 1) * 4 + i) ^ (lane & 7))) << 4);
      }
      const uint32_t b_tfull = BAR(8), b_tempty = BAR(10), b_sfull = BAR(12 + 2 * qd), b_sfree = BAR(20 + 2 * qd);
      if (dbg_mode >> 8) __nanosleep((unsigned)(sw * (dbg_mode >> 8) * 16));      // experiment: start the four warps of a quarter out of phase
      for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        uint32_t a_w = a_aq;
#pragma unroll 1
        for (int ch2 = 0; ch2 < NCH; ch2 += 2) {
          const uint32_t par = (uint32_t)((ch2 >> 1) & 1);   // (chunk sequence number >> 1) & 1: NCH / 2 is even
#pragma unroll
          for (int ts = 0; ts < 2; ++ts) {
            tick(0);
            mbar_wait(b_tfull + 8u * ts, par);          // accumulators of this chunk ready
            tick(1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tmem_ld32(t_lane + (uint32_t)(ts * 2 * TC_N), reA, reB);
            tmem_ld32(t_lane + (uint32_t)(ts * 2 * TC_N + TC_N), imA, imB);
            tmem_wait_ld(reA, imA);
            tmem_wait_ld(reB, imB);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) mbar_arrive(b_tempty + 8u * ts);   // TMEM stage back to the MMA warp
            tick(2);
            if (skip_math) { if (reA[0] + imB[15] == 123.456f) out[0] = 1.f; continue; }
            float o[16];
            if ((dbg_mode & 64) && sw > ((dbg_mode >> 2) & 1)) {   // experiment: only one (64) or two (68) warps per quarter run the lg2 phase
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = reA[i] + imB[i];
            } else {
              half_chunk(reA, imA, a_w, o, 0);
              half_chunk(reB, imB, a_w + 32u, o + 8, 4);
            }
            a_w += (uint32_t)(TC_QC * 4);
            if (PROF) { asm volatile("" :: "f"(o[0]), "f"(o[3]), "f"(o[7]), "f"(o[11]), "f"(o[15])); }
            tick(4);
            if (it > 0 || ch2 > 0) mbar_wait(b_sfree + 8u * ts, par ^ 1u);   // staging buffer ts (last used two chunks ago) read by its store
            tick(5);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              sts128(a_box[i] + (uint32_t)(ts * TC_TBUF), make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
            __syncwarp();
            if (lane == 0) mbar_arrive(b_sfull + 8u * ts);    // (release) this warp's 16 queries of the quarter's box pair are staged
            tick(6);
          }
        }
      }
      if (PROF && lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
        printf("PROF cta %3d warp %2d (quarter %d sub %d) chunks %llu | cycles per chunk: loop %5u  t_full %5u  ld %5u  math %5u  free %5u  sts %5u\n",
               blockIdx.x, warp, qd, sw, it * NCH, pf[0] / (uint32_t)(it * NCH), pf[1] / (uint32_t)(it * NCH), pf[2] / (uint32_t)(it * NCH),
               pf[4] / (uint32_t)(it * NCH), pf[5] / (uint32_t)(it * NCH), pf[6] / (uint32_t)(it * NCH));
    } else {
      for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const unsigned long long tile_col0 = cb + tile * TC_M;
        const unsigned long long warp_col0 = tile_col0 + (unsigned long long)(qd * 32 + sw * 8);   // first of the warp's 8 store columns
        const int wcols = (warp_col0 >= ce) ? 0 : (int)((ce - warp_col0) < 8ull ? (ce - warp_col0) : 8ull);
        const bool col_ok = (tile_col0 + m) < ce;
        for (int ch = 0; ch < n_chunks; ++ch) {
          const unsigned long long cseq = it * (unsigned long long)n_chunks + ch;
          const int ts = (int)(cseq & 1);
          const uint32_t c0 = t_lane + (uint32_t)(ts * 2 * TC_N);
          mbar_wait(BAR(8 + ts), (uint32_t)((cseq >> 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          tmem_ld32(c0, reA, reB);
          tmem_ld32(c0 + TC_N, imA, imB);
          tmem_wait_ld(reA, imA);
          tmem_wait_ld(reB, imB);
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (lane == 0) mbar_arrive(BAR(10 + ts));         // (warp-collective loads are complete) TMEM stage back to the MMA warp
          if (skip_math) { if (reA[0] + imB[15] == 123.456f) out[0] = 1.f; continue; }
          const uint32_t a_w = a_aq + (uint32_t)(ch * TC_QC * 4);
          float o[16];
          flush_begin();
          half_chunk(reA, imA, a_w, o, 0);
          half_chunk(reB, imB, a_w + 32u, o + 8, 4);
          if (LAYOUT == 0) {
            const uint32_t aw = a_st_w + (uint32_t)(ts * TC_SBUF * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) sts128(aw + (uint32_t)(i * 16), make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(12 + 2 * qd + ts));  // this warp's 16 queries of the quarter's 32 x 64 outputs are staged
            pending = true;
            pend_bar = BAR(12 + 2 * qd + ts);
            pend_par = (uint32_t)((cseq >> 1) & 1);
            pend_addr = a_st_r + (uint32_t)(ts * TC_SBUF * 4);
            pend_q = ch * TC_QC + 2 * lane;
            pend_ptr = out + (warp_col0 - cb) * (unsigned long long)nq + pend_q;
            pend_cols = wcols;
          } else if (col_ok) {
            // frequency-major: lanes = consecutive columns, coalesced 128-byte rows straight from registers
            float* p = out + (unsigned long long)(ch * TC_QC + sw * 16) * ld_cols + (tile_col0 + m - cb);
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (NQC > 0 || ch * TC_QC + sw * 16 + i < nq) p[(unsigned long long)i * ld_cols] = o[i];
          }
        }
      }
      flush_begin();
#pragma unroll
      for (int j = 0; j < 8; ++j) flush_col(j);
    }
  } else if (warp == TC_EW) {
    // ===================== MMA issuer: one thread drives the tensor core =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      long long mw_b = 0, mw_t = 0, mw_m = 0;         // PROF: cycles waiting for B tiles / for the TMEM stage / issuing (or executing, bit 32)
      unsigned long long it = 0;
      for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int abuf = (int)(it & 1);
        mbar_wait(BAR(0 + abuf), (uint32_t)((it >> 1) & 1));
        const uint32_t aA = smem_u32(sA) + (uint32_t)(abuf * TC_A_BYTES);
        for (int ch = 0; ch < n_chunks; ++ch) {
          const unsigned long long cseq = it * (unsigned long long)n_chunks + ch;
          const int st = (int)(cseq & 1);
          const uint32_t par = (uint32_t)((cseq >> 1) & 1);
          long long q0 = 0, q1 = 0, q2 = 0;
          if (PROF) q0 = clock64();
          mbar_wait(BAR(4 + st), par);               // B tile landed
          if (PROF) q1 = clock64();
          mbar_wait(BAR(10 + st), par ^ 1);          // TMEM stage drained (passes immediately the first two times)
          if (PROF) { q2 = clock64(); mw_b += q1 - q0; mw_t += q2 - q1; }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t aB = smem_u32(sB) + (uint32_t)(st * TC_B_BYTES);
#pragma unroll
          for (int part = 0; part < 2; ++part) {          // 0: Re = E * C^T, 1: Im = O * S^T
            const uint32_t d = tmem_base + (uint32_t)(st * 2 * TC_N + part * TC_N);
            const uint32_t a = aA + (uint32_t)(part * TC_A_MAT_BYTES), b = aB + (uint32_t)(part * TC_B_MAT_BYTES);
#pragma unroll
            for (int ks = 0; ks < TC_K / 8; ++ks)
              umma_tf32(d, make_desc(a + (uint32_t)(ks * 256)), make_desc(b + (uint32_t)(ks * 256)), idesc, ks > 0);
          }
          umma_commit(BAR(6 + st));                    // B stage consumed
          umma_commit(BAR(8 + st));                    // accumulators ready
          if (PROF) {                                   // timing only: the issuer waits for its own MMAs (serialises the pipeline!)
            if (dbg_mode & 32) { mbar_wait(BAR(8 + st), par); mw_m += clock64() - q2; }
            else mw_m += clock64() - q2;
          }
        }
        umma_commit(BAR(2 + abuf));                    // A buffer consumed
      }
      if (PROF && (blockIdx.x == 0 || blockIdx.x == 77) && it > 0)
        printf("PROF cta %3d MMA issuer | cycles per chunk: wait B %5lld  wait TMEM stage %5lld  issue%s %5lld\n", blockIdx.x,
               mw_b / (long long)(it * n_chunks), mw_t / (long long)(it * n_chunks), (dbg_mode & 32) ? " + execute" : "", mw_m / (long long)(it * n_chunks));
    }
    __syncwarp();
  } else if (warp == TC_EW + 2) {
    // ===================== A-tile builder: one warp, four rows (spectrogram columns) per lane =====================
    // A operands of one column: a TF32-exact mean estimate removed in float64, windowed, folded even/odd, split
    // hi/lo and laid out along K as [hi | hi | lo | M M] (E) and [hi | hi | lo | 0 0] (O).  The warp runs one tile
    // ahead of the MMA issuer, so neither the tensor core nor the epilogue warps ever wait for it.
    unsigned long long it = 0;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      mbar_wait(BAR(2 + buf), (uint32_t)(((it >> 1) & 1) ^ 1));   // the MMAs of the tile that last used this buffer retired
#pragma unroll 1
      for (int r = 0; r < ((dbg_mode & 2) && it >= 2 ? 0 : TC_M / 32); ++r) {
        const int m = r * 32 + lane;                       // row of the tile = spectrogram column = TMEM lane
        unsigned long long col = cb + tile * TC_M + m;
        if (col >= ce) col = ce - 1;
        const sig_t* xs = x + (col * g.hop - off);
        double xd[2 * TC_HALF];
        double mean_d = 0.0;
#pragma unroll
        for (int n = 0; n < 2 * TC_HALF; ++n) { xd[n] = __ldg(xs + n); mean_d += xd[n]; }
        mean_d *= (1.0 / (2 * TC_HALF));
        const float M = tf32_rna((float)(mean_d * inv_d));   // DC slot value; mu = M / inv is what the residual removes
        const double mu = (double)M * sqrt(pmax);
        float yv[2 * TC_HALF];
#pragma unroll
        for (int n = 0; n < 2 * TC_HALF; ++n) yv[n] = s_ws[n] * (float)(xd[n] - mu);
#pragma unroll
        for (int part = 0; part < 2; ++part) {             // 0: even (E), 1: odd (O)
          float hi[TC_HALF], lo[TC_HALF];
#pragma unroll
          for (int k = 0; k < TC_HALF; ++k) {
            const float v = part == 0 ? (yv[TC_HALF - 1 - k] + yv[TC_HALF + k]) : (yv[TC_HALF - 1 - k] - yv[TC_HALF + k]);
            hi[k] = tf32_rna(v);
            lo[k] = tf32_rna(v - hi[k]);
          }
          const float dc = part == 0 ? M : 0.f;
          float* rowp = sA + buf * (TC_A_BYTES / 4) + part * (TC_A_MAT_BYTES / 4) + (m >> 3) * (TC_K * 8) + (m & 7) * 4;
          *reinterpret_cast<float4*>(rowp + 0 * 32) = make_float4(hi[0], hi[1], hi[2], hi[3]);     // slots 0..9: hi
          *reinterpret_cast<float4*>(rowp + 1 * 32) = make_float4(hi[4], hi[5], hi[6], hi[7]);
          *reinterpret_cast<float4*>(rowp + 2 * 32) = make_float4(hi[8], hi[9], hi[0], hi[1]);     // slots 10..19: hi again
          *reinterpret_cast<float4*>(rowp + 3 * 32) = make_float4(hi[2], hi[3], hi[4], hi[5]);
          *reinterpret_cast<float4*>(rowp + 4 * 32) = make_float4(hi[6], hi[7], hi[8], hi[9]);
          *reinterpret_cast<float4*>(rowp + 5 * 32) = make_float4(lo[0], lo[1], lo[2], lo[3]);     // slots 20..29: lo
          *reinterpret_cast<float4*>(rowp + 6 * 32) = make_float4(lo[4], lo[5], lo[6], lo[7]);
          *reinterpret_cast<float4*>(rowp + 7 * 32) = make_float4(lo[8], lo[9], dc, dc);           // slots 30, 31: DC
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor-core (async proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(0 + buf));
    }
  } else if (warp >= TC_EW + 3) {
    // ===================== TMA-store issuer (fast path only) =====================
    // Four warps (one thread each), one per quarter: wait until the quarter's four warps have staged their box pair (32 columns x 64
    // queries, two 128-byte-swizzled boxes, one 3-D tensor store), fence (the proxy fence sits here, after the acquire of the barrier
    // the writers released, so the epilogue warps neither fence nor wait for a store), store, and hand the buffer back (free barrier)
    // as soon as the store has read it.
    if constexpr (TMA) {
      if (lane == 0 && !(dbg_mode & 1)) {
        constexpr uint32_t NCH = NQC / TC_QC;
        const uint32_t n_local = (n_tiles > blockIdx.x) ? (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
        const int q = warp - (TC_EW + 3);
        const uint32_t b_sfull = BAR(12 + 2 * q), b_sfree = BAR(20 + 2 * q);
        const uint32_t src0 = smem_u32(s_stg) + (uint32_t)(q * TC_TQ);
        long long sp[4] = {0, 0, 0, 0}, s0 = 0;        // PROF: wait staged / fence / store + commit / wait_group.read
        const long long s_begin = PROF ? clock64() : 0;
        for (uint32_t itl = 0; itl < n_local; ++itl) {
          const int col = (int)((blockIdx.x + ((dbg_mode & 128) ? 0u : itl) * gridDim.x) * (uint32_t)TC_M) + q * 32;   // first output row of the quarter (experiment 128: every tile lands on the CTA's first, in L2)
#pragma unroll 1
          for (uint32_t ch = 0; ch < NCH; ++ch) {
            const uint32_t ts = ch & 1u;
            if (PROF) s0 = clock64();
            mbar_wait(b_sfull + 8u * ts, (ch >> 1) & 1u);
            if (PROF) { const long long n = clock64(); sp[0] += n - s0; s0 = n; }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> TMA (async proxy) reads
            if (PROF) { const long long n = clock64(); sp[1] += n - s0; s0 = n; }
            const uint32_t src = src0 + ts * (uint32_t)TC_TBUF;
            if (!(dbg_mode & 8)) {
              if (tma3d) {
                tma_store_3d(&tmap, src, 0, col, (int)(2u * ch));
              } else {
                tma_store_2d(&tmap, src, (int)(ch * TC_QC), col);
                tma_store_2d(&tmap, src + TC_TBOX, (int)(ch * TC_QC) + 32, col);
              }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              if (PROF) { const long long n = clock64(); sp[2] += n - s0; s0 = n; }
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              if (PROF) { const long long n = clock64(); sp[3] += n - s0; s0 = n; }
            }
            mbar_arrive(b_sfree + 8u * ts);
          }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (PROF && (blockIdx.x == 0 || blockIdx.x == 77) && n_local > 0) {
          const long long tot = (long long)n_local * NCH;
          printf("PROF cta %3d store issuer lane %d | cycles per chunk: wait staged %5lld  fence %5lld  store + commit %5lld  wait_group.read %5lld  all %5lld\n",
                 blockIdx.x, q, sp[0] / tot, sp[1] / tot, sp[2] / tot, sp[3] / tot, (clock64() - s_begin) / tot);
        }
      }
      __syncwarp();
    }
  } else {
    // ===================== producer: B tiles by 1-D bulk copy =====================
    if (lane == 0) {
      unsigned long long it = 0;
      for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        for (int ch = 0; ch < n_chunks; ++ch) {
          const unsigned long long cseq = it * (unsigned long long)n_chunks + ch;
          const int st = (int)(cseq & 1);
          mbar_wait(BAR(6 + st), (uint32_t)(((cseq >> 1) & 1) ^ 1));
          if ((dbg_mode & 16) && cseq >= 2) { mbar_arrive(BAR(4 + st)); continue; }     // timing experiment: B tiles never reloaded
          mbar_expect_tx(BAR(4 + st), TC_B_BYTES);
          bulk_g2s(smem_u32(sB) + (uint32_t)(st * TC_B_BYTES), tcB + (size_t)ch * (TC_B_BYTES / 4), TC_B_BYTES, BAR(4 + st));
        }
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TC_EW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS));
}

size_t stft_tc_table_bytes(int nb_max) {
  (void)nb_max;
  return (size_t)(MAX_NQ / TC_QC) * TC_B_BYTES;      // one (C | S) pair per chunk of 64 queries
}

cudaError_t launch_stft_tc_prepare(const StftTables& t, const StftGeom& g, float* tcB, int nb_max, cudaStream_t st,
                                   int spec_mode) {
  (void)nb_max;
  stft_tc_prepare_kernel<<<64, 256, 0, st>>>(t, g, tcB, spec_mode);
  return cudaGetLastError();
}

// Tensor map of the spectrogram [capacity_cols][1024] floats for the TMA-store epilogue: boxes of 32 queries x 32 columns, 128-byte
// swizzle.  cuTensorMapEncodeTiled comes through the runtime's driver entry point (no link against libcuda).
static bool make_out_tmap(CUtensorMap* map, float* out, unsigned long long capacity_cols, bool three_d) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static const encode_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    cudaGetLastError();
    return (encode_fn)p;
  }();
  if (!fn || capacity_cols == 0) return false;
  if (three_d) {
    // [query group of 32][column][32 queries]: one box = both 32-query halves of a chunk for 32 columns, laid out in shared memory
    // half-major -- exactly the two 2-D boxes back to back
    const cuuint64_t dims[3] = {32, capacity_cols, 32};
    const cuuint64_t strides[2] = {4096, 128};
    const cuuint32_t box[3] = {32, 32, 2}, estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  const cuuint64_t dims[2] = {1024, capacity_cols};
  const cuuint64_t strides[1] = {4096};
  const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int L, int Q, bool TMA = false, bool PROF = false>
static cudaError_t launch_tc(const StftTables& t, const StftGeom& g, const sig_t* x, float* out, const float* tcB,
                             unsigned long long capacity_cols, unsigned long long ld_cols, int* d_err, cudaStream_t st,
                             const double* gmax_dev, int sms, int dbg, const CUtensorMap* map = nullptr, int tma3d = 0) {
  cudaError_t e = cudaFuncSetAttribute(stft_tc_kernel<L, Q, TMA, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  if (e != cudaSuccess) return e;
  CUtensorMap dummy{};
  const CUtensorMap& tm = map ? *map : dummy;
  stft_tc_kernel<L, Q, TMA, PROF><<<sms, TC_THREADS, TC_SMEM, st>>>(t, g, x, out, tcB, capacity_cols, ld_cols, d_err, dbg, gmax_dev, tm, tma3d);
  return cudaGetLastError();
}

cudaError_t launch_stft_tc_main(const StftTables& t, const StftGeom& g, const sig_t* x, float* out, const float* tcB,
                                unsigned long long capacity_cols, unsigned long long ld_cols,
                                int layout, int* d_err, cudaStream_t st, const double* gmax_dev) {
  int sms = device_sm_count();
  static const int dbg = env_int("FMCW_TC_DEBUG", 0);
  // small recordings (a fleet of radars, each on its own stream): a CTA per FMCW_TC_TILES_PER_CTA tiles of the buffer's capacity
  // instead of one per SM, so that the kernels of several recordings run side by side
  static const int tiles_env = env_int("FMCW_TC_TILES_PER_CTA", 0);
  const int tiles_per_cta = tiles_env > 0 ? tiles_env : (int)g.tiles_per_cta;
  if (tiles_per_cta > 1) {
    const unsigned long long want = (capacity_cols / TC_M + (unsigned long long)tiles_per_cta) / (unsigned long long)tiles_per_cta;
    if (want < (unsigned long long)sms) sms = (int)(want < 1 ? 1 : want);
  }
#define FMCW_TC_ARGS t, g, x, out, tcB, capacity_cols, ld_cols, d_err, st, gmax_dev, sms, dbg
  if (layout != 0) return launch_tc<1, 0>(FMCW_TC_ARGS);
  // the fast path stores 64-bit pairs: 1,024 queries and an 8-byte aligned spectrogram
  static const int use_tma = env_int("FMCW_TC_TMA", 1);
  if (use_tma && g.nq == 1024 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && capacity_cols < (1ull << 31)) {   // (tensor coordinates are int32)
    // TMA-store epilogue: the staged rows leave through the async proxy (no LDS / STG on the LSU pipe).  Rows between the
    // column count and capacity_cols that share the last tile are overwritten (with copies of the last column).
    CUtensorMap map;
    static const int want3d = env_int("FMCW_TC_TMA3D", 1);
    int tma3d = want3d && make_out_tmap(&map, out, capacity_cols, true);
    if (tma3d || make_out_tmap(&map, out, capacity_cols, false)) {
      static const int prof = env_int("FMCW_TC_PROF", 0);       // per-warp phase timing printed by CTAs 0 and 77 (diagnostics)
      if (prof) {
        static bool told = false;
        if (!told) { told = true; fprintf(stderr, "stft_tc: TMA store epilogue, %s boxes\n", tma3d ? "3-D" : "2-D"); }
        return launch_tc<0, 1024, true, true>(FMCW_TC_ARGS, &map, tma3d);
      }
      return launch_tc<0, 1024, true>(FMCW_TC_ARGS, &map, tma3d);
    }
  }
  if (g.nq == 1024 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) return launch_tc<0, 1024>(FMCW_TC_ARGS);
  return launch_tc<0, 0>(FMCW_TC_ARGS);
#undef FMCW_TC_ARGS
}

}  // namespace fmcw
