// Per-frame chain of the reference (radar_processing.m RP:199-260), ONE WARP PER FRAME, two chirps per lane.
//
//   unpack int16 I/Q -> calibration + IF scale + fast-time mean removal (RP:203-204, on exact integers held in
//   float) -> 2*blackman window -> 256-point range FFT (RP:205) -> max |.| over chirps (RP:210) -> f_search_peak
//   (RP:211) -> slow-time row at the selected bin (RP:259, float64 single-bin DFT) -> mean removal over chirps,
//   2*chebwin window, ND-point Doppler FFT, fftshift (RP:217-219) -> first max + threshold (RP:233-238).
//
// Why this shape (profiles/ncu_full_r1e.txt: the one-CTA-per-frame kernel was instruction-issue bound, 26.9 k warp
// instructions per frame, 68 % issue-slot utilisation, 11 % DRAM):
//   * 256 = 16 x 16 as before, sixteen lanes per chirp, but a lane now carries the SAME butterfly of TWO chirps
//     (4q + half and 4q + 2 + half) as packed pairs (re_A, re_B), (im_A, im_B): every add, every twiddle multiply,
//     the window and the |X|^2 are FADD2 / FMUL2 / FFMA2 on the sm_100 f32x2 pipe -- half the floating-point
//     instructions per chirp -- and the multiplications by -j / +j cost nothing (they pick which pair is added);
//   * int16 -> float without I2F: PRMT drops the (sign-flipped) code into the mantissa of 2^23, one FADD2
//     removes the bias; NTS*code - sum stays an exact integer inside one FFMA2;
//   * a warp owns a frame from the first load to the Doppler bin: no __syncthreads, no cross-warp reduction, the
//     running max of |X|^2 over all PN chirps never leaves the registers; the four warps of a CTA only share the
//     read-only tables.  The peak search runs on registers + four shuffles, arg-max by redux.sync.
#include <cstdlib>

#include "fmcw_internal.cuh"

namespace fmcw {

namespace {

struct cx2 { float2 re, im; };     // one complex number of chirp A (.x) and of chirp B (.y)

// signed 16-bit halves of a packed (I, Q) word as float / double: one conversion instruction each (I2F.*.S16 R.H0 / R.H1)
__device__ __forceinline__ float cvt_lo(uint32_t w) {
  short lo, hi; float f;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(w));
  asm("cvt.rn.f32.s16 %0, %1;" : "=f"(f) : "h"(lo));
  return f;
}
__device__ __forceinline__ float cvt_hi(uint32_t w) {
  short lo, hi; float f;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(w));
  asm("cvt.rn.f32.s16 %0, %1;" : "=f"(f) : "h"(hi));
  return f;
}
__device__ __forceinline__ double cvtd_lo(uint32_t w) {
  short lo, hi; double f;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(w));
  asm("cvt.rn.f64.s16 %0, %1;" : "=d"(f) : "h"(lo));
  return f;
}
__device__ __forceinline__ double cvtd_hi(uint32_t w) {
  short lo, hi; double f;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(w));
  asm("cvt.rn.f64.s16 %0, %1;" : "=d"(f) : "h"(hi));
  return f;
}
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ cx2 cadd(const cx2& a, const cx2& b) { return cx2{add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ cx2 csub(const cx2& a, const cx2& b) { return cx2{sub2(a.re, b.re), sub2(a.im, b.im)}; }
// a * (c + j s), the same constant for both chirps
__device__ __forceinline__ cx2 cmulc(const cx2& a, float c, float s) {
  const float2 c2 = make_float2(c, c), s2 = make_float2(s, s);
  return cx2{__ffma2_rn(a.im, neg2(s2), __fmul2_rn(a.re, c2)), __ffma2_rn(a.re, s2, __fmul2_rn(a.im, c2))};
}
__device__ __forceinline__ cx2 mul_mj(const cx2& a) { return cx2{a.im, neg2(a.re)}; }   // a * (-j)

// forward 4-point DFT in place
__device__ __forceinline__ void dft4(cx2& a0, cx2& a1, cx2& a2, cx2& a3) {
  const cx2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = cx2{add2(d02.re, d13.im), sub2(d02.im, d13.re)};     // d02 + (-j) d13
  a3 = cx2{sub2(d02.re, d13.im), add2(d02.im, d13.re)};     // d02 + (+j) d13
}
// a2 = a3 = 0 / a1 = a2 = a3 = 0 (zero padding of the range FFT)
__device__ __forceinline__ void dft4_z2(cx2& a0, cx2& a1, cx2& a2, cx2& a3) {
  const cx2 x0 = a0, x1 = a1;
  a0 = cadd(x0, x1);
  a2 = csub(x0, x1);
  a1 = cx2{add2(x0.re, x1.im), sub2(x0.im, x1.re)};
  a3 = cx2{sub2(x0.re, x1.im), add2(x0.im, x1.re)};
}
__device__ __forceinline__ void dft4_z1(cx2& a0, cx2& a1, cx2& a2, cx2& a3) { a1 = a0; a2 = a0; a3 = a0; }

// forward 16-point DFT, natural order in and out; v[4*NZ..15] are zero and never read
template <int NZ>
__device__ __forceinline__ void dft16(cx2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
    if (NZ >= 3) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    else if (NZ == 2) dft4_z2(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    else dft4_z1(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  }
  v[4 * 1 + 1] = cmulc(v[4 * 1 + 1], C1, -S1);    // W16^1
  v[4 * 1 + 2] = cmulc(v[4 * 1 + 2], R2, -R2);    // W16^2
  v[4 * 1 + 3] = cmulc(v[4 * 1 + 3], S1, -C1);    // W16^3
  v[4 * 2 + 1] = cmulc(v[4 * 2 + 1], R2, -R2);    // W16^2
  v[4 * 2 + 2] = mul_mj(v[4 * 2 + 2]);            // W16^4
  v[4 * 2 + 3] = cmulc(v[4 * 2 + 3], -R2, -R2);   // W16^6
  v[4 * 3 + 1] = cmulc(v[4 * 3 + 1], S1, -C1);    // W16^3
  v[4 * 3 + 2] = cmulc(v[4 * 3 + 2], -R2, -R2);   // W16^6
  v[4 * 3 + 3] = cmulc(v[4 * 3 + 3], -C1, S1);    // W16^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) { const cx2 t = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = t; }
}


constexpr int WF_XROW = 17;                      // float4 stride between rows of the in-warp transpose
constexpr int WF_XHALF = 16 * WF_XROW;           // float4 per half-warp slice (two chirps)
#ifndef FMCW_WF_SOA
#define FMCW_WF_SOA 1
#endif
constexpr bool WF_SOA = FMCW_WF_SOA != 0;        // transpose through two float2 planes instead of one float4 plane

struct WarpSmem {                                // byte offsets inside the dynamic shared memory
  int tw, wg, wh, dtw, dwin, per_warp0, per_warp;
  int xch, rmax, csum, row, total;
};
__host__ __device__ inline WarpSmem warp_smem_layout(uint32_t PN, int WF_WARPS) {
  WarpSmem L;
  int o = 0;
  // scalars, splat into both halves of a packed operand by the instruction itself (FFMA2 R, R.F32, ...): the LSU data pipe is
  // this kernel's busiest unit (profiles/ncu_chain_r2.txt), so no table carries duplicated values
  L.tw = o;   o += 15 * 16 * (int)sizeof(float2);          // (c, s) of W256^(s*k1), k1 = 1..15
  L.wg = o;   o += NR * (int)sizeof(float);                // gw
  L.wh = o;   o += NR * (int)sizeof(float2);               // (h_re, h_im)
  L.dtw = o;  o += MAX_ND * (int)sizeof(float2);
  L.dwin = o; o += MAX_ND * (int)sizeof(float);
  L.per_warp0 = o;
  int w = 0;
  L.xch = w;  w += 2 * WF_XHALF * (int)sizeof(float4);     // 8,704 B; pass 2: X[PN] double2
  const int need2 = (int)PN * 16;
  if (w < need2) w = need2;
  L.rmax = w; w += (NR + 16) * (int)sizeof(float);         // 8 pad floats each side for the neighbour loads
  L.csum = w; w += (int)PN * (int)sizeof(float2);
  L.row = w;  w += 2 * MAX_ND * (int)sizeof(float2);       // Doppler row + windowed, mean-removed copy
  L.per_warp = (w + 15) & ~15;
  L.total = L.per_warp0 + WF_WARPS * L.per_warp;
  return L;
}

// EXACT: NTS == 64*NZ and PN % 4 == 0 and no range-spectrum export: no predicates anywhere in pass 1.
template <int NZ, bool EXACT, int MINB, int WF_WARPS, int UNPK>
__global__ void __launch_bounds__(WF_WARPS * 32, MINB) frame_chain_warp_kernel(const ChainParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t NTS = p.NTS, PN = p.PN, ND = p.ND;
  const WarpSmem L = warp_smem_layout(PN, WF_WARPS);
  float2* s_tw = reinterpret_cast<float2*>(smem_raw + L.tw);
  float* s_wg = reinterpret_cast<float*>(smem_raw + L.wg);
  float2* s_wh = reinterpret_cast<float2*>(smem_raw + L.wh);
  float2* s_dtw = reinterpret_cast<float2*>(smem_raw + L.dtw);
  float* s_dwin = reinterpret_cast<float*>(smem_raw + L.dwin);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = lane & 15, half = lane >> 4;
  const int ndc = (int)min(PN, ND);

  // ---- read-only tables, once per CTA (the grid is persistent over frames) ----
  for (int i = tid; i < 15 * 16; i += WF_WARPS * 32) {
    s_tw[i] = p.tw_pair[16 + i];
  }
  for (int i = tid; i < NR; i += WF_WARPS * 32) {
    const float4 w = (i < (int)p.nts_fft) ? p.win_tab[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    s_wg[i] = w.x;
    s_wh[i] = make_float2(w.y, w.z);
  }
  if (tid < (int)ND) s_dtw[tid] = p.dop_tw[tid];
  if (tid < MAX_ND) s_dwin[tid] = (tid < ndc) ? p.dop_win[tid] : 0.f;
  __syncthreads();

  unsigned char* wbase = smem_raw + L.per_warp0 + warp * L.per_warp;
  float4* xch = reinterpret_cast<float4*>(wbase + L.xch) + half * WF_XHALF;
  float* s_rmax = reinterpret_cast<float*>(wbase + L.rmax) + 8;
  float2* s_csum = reinterpret_cast<float2*>(wbase + L.csum);
  float2* s_row = reinterpret_cast<float2*>(wbase + L.row);
  float2* s_xw = s_row + MAX_ND;
  double2* s_X = reinterpret_cast<double2*>(wbase + L.xch);

  if (lane < 8) { s_rmax[-8 + lane] = 0.f; s_rmax[NR + lane] = 0.f; }
  const float2 nts2 = make_float2((float)NTS, (float)NTS);
  const uint32_t n_quads = (PN + 3) >> 2;

  const uint64_t stride = (uint64_t)gridDim.x * WF_WARPS;
  for (uint64_t f = (uint64_t)blockIdx.x * WF_WARPS + warp; f < p.n_frames; f += stride) {
    const uint32_t* fbase = p.iq + ((f * p.n_rx + p.rx_sel) * (uint64_t)PN) * NTS;
    float mx[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) mx[i] = 0.f;

    // ================= pass 1: range FFT of every chirp, running max of |X|^2 =================
    constexpr bool PF = (NZ <= 2) && (MINB <= 3);       // software pipelining: the next quad's samples are requested before this FFT
    uint32_t wa[4 * NZ], wb[4 * NZ];
    auto load_quad = [&](uint32_t q) {
      const uint32_t ca = 4 * q + half, cb = ca + 2;
      if (EXACT) {
        const uint32_t* pa = fbase + (uint64_t)ca * NTS + s;
        const uint32_t* pb = pa + 2 * NTS;
        if (q < n_quads) {
#pragma unroll
          for (int r = 0; r < 4 * NZ; ++r) { wa[r] = __ldg(pa + 16 * r); wb[r] = __ldg(pb + 16 * r); }
        }
      } else {
        const bool oka = q < n_quads && ca < PN, okb = q < n_quads && cb < PN;
        const uint32_t* pa = fbase + (uint64_t)(oka ? ca : 0) * NTS;
        const uint32_t* pb = fbase + (uint64_t)(okb ? cb : 0) * NTS;
#pragma unroll
        for (int r = 0; r < 4 * NZ; ++r) {
          const uint32_t n = s + 16 * r;
          wa[r] = (oka && n < NTS) ? __ldg(pa + n) : 0u;
          wb[r] = (okb && n < NTS) ? __ldg(pb + n) : 0u;
        }
      }
    };
    if (PF) load_quad(0);
    for (uint32_t q = 0; q < n_quads; ++q) {
      if (!PF) load_quad(q);
      const uint32_t ca = 4 * q + half, cb = ca + 2;
      const bool oka = EXACT || ca < PN, okb = EXACT || cb < PN;
      float2 fi[4 * NZ], fq[4 * NZ];
      float2 sI = make_float2(0.f, 0.f), sQ = make_float2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 4 * NZ; ++r) {
        if (UNPK == 0) {
          fi[r] = make_float2(cvt_lo(wa[r]), cvt_lo(wb[r]));     // I2F.F32.S16 on the half-registers: no unpack instructions (XU pipe)
          fq[r] = make_float2(cvt_hi(wa[r]), cvt_hi(wb[r]));
        } else {
          // ALU + FMA pipes instead: the sign-flipped code dropped into the mantissa of 2^23, one packed subtract removes the bias
          const uint32_t ta = wa[r] ^ 0x80008000u, tb = wb[r] ^ 0x80008000u;
          const float2 bias2 = make_float2(8421376.0f, 8421376.0f);       // 2^23 + 2^15
          fi[r] = sub2(make_float2(__uint_as_float(__byte_perm(ta, 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(tb, 0x4B000000u, 0x7610))), bias2);
          fq[r] = sub2(make_float2(__uint_as_float(__byte_perm(ta, 0x4B000000u, 0x7632)), __uint_as_float(__byte_perm(tb, 0x4B000000u, 0x7632))), bias2);
        }
        sI = add2(sI, fi[r]);
        sQ = add2(sQ, fq[r]);
      }
      if (!EXACT) {
        // samples past the FFT length still enter the mean (RP:204 comes before the truncating fft of RP:205)
        for (uint32_t n = s + 64 * NZ; n < NTS; n += 16) {
          const uint32_t w0 = oka ? __ldg(fbase + (uint64_t)ca * NTS + n) : 0u;
          const uint32_t w1 = okb ? __ldg(fbase + (uint64_t)cb * NTS + n) : 0u;
          sI = add2(sI, make_float2((float)(short)(w0 & 0xffffu), (float)(short)(w1 & 0xffffu)));
          sQ = add2(sQ, make_float2((float)((int)w0 >> 16), (float)((int)w1 >> 16)));
        }
      }
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) {
        sI.x += __shfl_xor_sync(0xffffffffu, sI.x, m);
        sI.y += __shfl_xor_sync(0xffffffffu, sI.y, m);
        sQ.x += __shfl_xor_sync(0xffffffffu, sQ.x, m);
        sQ.y += __shfl_xor_sync(0xffffffffu, sQ.y, m);
      }
      if (s == 0) {
        if (oka) s_csum[ca] = make_float2(sI.x, sQ.x);
        if (okb) s_csum[cb] = make_float2(sI.y, sQ.y);
      }

      cx2 v[16];
      const float2 nsI = neg2(sI), nsQ = neg2(sQ);
#pragma unroll
      for (int r = 0; r < 4 * NZ; ++r) {
        const int n = s + 16 * r;
        const float g1 = s_wg[n];
        const float2 h1 = s_wh[n];
        const float2 wg = make_float2(g1, g1);
        const float4 wh = make_float4(h1.x, h1.x, h1.y, h1.y);
        // (code - mean)*NTS is an exact integer (|.| < 2^24); one rounding in the windowing multiply-add
        const float2 dI = __ffma2_rn(nts2, fi[r], nsI), dQ = __ffma2_rn(nts2, fq[r], nsQ);
        v[r].re = __ffma2_rn(wg, dI, make_float2(-wh.x, -wh.y));
        v[r].im = __ffma2_rn(wg, dQ, make_float2(-wh.z, -wh.w));
        if (!EXACT) {   // chirps past PN and samples past NTS carry nothing (the table is zero past nts_fft)
          const bool live = (uint32_t)n < p.nts_fft;
          if (!(live && oka)) { v[r].re.x = 0.f; v[r].im.x = 0.f; }
          if (!(live && okb)) { v[r].re.y = 0.f; v[r].im.y = 0.f; }
        }
      }
      dft16<NZ>(v);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) {
        const float2 t = s_tw[(k1 - 1) * 16 + s];
        const float2 c2 = make_float2(t.x, t.x), s2 = make_float2(t.y, t.y);
        const cx2 a = v[k1];
        v[k1].re = __ffma2_rn(a.im, neg2(s2), __fmul2_rn(a.re, c2));
        v[k1].im = __ffma2_rn(a.re, s2, __fmul2_rn(a.im, c2));
      }
      if (WF_SOA) {
        // Re pairs and Im pairs in two float2 planes (row stride 17): 64-bit stores and loads straight from / into the register
        // pairs the packed arithmetic uses, conflict-free both ways (a 128-bit store needs four consecutive registers, which
        // cost ~4 MOVs apiece at this register pressure)
        float2* xre = reinterpret_cast<float2*>(xch);
        float2* xim = xre + WF_XHALF;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) { xre[k1 * WF_XROW + s] = v[k1].re; xim[k1 * WF_XROW + s] = v[k1].im; }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) { v[n2].re = xre[s * WF_XROW + n2]; v[n2].im = xim[s * WF_XROW + n2]; }
      } else {
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) xch[k1 * WF_XROW + s] = make_float4(v[k1].re.x, v[k1].re.y, v[k1].im.x, v[k1].im.y);
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
          const float4 t = xch[s * WF_XROW + n2];
          v[n2].re = make_float2(t.x, t.y);
          v[n2].im = make_float2(t.z, t.w);
        }
      }
      __syncwarp();
      if (PF) load_quad(q + 1);   // in flight during the second radix-16 stage and the next unpack
      dft16<4>(v);   // v[k2] = X[s + 16*k2] of both chirps
      if (!EXACT && p.spec_out != nullptr && f == p.spec_frame) {
        if (oka && ca == p.spec_chirp) {
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) p.spec_out[s + 16 * k2] = sqrtf(fmaf(v[k2].re.x, v[k2].re.x, v[k2].im.x * v[k2].im.x));
        }
        if (okb && cb == p.spec_chirp) {
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) p.spec_out[s + 16 * k2] = sqrtf(fmaf(v[k2].re.y, v[k2].re.y, v[k2].im.y * v[k2].im.y));
        }
      }
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const float2 pw = __ffma2_rn(v[k2].re, v[k2].re, __fmul2_rn(v[k2].im, v[k2].im));
        mx[k2] = fmaxf(mx[k2], fmaxf(pw.x, pw.y));
      }
    }

    // ================= max over chirps (RP:210), range_fft column (RP:265) =================
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) mx[k2] = fmaxf(mx[k2], __shfl_xor_sync(0xffffffffu, mx[k2], 16));
#pragma unroll
    for (int j = 0; j < 8; ++j) {            // each half takes eight of the sixteen k2 (both hold the same maxima)
      const float m0 = half ? mx[2 * j + 1] : mx[2 * j];
      s_rmax[s + 16 * (2 * j + half)] = sqrtf(m0);
    }
    __syncwarp();
    // lane l owns bins 8l .. 8l+7; r[2..9] are its own, r[0,1] / r[10,11] the neighbours' (zero outside 0..255)
    float r[12];
    {
      const float4 a = *reinterpret_cast<const float4*>(s_rmax + 8 * lane);
      const float4 b = *reinterpret_cast<const float4*>(s_rmax + 8 * lane + 4);
      const float2 lo = *reinterpret_cast<const float2*>(s_rmax + 8 * lane - 2);
      const float2 hi = *reinterpret_cast<const float2*>(s_rmax + 8 * lane + 8);
      r[0] = lo.x; r[1] = lo.y; r[2] = a.x; r[3] = a.y; r[4] = a.z; r[5] = a.w;
      r[6] = b.x; r[7] = b.y; r[8] = b.z; r[9] = b.w; r[10] = hi.x; r[11] = hi.y;
      if (p.range_max_abs) {
        float4* dst = reinterpret_cast<float4*>(p.range_max_abs + f * NR + 8 * lane);
        dst[0] = a;
        dst[1] = b;
      }
    }

    // ================= f_search_peak (RP:211; shim definition, see oracle) =================
    unsigned best_v = 0u, best_b = 0xffffffffu;     // strongest candidate of this lane (value bits, bin); lowest bin on ties
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int b = 8 * lane + i;
      const float fp = r[i + 2];
      const bool cand = b >= 2 && b <= NR - 3 && b >= p.bin_lo && b <= p.bin_hi && fp >= p.range_thr && fp >= r[i] &&
                        fp >= r[i + 1] && fp > r[i + 3] && fp > r[i + 4];
      if (cand) {
        const unsigned vb = (p.peak_mode == 0) ? __float_as_uint(fp) : 1u;
        if (best_b == 0xffffffffu || vb > best_v) { best_v = vb; best_b = (unsigned)b; }
      }
    }
    const unsigned top_v = __reduce_max_sync(0xffffffffu, best_b == 0xffffffffu ? 0u : best_v);
    const unsigned top_b = __reduce_min_sync(0xffffffffu, (best_b != 0xffffffffu && best_v == top_v) ? best_b : 0xffffffffu);
    const bool det = top_b != 0xffffffffu;
    const int kbin = det ? (int)top_b : -1;

    if (lane == 0) {
      if (p.detected) p.detected[f] = det ? 1 : 0;
      if (p.range_bin) p.range_bin[f] = kbin;
      if (p.range_mag) p.range_mag[f] = det ? s_rmax[kbin] : 0.f;
    }
    if (!det) {
      for (uint32_t c = lane; c < PN; c += 32) { p.slow64[f * PN + c] = 0.0; if (p.slow_mag) p.slow_mag[f * PN + c] = 0.f; }
      if (p.doppler_row) for (uint32_t i = lane; i < ND; i += 32) p.doppler_row[f * ND + i] = make_float2(0.f, 0.f);
      if (p.doppler_bin && lane == 0) p.doppler_bin[f] = (int)ND / 2;
      __syncwarp();
      continue;
    }

    // ================= pass 2: slow-time row at the selected bin (single-bin DFT, float64) =================
    // X_c[k*] = sum_n G[n] d_c[n] - H[k*],  G[n] = gw[n] W^(n k*),  d_c[n] = NTS*code - sum (the exact integer of
    // pass 1),  H = FFT of the calibration term (tabulated per bin at create).  This row feeds the STFT, whose bins
    // sit 100+ dB under its DC term, so it is carried in float64.
    __syncwarp();                              // every lane is done with the transpose slices
    // 32 lanes per chirp, lane l owns samples l + 32 i: G stays in registers (2 NZ values), a chirp is four (NZ = 2) fully
    // coalesced 128-byte loads, and eight chirps are reduced over the warp together (a halving butterfly: 36 shuffles per
    // eight chirps).  The LSU data pipe is this kernel's busiest unit; this pass no longer touches shared memory in its loop.
    constexpr int NG = 2 * NZ;
    double2 G[NG];
    double2 Gs = make_double2(0.0, 0.0);       // sum_n G[n]
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      const int n = lane + 32 * i;
      G[i] = make_double2(0.0, 0.0);
      if (EXACT || n < (int)p.nts_fft) {
        const double gw = __ldg(p.win_tab_d + 3 * n);
        const double2 tw = __ldg(p.tw_d + (((uint32_t)n * (uint32_t)kbin) & (NR - 1)));
        G[i] = make_double2(gw * tw.x, gw * tw.y);
      }
      Gs.x += G[i].x;
      Gs.y += G[i].y;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      Gs.x += __shfl_xor_sync(0xffffffffu, Gs.x, m);
      Gs.y += __shfl_xor_sync(0xffffffffu, Gs.y, m);
    }
    const double2 Hk = __ldg(p.hfft_d + kbin);
    {
      constexpr bool PF2 = (NZ <= 2);          // NZ = 4: eight chirps x eight words already take 64 registers
      uint32_t wcur[8][NG], wnxt[PF2 ? 8 : 1][PF2 ? NG : 1];
      auto load_batch = [&](uint32_t c0, auto& w) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t c = c0 + j;
          const bool live = EXACT || c < PN;
          const uint32_t* cbp = fbase + (uint64_t)(live ? c : 0) * NTS + lane;
#pragma unroll
          for (int i = 0; i < NG; ++i)
            w[j][i] = (live && (EXACT || (uint32_t)(lane + 32 * i) < p.nts_fft)) ? __ldg(cbp + 32 * i) : 0u;
        }
      };
      const bool b16 = (lane & 16) != 0, b8 = (lane & 8) != 0, b4 = (lane & 4) != 0;
      const uint32_t my_chirp = (b16 ? 4u : 0u) + (b8 ? 2u : 0u) + (b4 ? 1u : 0u);      // the chirp of the batch this lane ends up with
      if (PF2) load_batch(0, wcur);
      for (uint32_t c0 = 0; c0 < PN; c0 += 8) {
        if (!PF2) load_batch(c0, wcur);
        if constexpr (PF2) { if (c0 + 8 < PN) load_batch(c0 + 8, wnxt); }         // the next eight chirps are in flight during this step
        double ar[8], ai[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double r0 = 0.0, i0 = 0.0;
#pragma unroll
          for (int i = 0; i < NG; ++i) {
            const double cI = cvtd_lo(wcur[j][i]), cQ = cvtd_hi(wcur[j][i]);
            r0 = fma(G[i].x, cI, fma(-G[i].y, cQ, r0));
            i0 = fma(G[i].x, cQ, fma(G[i].y, cI, i0));
          }
          ar[j] = r0; ai[j] = i0;
        }
        // halving butterfly: after xor 16 a lane keeps four chirps, after xor 8 two, after xor 4 one; then a plain reduction
        double r4[4], i4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double keep_r = b16 ? ar[j + 4] : ar[j], send_r = b16 ? ar[j] : ar[j + 4];
          const double keep_i = b16 ? ai[j + 4] : ai[j], send_i = b16 ? ai[j] : ai[j + 4];
          r4[j] = keep_r + __shfl_xor_sync(0xffffffffu, send_r, 16);
          i4[j] = keep_i + __shfl_xor_sync(0xffffffffu, send_i, 16);
        }
        double r2[2], i2[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double keep_r = b8 ? r4[j + 2] : r4[j], send_r = b8 ? r4[j] : r4[j + 2];
          const double keep_i = b8 ? i4[j + 2] : i4[j], send_i = b8 ? i4[j] : i4[j + 2];
          r2[j] = keep_r + __shfl_xor_sync(0xffffffffu, send_r, 8);
          i2[j] = keep_i + __shfl_xor_sync(0xffffffffu, send_i, 8);
        }
        double ar1 = (b4 ? r2[1] : r2[0]) + __shfl_xor_sync(0xffffffffu, b4 ? r2[0] : r2[1], 4);
        double ai1 = (b4 ? i2[1] : i2[0]) + __shfl_xor_sync(0xffffffffu, b4 ? i2[0] : i2[1], 4);
        ar1 += __shfl_xor_sync(0xffffffffu, ar1, 2);
        ai1 += __shfl_xor_sync(0xffffffffu, ai1, 2);
        ar1 += __shfl_xor_sync(0xffffffffu, ar1, 1);
        ai1 += __shfl_xor_sync(0xffffffffu, ai1, 1);
        const uint32_t c = c0 + my_chirp;
        if ((lane & 3) == 0 && c < PN) {
          // sum_n G[n] (NTS c[n] - sum c) = NTS * acc - (sum c) * Gsum: the codes enter as they are, the mean leaves here
          const float2 csf = s_csum[c];
          const double sI = (double)csf.x, sQ = (double)csf.y;
          s_X[c] = make_double2((double)NTS * ar1 - (sI * Gs.x - sQ * Gs.y) - Hk.x, (double)NTS * ai1 - (sI * Gs.y + sQ * Gs.x) - Hk.y);
        }
        if constexpr (PF2) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < NG; ++i) wcur[j][i] = wnxt[j][i];
        }
      }
    }
    __syncwarp();
    // magnitudes (RP:259 + RP:270), mean over all PN chirps (RP:217)
    float sr = 0.f, si = 0.f;
    for (uint32_t c = lane; c < PN; c += 32) {
      const double2 X = s_X[c];
      const double mag = sqrt(X.x * X.x + X.y * X.y);
      p.slow64[f * PN + c] = mag;
      if (p.slow_mag) p.slow_mag[f * PN + c] = (float)mag;
      const float xr = (float)X.x, xi = (float)X.y;
      sr += xr;
      si += xi;
      if (c < (uint32_t)ndc) s_row[c] = make_float2(xr, xi);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, m);
      si += __shfl_xor_sync(0xffffffffu, si, m);
    }
    const float mr = sr / (float)PN, mi = si / (float)PN;
    __syncwarp();
    for (int c = lane; c < ndc; c += 32) {
      const float2 rr = s_row[c];
      const float w = s_dwin[c];
      s_xw[c] = make_float2((rr.x - mr) * w, (rr.y - mi) * w);
    }
    __syncwarp();

    // Doppler FFT over the first min(PN, ND) chirps, fftshift, first max, threshold (RP:219, 233-238)
    unsigned dv = 0u, di_ = 0xffffffffu;
    for (int i0 = 0; i0 < (int)ND; i0 += 32) {
      const int i = i0 + lane;
      if (i < (int)ND) {
        const int k = (i + (int)ND / 2) & ((int)ND - 1);
        float dr = 0.f, di = 0.f;
        for (int c = 0; c < ndc; ++c) {
          const float2 x = s_xw[c];
          const float2 t = s_dtw[(c * k) & ((int)ND - 1)];
          dr = fmaf(x.x, t.x, fmaf(-x.y, t.y, dr));
          di = fmaf(x.x, t.y, fmaf(x.y, t.x, di));
        }
        if (p.doppler_row) p.doppler_row[f * ND + i] = make_float2(dr, di);
        const unsigned ab = __float_as_uint(sqrtf(fmaf(dr, dr, di * di)));
        if (di_ == 0xffffffffu || ab > dv) { dv = ab; di_ = (unsigned)i; }
      }
    }
    const unsigned tv = __reduce_max_sync(0xffffffffu, di_ == 0xffffffffu ? 0u : dv);
    const unsigned ti = __reduce_min_sync(0xffffffffu, (di_ != 0xffffffffu && dv == tv) ? di_ : 0xffffffffu);
    if (lane == 0 && p.doppler_bin) p.doppler_bin[f] = (__uint_as_float(tv) >= p.dop_thr) ? (int)ti : (int)ND / 2;
    __syncwarp();   // the per-warp shared memory is rewritten by the next frame
  }
}

template <int NZ, bool EXACT, int MINB, int WF_WARPS = 4, int UNPK = 0>
cudaError_t launch_variant(const ChainParams& p, int sms, cudaStream_t st) {
  const WarpSmem L = warp_smem_layout(p.PN, WF_WARPS);
  const int per_sm = MINB;
  const uint64_t ctas_needed = (p.n_frames + WF_WARPS - 1) / WF_WARPS;
  const uint64_t max_grid = (uint64_t)sms * per_sm;
  const unsigned grid = (unsigned)(ctas_needed < max_grid ? ctas_needed : max_grid);
  cudaError_t e = cudaFuncSetAttribute(frame_chain_warp_kernel<NZ, EXACT, MINB, WF_WARPS, UNPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) return e;
  frame_chain_warp_kernel<NZ, EXACT, MINB, WF_WARPS, UNPK><<<grid, WF_WARPS * 32, L.total, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

bool chain_warp_supported(const ChainParams& p) {
  // float holds NTS*code - sum exactly while NTS * 32768 < 2^24; the per-warp shared memory grows with PN
  return p.NTS <= 511 && p.PN <= 1024 && warp_smem_layout(p.PN, 4).total <= 100 * 1024;
}

cudaError_t launch_frame_chain_warp(const ChainParams& p, cudaStream_t st) {
  if (p.n_frames == 0) return cudaSuccess;
  const int sms = device_sm_count();
  const int nz = p.nts_fft <= 64 ? 1 : p.nts_fft <= 128 ? 2 : 4;
  const bool exact = p.NTS == (uint32_t)(64 * nz) && (p.PN & 3u) == 0 && p.spec_out == nullptr;
  static const int minb = env_int("FMCW_CHAIN_MINB", 3) == 4 ? 4 : 3;
#define FMCW_WF(NZ_)                                                                                        \
  do {                                                                                                      \
    if (minb == 4 && NZ_ != 4)                                                                              \
      return exact ? launch_variant<NZ_, true, (NZ_ == 4 ? 3 : 4)>(p, sms, st) : launch_variant<NZ_, false, (NZ_ == 4 ? 3 : 4)>(p, sms, st); \
    return exact ? launch_variant<NZ_, true, 3>(p, sms, st) : launch_variant<NZ_, false, 3>(p, sms, st);    \
  } while (0)
  static const int unpk = env_int("FMCW_CHAIN_UNPACK", 0);
  if (unpk == 1 && nz == 2 && exact) return launch_variant<2, true, 3, 4, 1>(p, sms, st);
  if (nz == 1) FMCW_WF(1);
  if (nz == 2) FMCW_WF(2);
  FMCW_WF(4);
#undef FMCW_WF
}

}  // namespace fmcw
