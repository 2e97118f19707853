// Per-frame chain of the reference (radar_processing.m RP:199-260), one CTA per frame:
//   unpack int16 I/Q -> calibration + IF scale + fast-time mean removal (RP:203-204, done on exact
//   integers) -> 2*blackman window -> 256-point range FFT (RP:205) -> max |.| over chirps (RP:210)
//   -> f_search_peak (RP:211) -> slow-time row at the selected bin (RP:259) -> mean removal over chirps,
//   2*chebwin window, ND-point Doppler FFT, fftshift (RP:217-219) -> first-max + threshold (RP:233-238).
//
// FFT core: 256 = 16 x 16.  Sixteen lanes own one chirp; each lane runs a radix-16 butterfly in
// registers on the stride-16 samples (pruned for the zero padding), multiplies by W_256^(s*k1),
// transposes through a private, bank-conflict-free slice of shared memory (only __syncwarp, the two
// chirps of a warp never leave the warp) and runs the second radix-16.  Nothing of the 256 x PN
// range cube is written to HBM (RP:207's range_tx1rx1_complete is never materialised): the slow-time
// row of the one selected bin is re-evaluated as a single-bin DFT from the L1/L2-resident samples.
#include <cstdlib>

#include "fmcw_internal.cuh"

namespace fmcw {

// complex add / subtract as one packed FADD2 (sm_100 f32x2 pipe)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  float2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  float2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return r;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }   // a * (-j)
__device__ __forceinline__ float2 mul_pj(float2 a) { return make_float2(-a.y, a.x); }   // a * (+j)

// forward 4-point DFT in place: (a0,a1,a2,a3) -> (X0,X1,X2,X3)
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = make_float2(d02.x + d13.y, d02.y - d13.x);     // d02 + (-j) d13
  a3 = make_float2(d02.x - d13.y, d02.y + d13.x);     // d02 + (+j) d13
}
// same with a2 = a3 = 0 / with a1 = a2 = a3 = 0 (zero padding of the range FFT)
__device__ __forceinline__ void dft4_z2(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 x0 = a0, x1 = a1;
  a0 = cadd(x0, x1);
  a2 = csub(x0, x1);
  a1 = make_float2(x0.x + x1.y, x0.y - x1.x);
  a3 = make_float2(x0.x - x1.y, x0.y + x1.x);
}
__device__ __forceinline__ void dft4_z1(float2& a0, float2& a1, float2& a2, float2& a3) { a1 = a0; a2 = a0; a3 = a0; }

// forward 16-point DFT, natural order in and out.  NZ = number of non-zero groups of four inputs
// (v[4*NZ..15] are treated as zero and never read).
template <int NZ>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
  // n = 4*n1 + n2: stage A is a DFT over n1 for every n2; result a[n2][k1] kept in v[4*k1 + n2]
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
    if (NZ >= 3) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    else if (NZ == 2) dft4_z2(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    else dft4_z1(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  }
  // twiddles W_16^(n2*k1)
  v[4 * 1 + 1] = cmul(v[4 * 1 + 1], make_float2(C1, -S1));   // m = 1
  v[4 * 1 + 2] = cmul(v[4 * 1 + 2], make_float2(R2, -R2));   // m = 2
  v[4 * 1 + 3] = cmul(v[4 * 1 + 3], make_float2(S1, -C1));   // m = 3
  v[4 * 2 + 1] = cmul(v[4 * 2 + 1], make_float2(R2, -R2));   // m = 2
  v[4 * 2 + 2] = mul_mj(v[4 * 2 + 2]);                       // m = 4
  v[4 * 2 + 3] = cmul(v[4 * 2 + 3], make_float2(-R2, -R2));  // m = 6
  v[4 * 3 + 1] = cmul(v[4 * 3 + 1], make_float2(S1, -C1));   // m = 3
  v[4 * 3 + 2] = cmul(v[4 * 3 + 2], make_float2(-R2, -R2));  // m = 6
  v[4 * 3 + 3] = cmul(v[4 * 3 + 3], make_float2(-C1, S1));   // m = 9
  // stage B: DFT over n2 for every k1 -> X[k1 + 4*k2] in v[4*k1 + k2]
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // transpose the 4x4 register tile to natural order
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) { float2 t = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = t; }
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
int device_sm_count() {
  static const int sms = [] { int dev = 0, n = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n > 0 ? n : 148; }();
  return sms;
}

static size_t chain_smem_bytes_nw(uint32_t PN, int nw) {
  size_t b = (size_t)nw * 2 * XCH_CHIRP * sizeof(float2);             // transpose slices (4,352 B per warp)
  b += 256 * sizeof(float2);                                          // tw_pair
  b += 2 * 272 * sizeof(float);                                       // skewed W_256
  b += NR * sizeof(float4);                                           // window / calibration table
  b += NR * sizeof(float);                                            // rmax
  b += (size_t)PN * sizeof(int2);                                     // per-chirp integer sums
  b += (size_t)PN * sizeof(float2);                                   // slow-time row
  b += MAX_ND * sizeof(float2) + MAX_ND * sizeof(float);              // Doppler twiddles + window
  b += 64 * sizeof(unsigned long long);                               // reduction scratch
  return b;
}
size_t chain_smem_bytes(uint32_t PN) { return chain_smem_bytes_nw(PN, CHAIN_WARPS); }

// NT threads per CTA (256 or 128): the smaller CTA interleaves the per-frame serial sections (peak search, Doppler
// row) of more frames on one SM.
template <int NZ, int NT>
__global__ void __launch_bounds__(NT, 768 / NT) frame_chain_kernel(const ChainParams p) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* xch = reinterpret_cast<float2*>(smem_raw);
  float2* s_twp = xch + NW * 2 * XCH_CHIRP;
  float* s_twre = reinterpret_cast<float*>(s_twp + 256);
  float* s_twim = s_twre + 272;
  float4* s_win = reinterpret_cast<float4*>(s_twim + 272);
  float* s_rmax = reinterpret_cast<float*>(s_win + NR);
  int2* s_csum = reinterpret_cast<int2*>(s_rmax + NR);
  float2* s_row = reinterpret_cast<float2*>(s_csum + p.PN);
  float2* s_dtw = s_row + p.PN;
  float* s_dwin = reinterpret_cast<float*>(s_dtw + MAX_ND);
  unsigned long long* s_red = reinterpret_cast<unsigned long long*>(s_dwin + MAX_ND);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = lane & 15, half = lane >> 4;
  const uint32_t NTS = p.NTS, PN = p.PN, ND = p.ND;
  const int ndc = (int)min(PN, ND);

  // ---- tables (once per CTA; the grid is persistent over frames) ----
  for (int i = tid; i < 256; i += NT) s_twp[i] = p.tw_pair[i];
  for (int i = tid; i < 272; i += NT) { s_twre[i] = p.tw_re[i]; s_twim[i] = p.tw_im[i]; }
  for (int i = tid; i < NR; i += NT) s_win[i] = (i < (int)p.nts_fft) ? p.win_tab[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < (int)ND) s_dtw[tid] = p.dop_tw[tid];
  if (tid < ndc) s_dwin[tid] = p.dop_win[tid];
  __syncthreads();

  float2* my_x = xch + (warp * 2 + half) * XCH_CHIRP;

  for (uint64_t f = blockIdx.x; f < p.n_frames; f += gridDim.x) {
    const uint32_t* fbase = p.iq + ((f * p.n_rx + p.rx_sel) * (uint64_t)PN) * NTS;
    float mx[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) mx[i] = 0.f;

    // ================= pass 1: range FFT of every chirp, running max of |X|^2 =================
    // software pipelining: the samples of the next chirp pair are requested before this pair's FFT
    constexpr bool PF = (NZ <= 2);       // (the 16-word prefetch of NZ = 4 would spill)
    uint32_t wnext[4 * NZ];
    if (PF) {
      const uint32_t c = warp * 2 + half;
      const uint32_t* cb = fbase + (uint64_t)(c < PN ? c : 0) * NTS;
#pragma unroll
      for (int r = 0; r < 4 * NZ; ++r) {
        const uint32_t n = s + 16 * r;
        wnext[r] = (c < PN && n < NTS) ? __ldg(cb + n) : 0u;
      }
    }
    for (uint32_t pair = warp; pair * 2 < PN; pair += NW) {
      const uint32_t c = pair * 2 + half;
      const bool active = c < PN;
      const uint32_t* cb = fbase + (uint64_t)(active ? c : 0) * NTS;
      int ci[4 * NZ], cq[4 * NZ];
      int sumI = 0, sumQ = 0;
      if (!PF) {
#pragma unroll
        for (int r = 0; r < 4 * NZ; ++r) {
          const uint32_t n = s + 16 * r;
          wnext[r] = (active && n < NTS) ? __ldg(cb + n) : 0u;
        }
      }
#pragma unroll
      for (int r = 0; r < 4 * NZ; ++r) {
        const uint32_t w = wnext[r];
        ci[r] = (int)(short)(w & 0xffffu);
        cq[r] = (int)w >> 16;
        sumI += ci[r];
        sumQ += cq[r];
      }
      if (PF) {
        const uint32_t cn = (pair + NW) * 2 + half;
        const uint32_t* cbn = fbase + (uint64_t)(cn < PN ? cn : 0) * NTS;
#pragma unroll
        for (int r = 0; r < 4 * NZ; ++r) {
          const uint32_t n = s + 16 * r;
          wnext[r] = (cn < PN && n < NTS) ? __ldg(cbn + n) : 0u;
        }
      }
      if (active)
        for (uint32_t n = s + NR; n < NTS; n += 16) {   // samples past the FFT length still enter the mean (RP:204)
          uint32_t w = __ldg(cb + n);
          sumI += (int)(short)(w & 0xffffu);
          sumQ += (int)w >> 16;
        }
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) {
        sumI += __shfl_xor_sync(0xffffffffu, sumI, m);
        sumQ += __shfl_xor_sync(0xffffffffu, sumQ, m);
      }
      if (active && s == 0) s_csum[c] = make_int2(sumI, sumQ);

      float2 v[16];
#pragma unroll
      for (int r = 0; r < 4 * NZ; ++r) {
        const float4 wt = s_win[s + 16 * r];
        // (code - mean)*NTS is an exact integer; one rounding in the fused multiply-add
        const float dI = (float)((int)NTS * ci[r] - sumI), dQ = (float)((int)NTS * cq[r] - sumQ);
        const bool live = (uint32_t)(s + 16 * r) < p.nts_fft;
        v[r] = live ? make_float2(fmaf(wt.x, dI, -wt.y), fmaf(wt.x, dQ, -wt.z)) : make_float2(0.f, 0.f);
      }
      dft16<NZ>(v);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        float2 a = (k1 == 0) ? v[0] : cmul(v[k1], s_twp[k1 * 16 + s]);
        my_x[k1 * XCH_STRIDE + s] = a;
      }
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) v[n2] = my_x[s * XCH_STRIDE + n2];
      __syncwarp();
      dft16<4>(v);   // v[k2] = X[s + 16*k2]
      if (p.spec_out != nullptr && active && f == p.spec_frame && c == p.spec_chirp) {
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) p.spec_out[s + 16 * k2] = sqrtf(fmaf(v[k2].x, v[k2].x, v[k2].y * v[k2].y));
      }
      if (active) {
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) mx[k2] = fmaxf(mx[k2], fmaf(v[k2].x, v[k2].x, v[k2].y * v[k2].y));
      }
    }
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) mx[k2] = fmaxf(mx[k2], __shfl_xor_sync(0xffffffffu, mx[k2], 16));
    __syncthreads();   // every warp is done with its transpose slice; reuse it for the cross-warp max
    float* s_mx = reinterpret_cast<float*>(xch);
    if (half == 0) {
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) s_mx[warp * NR + s + 16 * k2] = mx[k2];
    }
    __syncthreads();
    for (int b = tid; b < NR; b += NT) {
      float m = s_mx[b];
#pragma unroll
      for (int w = 1; w < NW; ++w) m = fmaxf(m, s_mx[w * NR + b]);
      const float r = sqrtf(m);
      s_rmax[b] = r;
      if (p.range_max_abs) p.range_max_abs[f * NR + b] = r;
    }
    __syncthreads();

    // ================= f_search_peak (RP:211; shim definition, see oracle) =================
    unsigned long long key = 0ull;
    for (int b = tid; b < NR; b += NT) {
      if (b >= 2 && b <= NR - 3 && b >= p.bin_lo && b <= p.bin_hi) {
        const float fp = s_rmax[b];
        if (fp >= p.range_thr && fp >= s_rmax[b - 2] && fp >= s_rmax[b - 1] && fp > s_rmax[b + 1] && fp > s_rmax[b + 2]) {
          const unsigned long long lo = 0xffffffffull - (unsigned)b;
          const unsigned long long kb = (p.peak_mode == 0) ? (((unsigned long long)__float_as_uint(fp) << 32) | lo) : ((1ull << 32) | lo);
          key = kb > key ? kb : key;
        }
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { unsigned long long o = shfl_xor_u64(key, m); key = o > key ? o : key; }
    if (lane == 0) s_red[warp] = key;
    __syncthreads();
    key = s_red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) key = s_red[w] > key ? s_red[w] : key;
    const bool det = key != 0ull;
    const int kbin = det ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : -1;
    __syncthreads();   // s_red is reused below

    if (tid == 0) {
      if (p.detected) p.detected[f] = det ? 1 : 0;
      if (p.range_bin) p.range_bin[f] = kbin;
      if (p.range_mag) p.range_mag[f] = det ? s_rmax[kbin] : 0.f;
    }
    if (!det) {
      for (uint32_t c = tid; c < PN; c += NT) { p.slow64[f * PN + c] = 0.0; if (p.slow_mag) p.slow_mag[f * PN + c] = 0.f; }
      if (p.doppler_row && tid < (int)ND) p.doppler_row[f * ND + tid] = make_float2(0.f, 0.f);
      if (p.doppler_bin && tid == 0) p.doppler_bin[f] = (int)ND / 2;
      continue;
    }

    // ================= pass 2: slow-time row at the selected bin (single-bin DFT, float64) =================
    // X_c[k*] = sum_n (gw[n]*d_c[n] - h[n]) W^(n k*) = sum_n G[n]*d_c[n] - H with G[n] = gw[n] W^(n k*) tabulated once
    // per frame; d_c[n] = NTS*code - sum is the exact integer of pass 1.  This row feeds the STFT, whose bins sit
    // 100+ dB under its DC term, so it is carried in float64 (8,192 complex MACs per frame).
    double2* s_G = reinterpret_cast<double2*>(xch);          // the transpose slices are free again
    double2* s_h = reinterpret_cast<double2*>(xch) + NR;     // [CHAIN_WARPS]
    {
      double hr = 0.0, hi = 0.0;
      for (int n = tid; n < (int)p.nts_fft; n += NT) {
        const double gw = p.win_tab_d[3 * n], h_re = p.win_tab_d[3 * n + 1], h_im = p.win_tab_d[3 * n + 2];
        const double2 tw = p.tw_d[((uint32_t)n * (uint32_t)kbin) & (NR - 1)];
        s_G[n] = make_double2(gw * tw.x, gw * tw.y);
        hr += h_re * tw.x - h_im * tw.y;
        hi += h_re * tw.y + h_im * tw.x;
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        hr += __shfl_xor_sync(0xffffffffu, hr, m);
        hi += __shfl_xor_sync(0xffffffffu, hi, m);
      }
      if (lane == 0) s_h[warp] = make_double2(hr, hi);
    }
    __syncthreads();
    {
      double Hr = 0.0, Hi = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) { Hr += s_h[w].x; Hi += s_h[w].y; }
      // eight lanes per chirp, four chirps per warp
      const int grp = lane >> 3, j8 = lane & 7;
      for (uint32_t c0 = warp * 4; c0 < PN; c0 += NW * 4) {
        const uint32_t c = c0 + grp;
        const bool live = c < PN;
        const uint32_t* cb = fbase + (uint64_t)(live ? c : 0) * NTS;
        const int2 cs = live ? s_csum[c] : make_int2(0, 0);
        double ar = 0.0, ai = 0.0;
#pragma unroll 4
        for (uint32_t n = j8; n < p.nts_fft; n += 8) {
          const uint32_t w = __ldg(cb + n);
          const double2 G = s_G[n];
          const double dI = (double)((int)NTS * (int)(short)(w & 0xffffu) - cs.x);
          const double dQ = (double)((int)NTS * ((int)w >> 16) - cs.y);
          ar = fma(G.x, dI, fma(-G.y, dQ, ar));
          ai = fma(G.x, dQ, fma(G.y, dI, ai));
        }
#pragma unroll
        for (int m = 4; m >= 1; m >>= 1) {
          ar += __shfl_xor_sync(0xffffffffu, ar, m);
          ai += __shfl_xor_sync(0xffffffffu, ai, m);
        }
        if (live && j8 == 0) {
          const double xr = ar - Hr, xi = ai - Hi;
          s_row[c] = make_float2((float)xr, (float)xi);
          const double mag = sqrt(xr * xr + xi * xi);          // RP:259 + RP:270
          p.slow64[f * PN + c] = mag;
          if (p.slow_mag) p.slow_mag[f * PN + c] = (float)mag;
        }
      }
    }
    __syncthreads();

    // mean over all PN chirps (RP:217)
    float sr = 0.f, si = 0.f;
    for (uint32_t c = tid; c < PN; c += NT) {
      const float2 r = s_row[c];
      sr += r.x;
      si += r.y;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, m);
      si += __shfl_xor_sync(0xffffffffu, si, m);
    }
    float2* s_sum = reinterpret_cast<float2*>(s_red);
    if (lane == 0) s_sum[warp] = make_float2(sr, si);
    __syncthreads();
    float mr = 0.f, mi = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) { mr += s_sum[w].x; mi += s_sum[w].y; }
    mr /= (float)PN;
    mi /= (float)PN;
    __syncthreads();

    // Doppler FFT over the first min(PN, ND) chirps, fftshift, first max, threshold (RP:219, 233-238)
    if (warp == 0) {
      unsigned long long dk = 0ull;
      for (int i0 = 0; i0 < (int)ND; i0 += 32) {
        const int i = i0 + lane;
        if (i < (int)ND) {
          const int k = (i + (int)ND / 2) & ((int)ND - 1);
          float dr = 0.f, di = 0.f;
          for (int c = 0; c < ndc; ++c) {
            const float2 r = s_row[c];
            const float w = s_dwin[c];
            const float xr = (r.x - mr) * w, xi = (r.y - mi) * w;
            const float2 t = s_dtw[(c * k) & ((int)ND - 1)];
            dr = fmaf(xr, t.x, fmaf(-xi, t.y, dr));
            di = fmaf(xr, t.y, fmaf(xi, t.x, di));
          }
          if (p.doppler_row) p.doppler_row[f * ND + i] = make_float2(dr, di);
          const float a = sqrtf(fmaf(dr, dr, di * di));
          const unsigned long long kk = ((unsigned long long)__float_as_uint(a) << 32) | (0xffffffffull - (unsigned)i);
          dk = kk > dk ? kk : dk;
        }
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) { unsigned long long o = shfl_xor_u64(dk, m); dk = o > dk ? o : dk; }
      if (lane == 0 && p.doppler_bin) {
        const float val = __uint_as_float((unsigned)(dk >> 32));
        const int idx = (int)(0xffffffffu - (unsigned)(dk & 0xffffffffull));
        p.doppler_bin[f] = (val >= p.dop_thr) ? idx : (int)ND / 2;
      }
    }
    __syncthreads();   // s_row / s_csum are rewritten by the next frame
  }
}

// Full range-Doppler dB map of ONE frame (an export, not on the throughput path): one thread per range bin.  The range row
// X_c[r] of every chirp by direct float64 DFT of the exact integer samples (the tables of the slow-time pass: G = gw * W^(n r),
// H = FFT of the calibration term), then RP:217-219 per row and 20*log10(|.|).
__global__ void __launch_bounds__(NR) range_doppler_map_kernel(const ChainParams p, uint64_t frame, float* __restrict__ out_db) {
  extern __shared__ __align__(16) unsigned char rd_raw[];
  int2* s_sum = reinterpret_cast<int2*>(rd_raw);                       // [PN] integer sums of every chirp (RP:204)
  const int r = threadIdx.x;
  const uint32_t NTS = p.NTS, PN = p.PN, ND = p.ND;
  const int ndc = (int)min(PN, ND);
  const uint32_t* fbase = p.iq + ((frame * p.n_rx + p.rx_sel) * (uint64_t)PN) * NTS;
  for (uint32_t c = r; c < PN; c += NR) {
    int sI = 0, sQ = 0;
    for (uint32_t n = 0; n < NTS; ++n) { const uint32_t w = __ldg(fbase + (uint64_t)c * NTS + n); sI += (int)(short)(w & 0xffffu); sQ += (int)w >> 16; }
    s_sum[c] = make_int2(sI, sQ);
  }
  __syncthreads();
  const double2 H = p.hfft_d[r];
  double mr = 0.0, mi = 0.0;
  double xr[MAX_ND], xi[MAX_ND];
  for (uint32_t c = 0; c < PN; ++c) {
    const int2 cs = s_sum[c];
    double ar = 0.0, ai = 0.0;
    for (uint32_t n = 0; n < p.nts_fft; ++n) {
      const uint32_t w = __ldg(fbase + (uint64_t)c * NTS + n);
      const double gw = p.win_tab_d[3 * n];
      const double2 tw = p.tw_d[(n * (uint32_t)r) & (NR - 1)];
      const double dI = (double)((int)NTS * (int)(short)(w & 0xffffu) - cs.x), dQ = (double)((int)NTS * ((int)w >> 16) - cs.y);
      const double gr = gw * tw.x, gi = gw * tw.y;
      ar = fma(gr, dI, fma(-gi, dQ, ar));
      ai = fma(gr, dQ, fma(gi, dI, ai));
    }
    ar -= H.x; ai -= H.y;
    mr += ar; mi += ai;                                                 // mean over ALL chirps (RP:217)
    if ((int)c < ndc) { xr[c] = ar; xi[c] = ai; }
  }
  mr /= (double)PN; mi /= (double)PN;
  for (int i = 0; i < (int)ND; ++i) {
    const int k = (i + (int)ND / 2) & ((int)ND - 1);                    // fftshift
    double dr = 0.0, di = 0.0;
    for (int c = 0; c < ndc; ++c) {
      const double w = p.dop_win_d[c];
      const double a = (xr[c] - mr) * w, b = (xi[c] - mi) * w;
      double sn, cs;
      sincospi(-2.0 * (double)((c * k) & ((int)ND - 1)) / (double)ND, &sn, &cs);
      dr += a * cs - b * sn;
      di += a * sn + b * cs;
    }
    out_db[r * ND + i] = (float)(10.0 * log10(dr * dr + di * di));      // 20*log10(|.|); -inf for an exact zero
  }
}

cudaError_t launch_range_doppler_map(const ChainParams& p, uint64_t frame, float* out_db, cudaStream_t st) {
  range_doppler_map_kernel<<<1, NR, (size_t)p.PN * sizeof(int2), st>>>(p, frame, out_db);
  return cudaGetLastError();
}

template <int NZ, int NT>
static cudaError_t launch_chain_variant(const ChainParams& p, int sms, cudaStream_t st) {
  const size_t smem = chain_smem_bytes_nw(p.PN, NT / 32);
  // persistent over frames: two resident sets of CTAs (tables are loaded once per CTA; 1 / 2 / 4 / 8 sets measured
  // 0.191 / 0.182 / 0.183 / 0.184 ms on C2)
  const uint64_t max_grid = (uint64_t)sms * (768 / NT) * 2;
  const unsigned grid = (unsigned)(p.n_frames < max_grid ? p.n_frames : max_grid);
  cudaError_t e = cudaFuncSetAttribute(frame_chain_kernel<NZ, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  frame_chain_kernel<NZ, NT><<<grid, NT, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_frame_chain(const ChainParams& p, cudaStream_t st) {
  if (p.n_frames == 0) return cudaSuccess;
  // FMCW_CHAIN_VARIANT=0: the one-CTA-per-frame kernel below (kept for A/B runs and odd shapes)
  static const int variant = env_int("FMCW_CHAIN_VARIANT", 1) == 0 ? 0 : 1;
  if (variant == 1 && chain_warp_supported(p)) return launch_frame_chain_warp(p, st);
  const int sms = device_sm_count();
  static const int nt = env_int("FMCW_CHAIN_THREADS", 128) == 256 ? 256 : 128;   // 128 measured fastest (64 / 96 / 128 / 256: 0.192 / 0.186 / 0.185 / 0.199 ms on C2)
#define FMCW_CHAIN_NT(NT)                                                  \
  do {                                                                     \
    if (p.nts_fft <= 64) return launch_chain_variant<1, NT>(p, sms, st);   \
    if (p.nts_fft <= 128) return launch_chain_variant<2, NT>(p, sms, st);  \
    return launch_chain_variant<4, NT>(p, sms, st);                        \
  } while (0)
  if (nt == 256) FMCW_CHAIN_NT(256);
  FMCW_CHAIN_NT(128);
#undef FMCW_CHAIN_NT
}

}  // namespace fmcw
