// Synthetic scene generator; device twin of fmcw_radar_processing_b200/synth.py.
// Every sample is a pure function of (seed, global frame, rx, chirp, sample).  Multiplies and adds
// are issued separately (__dmul_rn / __dadd_rn, no FMA contraction) so the float64 value that is
// rounded to an ADC code differs from NumPy's only through the last ulp of sin/cos.
#include "fmcw_internal.cuh"

namespace fmcw {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double ih4(unsigned long long h) {
  const unsigned long long u = (h & 0xFFFFull) + ((h >> 16) & 0xFFFFull) + ((h >> 32) & 0xFFFFull) + (h >> 48);
  return __dmul_rn(__dadd_rn((double)u, -131070.0), 1.0 / 37837.226637012174);
}

__global__ void __launch_bounds__(256) synth_kernel(const double* __restrict__ tables, uint32_t n_scat,
                                                    unsigned long long seedmix, uint64_t frame0, uint64_t n_frames,
                                                    uint32_t n_rx, uint32_t PN, uint32_t NTS, double sigma, double dc,
                                                    double rx_step, uint32_t* __restrict__ out) {
  const unsigned long long per_frame = (unsigned long long)n_rx * PN * NTS;
  const unsigned long long total = per_frame * n_frames;
  const double two_pi = 6.283185307179586;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long k = i / per_frame;
    const unsigned long long lin = i - k * per_frame;
    const uint32_t n = (uint32_t)(lin % NTS);
    const uint32_t m = (uint32_t)((lin / NTS) % PN);
    const uint32_t r = (uint32_t)(lin / ((unsigned long long)NTS * PN));
    double sI = dc, sQ = dc;
    for (uint32_t s = 0; s < n_scat; ++s) {
      const double* tb = tables + (k * n_scat + s) * 4;
      const double A = tb[0];
      if (A == 0.0) continue;
      double p = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(tb[1], (double)n), __dmul_rn(tb[2], (double)m)), tb[3]),
                           __dmul_rn((double)r, rx_step));
      p = __dadd_rn(p, -floor(p));
      const double ang = __dmul_rn(two_pi, p);
      double sn, cs;
      sincos(ang, &sn, &cs);
      sI = __dadd_rn(sI, __dmul_rn(A, cs));
      sQ = __dadd_rn(sQ, __dmul_rn(A, sn));
    }
    const unsigned long long idx = (frame0 + k) * per_frame + lin;
    const unsigned long long kI = seedmix + idx * 2ull;
    sI = __dadd_rn(sI, __dmul_rn(sigma, ih4(splitmix64(kI))));
    sQ = __dadd_rn(sQ, __dmul_rn(sigma, ih4(splitmix64(kI + 1ull))));
    double cI = floor(__dadd_rn(sI, 0.5)), cQ = floor(__dadd_rn(sQ, 0.5));
    cI = fmin(fmax(cI, 0.0), 4095.0);
    cQ = fmin(fmax(cQ, 0.0), 4095.0);
    out[i] = ((uint32_t)(int)cI & 0xffffu) | ((uint32_t)(int)cQ << 16);
  }
}

cudaError_t launch_synth(const double* tables, uint32_t n_scat, uint64_t seed, uint64_t frame0, uint64_t n_frames,
                         uint32_t n_rx, uint32_t PN, uint32_t NTS, double sigma, double dc, double rx_step,
                         int16_t* out, cudaStream_t st) {
  if (n_frames == 0) return cudaSuccess;
  const unsigned long long seedmix = (unsigned long long)seed * 0x9E3779B97F4A7C15ull;
  const unsigned long long total = (unsigned long long)n_frames * n_rx * PN * NTS;
  unsigned long long blocks = (total + 255) / 256;
  if (blocks > 148ull * 32) blocks = 148ull * 32;
  synth_kernel<<<(unsigned)blocks, 256, 0, st>>>(tables, n_scat, seedmix, frame0, n_frames, n_rx, PN, NTS, sigma, dc,
                                                 rx_step, reinterpret_cast<uint32_t*>(out));
  return cudaGetLastError();
}

}  // namespace fmcw
