// Internal declarations shared by the kernels and the C ABI of libfmcw_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fmcw {

typedef double sig_t;           // slow-time magnitude signal: float64 from the single-bin DFT to the STFT operands

constexpr int NR = 256;           // range_fft_size (RP:118); the radix-16 x radix-16 core is built for it
constexpr int MAX_ND = 64;        // Doppler_fft_size upper bound
constexpr int MAX_NQ = 1024;      // MAX_FREQ_BINS upper bound (RP:293)
constexpr int MAX_CHUNKS = 32;    // query chunks per spectrogram column
constexpr int CHAIN_THREADS = 256;
constexpr int CHAIN_WARPS = CHAIN_THREADS / 32;
constexpr int XCH_STRIDE = 17;    // float2 stride between k1 rows of the in-warp transpose (bank-conflict free)
constexpr int XCH_CHIRP = 16 * XCH_STRIDE;   // 272 float2 per chirp

// ---- per-frame chain (RP:199-260) -------------------------------------------------------------
struct ChainParams {
  const uint32_t* iq;        // int16 (I,Q) pairs, [frame][rx][chirp][sample]
  uint64_t n_frames;
  uint32_t NTS, PN, n_rx, rx_sel, ND;
  uint32_t nts_fft;          // min(NTS, NR): samples that enter the FFT (fft(x,256,1) truncates, RP:205)
  const float4* win_tab;     // [nts_fft] {gw, h_re, h_im, 0}: xw = gw*(NTS*code - sum) - h
  const double* win_tab_d;   // [nts_fft][3] the same table in float64 (slow-time row)
  const double2* tw_d;       // [256] W_256^k in float64
  const double2* hfft_d;     // [256] FFT of the calibration term h[n] (float64), subtracted from the slow-time row
  const float2* tw_pair;     // [16][16] W_256^(s*k1)
  const float*  tw_re;       // [272] skewed W_256^k table (index k + k/16)
  const float*  tw_im;
  const float2* dop_tw;      // [ND] W_ND^k
  const float*  dop_win;     // [min(PN,ND)] first taps of 2*chebwin(PN) (RP:139, 219)
  const double* dop_win_d;   // the same taps in float64 (range-Doppler map export)
  int32_t bin_lo, bin_hi;    // range gate, 0-based inclusive (RP:126-127 through f_search_peak)
  float range_thr, dop_thr;
  int32_t peak_mode;
  // outputs, device pointers, any may be null
  float* range_max_abs; int32_t* detected; int32_t* range_bin; float* range_mag;
  int32_t* doppler_bin; float2* doppler_row; float* slow_mag;
  sig_t* slow64;             // [n_frames][PN] float64 magnitudes (always written; feeds the compaction)
  // optional: range spectrum of one (frame, chirp) (RP:410-411)
  float* spec_out; uint64_t spec_frame; uint32_t spec_chirp;
};

size_t chain_smem_bytes(uint32_t PN);
cudaError_t launch_frame_chain(const ChainParams& p, cudaStream_t st);
// Process-wide settings, read once (C++11 thread-safe static initialisation): handles may be used from any thread.
int device_sm_count();                      // SMs of the current device (148 on B200)
int env_int(const char* name, int dflt);    // atoi(getenv(name)) or dflt
// one warp per frame, two chirps per lane (frame_chain_warp.cu); launch_frame_chain dispatches to it when supported
bool chain_warp_supported(const ChainParams& p);
cudaError_t launch_frame_chain_warp(const ChainParams& p, cudaStream_t st);

// full range-Doppler dB map of one frame (frame_chain.cu)
cudaError_t launch_range_doppler_map(const ChainParams& p, uint64_t frame, float* out_db, cudaStream_t st);

// ---- compaction of the detected frames' slow-time rows (RP:257-260) ------------------------------
struct CompactParams {
  const int32_t* detected; uint64_t n_frames; uint32_t PN;
  const sig_t* slow_mag;      // [n_frames][PN]
  sig_t* xc;                  // compacted magnitudes
  uint32_t* det_list;         // [n_frames] frame index of the k-th detection
  unsigned long long* n_det;  // device scalar
  uint32_t* counts;           // [ceil(n/1024) * 33] scratch of the two-launch form (long recordings), may be null
};
cudaError_t launch_compact(const CompactParams& p, cudaStream_t st);

// ---- STFT plan (device resident; RP:273, 293-299 restated) ----------------------------------------
struct StftPlan {
  unsigned long long L_total, nfft, ncol_total, col_begin, col_end, sample_offset, L_avail, L_local;
  int log2nfft, nb, nq, n_chunks, valid;
  unsigned int n_hard, n_refined, task_counter, ticket_r, ticket_h;
  int spec_state;               // 1: planned ahead for the length in L_total (not yet confirmed), 2: confirmed, 0: none
  float lb_max;                 // max_t max(S0^2, 2|S(w1)|^2) (lower bound of the global max)
  double pmax_raw;              // final global max of c_j |S|^2
  int chunk_q0[MAX_CHUNKS + 1]; // query range of each chunk (multiples of 32 except the last end)
  int chunk_p0[MAX_CHUNKS + 1]; // first bin position each chunk needs
};

struct StftTables {             // device arrays owned by the handle
  StftPlan* plan;
  int* bins;                    // [nb_max] fine-grid bin of position p
  float* kcb;                   // [nb_max] K*log2(c_p), c_p = 1 for DC/Nyquist else 2
  int* qpos;                    // [nq] position p_q of the lower bracket bin of query q
  float* aq;                    // [nq] interpolation weight
  int* qend;                    // [nb_max+1] queries [qend[p-1], qend[p]) complete when position p is known
  float* coef;                  // [nb_max][2*half] cos((m+d)w), sin((m+d)w)
  float* win;                   // [win] kaiser window (host computed, float64 -> float32)
  double* win_d;                // [win] the same window in float64 (DC response tables, float64 STFT kernel)
  float* wdc;                   // [nb_max] window DC response per bin position (generic-window kernel)
  unsigned int* hard_list;      // columns whose max needs the exhaustive search
  unsigned int hard_cap;
  float* col_ub;                // [local columns] trivial upper bound 2*(sum|y|)^2 of each column's maximum
  float* tcB;                   // tensor-core path: per chunk of 64 queries the B operands C|S in the UMMA smem layout
  int nb_max;
};

struct StftGeom {
  uint32_t win, hop, nq;
  double fs;
  float rho;   // rigorous bound of max_{w >= pi/(win-1)} |W(w)| / W(0) of the STFT window (host, float64)
  uint32_t tiles_per_cta;   // FMCW_OPT_STFT_TILES_PER_CTA (0 / 1: one CTA per SM): grid of the persistent STFT kernel for small recordings
};



cudaError_t launch_shard_pack(const sig_t* xc, const unsigned long long* d_ndet, uint32_t PN, uint32_t win, double* msg,
                              cudaStream_t st);
cudaError_t launch_stft_set_max_dev(const StftTables& t, const double* src, cudaStream_t st);
cudaError_t launch_stft_plan(const StftTables& t, const StftGeom& g, const unsigned long long* d_ndet, uint32_t PN,
                             unsigned long long L_total_host, unsigned long long sample_offset,
                             unsigned long long L_local_host, unsigned long long L_avail_host, int n_chunks,
                             cudaStream_t st, const double* gathered = nullptr, uint32_t world = 0, uint32_t rank = 0,
                             sig_t* xc = nullptr, int spec_mode = 0, const unsigned long long* wait_flags = nullptr,
                             const unsigned long long* wait_step = nullptr);
cudaError_t launch_stft_max(const StftTables& t, const StftGeom& g, const sig_t* x, cudaStream_t st,
                            double* export_dst = nullptr);
cudaError_t launch_stft_set_max(const StftTables& t, double pmax_raw, cudaStream_t st);
cudaError_t launch_stft_main(const StftTables& t, const StftGeom& g, const sig_t* x, float* out,
                             unsigned long long capacity_cols, unsigned long long ld_cols, int layout,
                             int* d_err, cudaStream_t st, const double* gmax_dev = nullptr, int precise = 0);

cudaError_t launch_stft_finegrid(const StftTables& t, const StftGeom& g, const sig_t* x, float* psd, unsigned long long first_bin,
                                 unsigned long long bin_step, unsigned n_rows, unsigned long long capacity_cols, int* d_err,
                                 cudaStream_t st);

// tensor-core (tcgen05) STFT main kernel, window_length = 20 (stft_tc.cu)
size_t stft_tc_table_bytes(int nb_max);
cudaError_t launch_stft_tc_prepare(const StftTables& t, const StftGeom& g, float* tcB, int nb_max, cudaStream_t st,
                                   int spec_mode = 0);
cudaError_t launch_stft_tc_main(const StftTables& t, const StftGeom& g, const sig_t* x, float* out, const float* tcB,
                                unsigned long long capacity_cols, unsigned long long ld_cols,
                                int layout, int* d_err, cudaStream_t st, const double* gmax_dev = nullptr);
int stft_variant();            // FMCW_STFT_VARIANT: -1 (default) tensor cores, 0..4 CUDA-core variants

cudaError_t launch_f32_to_sig(const float* src, sig_t* dst, unsigned long long n, cudaStream_t st);

// ---- peer-memory mailboxes of the sharded path (mailbox.cu) ------------------------------------------------
constexpr uint32_t MAILBOX_MAX_WORLD = 64;
constexpr uint32_t MAILBOX_MAX_WIN = 400;
constexpr unsigned long long MAILBOX_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
constexpr size_t MAILBOX_BYTES = (size_t)MAILBOX_MAX_WORLD * (8 + 8 + 8) + (size_t)MAILBOX_MAX_WORLD * MAILBOX_MAX_WIN * 8;
struct MailboxSet { void* ptr[MAILBOX_MAX_WORLD]; };   // the mailbox of every rank as mapped on this rank
__host__ __device__ inline unsigned long long* mailbox_flag_heads(void* m) { return reinterpret_cast<unsigned long long*>(m); }
__host__ __device__ inline unsigned long long* mailbox_flag_max(void* m) { return reinterpret_cast<unsigned long long*>(m) + MAILBOX_MAX_WORLD; }
__host__ __device__ inline double* mailbox_max(void* m) { return reinterpret_cast<double*>(m) + 2 * MAILBOX_MAX_WORLD; }
__host__ __device__ inline double* mailbox_heads(void* m) { return reinterpret_cast<double*>(m) + 3 * MAILBOX_MAX_WORLD; }
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long mailbox_ld_flag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long mailbox_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// true when flags[i] >= step before the timeout (threads i >= n return true at once)
__device__ __forceinline__ bool mailbox_wait(const unsigned long long* flags, uint32_t i, uint32_t n, unsigned long long step) {
  if (i >= n) return true;
  const unsigned long long t0 = mailbox_global_ns();
  while (mailbox_ld_flag(flags + i) < step) {
    __nanosleep(64);
    if (mailbox_global_ns() - t0 > MAILBOX_TIMEOUT_NS) return false;
  }
  return true;
}
#endif
cudaError_t launch_mailbox_post_heads(const sig_t* xc, const unsigned long long* d_ndet, uint32_t PN, uint32_t win,
                                      const MailboxSet& mb, uint32_t world, uint32_t rank, unsigned long long* d_step,
                                      cudaStream_t st);
cudaError_t launch_mailbox_post_max(const double* local_max, const MailboxSet& mb, uint32_t world, uint32_t rank,
                                    const unsigned long long* d_step, cudaStream_t st);
cudaError_t launch_mailbox_collect_max(void* own, uint32_t world, const unsigned long long* d_step, double* gmax, int* d_err,
                                       cudaStream_t st);

// ---- synthetic scene generator ------------------------------------------------------------------
cudaError_t launch_synth(const double* tables, uint32_t n_scat, uint64_t seed, uint64_t frame0, uint64_t n_frames,
                         uint32_t n_rx, uint32_t PN, uint32_t NTS, double sigma, double dc, double rx_step,
                         int16_t* out, cudaStream_t st);

}  // namespace fmcw
