/* libfmcw_cuda -- C ABI of the B200-native range-Doppler-STFT chain.
 *
 * Drop-in boundary for the per-frame loop and the STFT block of the reference
 * (alepnabil/fmcw_radar_processing, radar-etl-pipeline/radar_processing.m, "RP"):
 *   RP:197-261  per-frame chain ('no' branch)      -> fmcw_process_frames / fmcw_run
 *   RP:265      max over chirps for all frames     -> fmcw_frame_out.range_max_abs
 *   RP:270-299  STFT + dB + log-frequency resample -> fmcw_stft / fmcw_run
 *   RP:457-566  same per 100-frame batch ('yes')   -> fmcw_run on a frame sub-range
 *   RP:410-411  range spectrum of one (frame,chirp)-> fmcw_range_spectrum
 *   RP:332-348  fine-grid psd band of the picture  -> fmcw_stft_finegrid
 *   RP:89-179   configuration                      -> fmcw_config (fmcw_configurations, RP:645-672)
 * The reference has no FFI of its own (pure MATLAB); the MEX gateway (mex/) and the
 * Node-API addon (node/) bind exactly these entry points, see INTEGRATION.md.
 *
 * Conventions: plain C, no exceptions, every call returns an fmcw_status.  All buffers
 * are caller-owned; each may be a host or a device pointer (detected with
 * cudaPointerGetAttributes).  Indices are 0-based (MATLAB index - 1).  A handle is
 * not re-entrant (a second concurrent call returns FMCW_ERR_BUSY); distinct handles
 * are independent and may be used from any thread.  There is no CPU fallback.
 */
#ifndef FMCW_CUDA_H
#define FMCW_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FMCW_API __declspec(dllexport)
#else
#define FMCW_API __attribute__((visibility("default")))
#endif

typedef enum fmcw_status {
  FMCW_OK = 0,
  FMCW_ERR_CONFIG = 1,   /* inconsistent or unsupported fmcw_config field */
  FMCW_ERR_POINTER = 2,  /* NULL or mismatched pointer argument */
  FMCW_ERR_CUDA = 3,     /* a CUDA runtime call or kernel failed */
  FMCW_ERR_NCCL = 4,     /* reserved for the collective layer */
  FMCW_ERR_OOM = 5,      /* device or host allocation failed */
  FMCW_ERR_BUSY = 6,     /* another call is in flight on this handle */
  FMCW_ERR_SIZE = 7,     /* a size argument is out of range (e.g. capacity too small) */
  FMCW_ERR_NO_DATA = 8,  /* fewer than window_length slow-time samples (RP:269, RP:534) */
  FMCW_ERR_STATE = 9     /* call order violated (e.g. STFT before frames were processed) */
} fmcw_status;

enum { FMCW_PEAK_STRONGEST = 0, FMCW_PEAK_FIRST = 1 };
enum { FMCW_LAYOUT_TIME_MAJOR = 0,  /* intensity[col][n_freq]: MATLAB's 1024 x ncol column-major */
       FMCW_LAYOUT_FREQ_MAJOR = 1   /* intensity[n_freq][ld_cols]: jsonencode / NumPy row order  */ };

/* Field names follow the reference's fmcw_configurations struct (RP:645-672); the values
 * are what RP:89-179 computes.  struct_size must be sizeof(fmcw_config) of the caller. */
typedef struct fmcw_config {
  uint32_t struct_size;
  uint32_t num_Tx_antennas;             /* RP:102 (unused by the chain) */
  uint32_t num_Rx_antennas;             /* RP:103 */
  uint32_t num_ADC_samples_per_chirp;   /* NTS, RP:109 */
  uint32_t num_chirps_per_frame;        /* PN,  RP:112 */
  uint32_t range_fft_size;              /* RP:118, must be 256 */
  uint32_t Doppler_fft_size;            /* RP:119, power of two <= 64 */
  uint32_t max_num_targets;             /* RP:129, must be 1 */
  uint32_t window_length;               /* RP:178 */
  uint32_t overlap;                     /* RP:179 */
  uint32_t MAX_FREQ_BINS;               /* RP:293, <= 1024 */
  uint32_t rx_select;                   /* 0-based RX processed; the reference uses RX 1 -> 0 (RP:202) */
  uint32_t peak_mode;                   /* FMCW_PEAK_* (f_search_peak is not shipped upstream) */
  uint32_t reserved0;
  double frame_time;                    /* RP:91 */
  double PRT;                           /* RP:97 */
  double Bandwidth;                     /* RP:100 */
  double carrier_frequency;             /* RP:106 */
  double sampling_frequency;            /* RP:115 (unused by the chain) */
  double IF_scale;                      /* RP:121 */
  double range_threshold;               /* RP:123 */
  double Doppler_threshold;             /* RP:124 */
  double min_distance;                  /* RP:126 */
  double max_distance;                  /* RP:127 */
  double lambda;                        /* RP:133 */
  double Hz_to_mps_constant;            /* RP:135 */
  double R_max;                         /* RP:142 */
  double dist_per_bin;                  /* RP:147 */
  double fD_max;                        /* RP:152 */
  double fD_per_bin;                    /* RP:153 */
  double kaiser_beta;                   /* RP:276 literal 3 */
  double adc_scale;                     /* parser normalisation, 4095 (f_parse_data2 shim) */
} fmcw_config;

typedef struct fmcw_handle fmcw_handle;

/* Per-frame outputs (RP:197-261).  Any pointer may be NULL to skip that output. */
typedef struct fmcw_frame_out {
  float*   range_max_abs;  /* [n_frames][range_fft_size]  abs(max(range_fft,[],2)), RP:210/265 */
  int32_t* detected;       /* [n_frames] 1 if f_search_peak returned a target (RP:213) */
  int32_t* range_bin;      /* [n_frames] tgt_range_idx(1)-1, or -1 */
  float*   range_mag;      /* [n_frames] tgt_range_mag(1), or 0 (RP:245) */
  int32_t* doppler_bin;    /* [n_frames] tgt_doppler_idx(1)-1 in the fftshifted row (RP:233-238) */
  float*   doppler_row;    /* [n_frames][Doppler_fft_size][2] fftshifted complex row, RP:219 */
  float*   slow_time_mag;  /* [n_frames][PN] abs of the stored range-FFT row at the selected bin, RP:259/270 */
} fmcw_frame_out;

/* STFT outputs (RP:270-299). */
typedef struct fmcw_stft_out {
  float*   intensity;      /* interp_intensity, layout per `layout` */
  uint64_t capacity_cols;  /* columns the intensity buffer can hold */
  uint64_t ld_cols;        /* row stride for FMCW_LAYOUT_FREQ_MAJOR (0 = capacity_cols) */
  uint32_t layout;         /* FMCW_LAYOUT_* */
  uint32_t reserved0;
} fmcw_stft_out;

/* Scalars describing a finished run (host memory, filled by fmcw_get_info). */
typedef struct fmcw_run_info {
  uint64_t n_frames;       /* frames processed by the last fmcw_process_frames / fmcw_run */
  uint64_t n_detected;     /* frames with a target */
  uint64_t L_local;        /* slow-time samples held by this handle (n_detected * PN) */
  uint64_t L_total;        /* length of the whole (possibly sharded) slow-time signal */
  uint64_t sample_offset;  /* global index of this handle's first sample */
  uint64_t nfft;           /* 2^nextpow2(L_total), RP:273 */
  uint64_t ncol_total;     /* spectrogram columns of the whole signal */
  uint64_t col_begin;      /* first global column written by this handle */
  uint64_t ncol_local;     /* columns written by this handle */
  uint32_t n_dtft_bins;    /* distinct fine-grid bins evaluated per column */
  uint32_t n_refined;      /* columns that needed the exhaustive max search */
  double   pmax_raw;       /* max over the one-sided fine grid of c_j*|S|^2 (un-normalised) */
} fmcw_run_info;

FMCW_API const char* fmcw_version(void);
FMCW_API const char* fmcw_status_string(fmcw_status s);

/* calib_data: the parser's calibration row vector [I_rx1 Q_rx1 I_rx2 Q_rx2 ...] in normalised
 * units, length 2*num_Rx_antennas*N_cal with N_cal a multiple of NTS (RP:167-174).  Host pointer. */
FMCW_API fmcw_status fmcw_create(const fmcw_config* cfg, const double* calib_data, uint64_t calib_len,
                                 int device, fmcw_handle** out);
FMCW_API void        fmcw_destroy(fmcw_handle* h);
FMCW_API const char* fmcw_last_error(const fmcw_handle* h);
FMCW_API void*       fmcw_get_stream(fmcw_handle* h);      /* cudaStream_t all work is queued on */
FMCW_API fmcw_status fmcw_synchronize(fmcw_handle* h);

/* Options.  FMCW_OPT_ASYNC_HOST = 1: fmcw_process_frames / fmcw_run with (pinned) host buffers return as soon as the
 * copies and kernels are queued, so that two handles can overlap one recording's D2H with the next one's H2D; the
 * intensity copy then covers the upper bound of the column count and the results are valid after fmcw_synchronize
 * (sizes from fmcw_get_info).
 *
 * FMCW_OPT_STFT_PRECISION = 2: arithmetic of the STFT main kernel (RP:276-299).
 *   0 (default): tcgen05 tensor cores, operands split hi + lo in TF32 (2^-22 of the mean-removed column), float64 column
 *      mean through a float64-tabulated window response.  Against the double-precision reference: 1e-3 dB for bins above
 *      -60 dB and 1e-4 relative (of P/max) down to -140 dB on every scene; below that the split is the floor (measured
 *      4e-4 in (-180,-140] dB, 2e-3 in (-220,-180] dB on a scene whose columns carry strong 0-200 Hz content; 5e-5 down to
 *      -260 dB when the slow-time signal is a level plus noise).
 *   1: float64 CUDA-core kernel, any window: 1e-4 relative at every finite level (measured 5e-6); about 13x the time of
 *      the default kernel (3.4 ms instead of 0.27 ms per 5,000-frame recording on a B200).
 * tests/helpers.py::assert_spectrogram_contract asserts exactly these bounds.
 *
 * FMCW_OPT_RUN_GRAPH = 3: 1 = fmcw_run calls whose buffers are all device memory are recorded as ONE CUDA graph per set of
 *   buffers (input, outputs, info target) the third time the set is seen and replayed from then on.  For fleets of small
 *   recordings (BASELINE C5): a 500-frame recording is about 12 kernel launches and the GPU's launch rate, not its work, bounds
 *   the pass.  Results are identical; fmcw_get_timings reports 0 for replayed runs; a handle keeps at most 256 buffer sets and
 *   drops every recorded graph when one of its scratch buffers has to grow.
 *
 * FMCW_OPT_STFT_TILES_PER_CTA = 4: n > 1 = the persistent tensor-core STFT kernel is launched with one CTA per n tiles (128 columns)
 *   of the output buffer's capacity instead of one per SM when that is fewer: the kernels of several small recordings on
 *   different handles then run side by side (a fleet), at the price of one recording's own latency.  0 / 1 (default): one per SM. */
enum { FMCW_OPT_ASYNC_HOST = 1, FMCW_OPT_STFT_PRECISION = 2, FMCW_OPT_RUN_GRAPH = 3, FMCW_OPT_STFT_TILES_PER_CTA = 4 };
FMCW_API fmcw_status fmcw_set_option(fmcw_handle* h, int option, int64_t value);
FMCW_API fmcw_status fmcw_get_info(fmcw_handle* h, fmcw_run_info* info);   /* synchronises */

/* The scalars of a run without a host round trip: while a target is set, every fmcw_run on the handle ends by writing this
 * struct to `device_dst` (device memory, stream-ordered after the run's kernels).  A fleet driver gives every recording its own
 * slot and reads all of them with one copy after the pass.  status: 0, or the device-side failure code (-2 more DTFT bins than
 * the plan tables hold, -3 chunk too large, -4 capacity_cols too small, -5 too many columns for the exhaustive max search).
 * NULL disables. */
typedef struct fmcw_device_info {
  fmcw_run_info info;
  int32_t status;
  int32_t reserved0;
} fmcw_device_info;
FMCW_API fmcw_status fmcw_set_info_target(fmcw_handle* h, fmcw_device_info* device_dst);

/* Device time in ms of the stages of the last fmcw_run / fmcw_process_frames (CUDA events on the
 * handle's stream; synchronises): ms[0] frame chain, ms[1] compaction, ms[2] STFT plan + global max,
 * ms[3] STFT main kernel.  Stages that did not run report 0. */
FMCW_API fmcw_status fmcw_get_timings(fmcw_handle* h, float* ms);

/* iq: int16 [n_frames][num_Rx_antennas][PN][NTS][2] ADC codes (I,Q interleaved), host or device. */
FMCW_API fmcw_status fmcw_process_frames(fmcw_handle* h, const int16_t* iq, uint64_t n_frames,
                                         const fmcw_frame_out* out);

/* Fused single-GPU path: frames -> compaction -> STFT of the detected slow-time signal.
 * Asynchronous when every buffer is device memory; call fmcw_get_info for the sizes. */
FMCW_API fmcw_status fmcw_run(fmcw_handle* h, const int16_t* iq, uint64_t n_frames,
                              const fmcw_frame_out* fout, const fmcw_stft_out* sout);

/* STFT (RP:270-299 / RP:538-566) of the slow-time signal gathered by the last fmcw_process_frames on this handle
 * (the float64 magnitudes of the detected frames): the second half of fmcw_run. */
FMCW_API fmcw_status fmcw_stft_frames(fmcw_handle* h, const fmcw_stft_out* sout);

/* STFT of an arbitrary non-negative sequence x[L] (host or device): RP:270-299 on its own. */
FMCW_API fmcw_status fmcw_stft(fmcw_handle* h, const float* x, uint64_t L, const fmcw_stft_out* sout);

/* Sharded path (one handle per GPU; the collectives between the steps belong to the caller):
 *   fmcw_process_frames -> fmcw_get_info (n_detected) -> all-gather counts ->
 *   fmcw_get_slow_time / fmcw_set_halo (neighbour exchange of window_length-1 samples) ->
 *   fmcw_stft_local_max -> all-reduce(max) -> fmcw_stft_sharded. */
FMCW_API fmcw_status fmcw_get_slow_time(fmcw_handle* h, double* dst, uint64_t first, uint64_t count);
/* Streaming (recordings whose frames or spectrogram do not fit the GPU, BASELINE configs[3]): the caller keeps the
 * slow-time magnitudes (8 B per chirp of a detected frame) of the whole recording, collected chunk by chunk with
 * fmcw_process_frames + fmcw_get_slow_time, and hands any piece of them back: x = L_local samples followed by n_halo
 * (< window_length) samples that follow the piece.  L_local must be a multiple of num_chirps_per_frame.  The piece then
 * behaves like a shard: fmcw_stft_local_max / fmcw_stft_sharded with the piece's global sample offset.  Host or device x. */
FMCW_API fmcw_status fmcw_load_slow_time(fmcw_handle* h, const double* x, uint64_t L_local, uint64_t n_halo);
FMCW_API fmcw_status fmcw_set_halo(fmcw_handle* h, const double* src, uint64_t count);
FMCW_API fmcw_status fmcw_stft_local_max(fmcw_handle* h, uint64_t L_total, uint64_t sample_offset,
                                         double* pmax_raw_local);
FMCW_API fmcw_status fmcw_stft_sharded(fmcw_handle* h, uint64_t L_total, uint64_t sample_offset,
                                       double pmax_raw_global, const fmcw_stft_out* sout);

/* Asynchronous form of the sharded path: every hand-off stays in device memory, the host only enqueues.
 *   fmcw_process_frames (device buffers) -> fmcw_shard_pack(msg) -> all-gather(msg) -> fmcw_shard_plan(gathered,
 *   world, rank, local_max) -> all-reduce(max, local_max) -> fmcw_shard_stft(global_max, out).
 * msg: double[1 + window_length-1] = {L_local, first window_length-1 samples} (the slow-time signal is float64);
 * gathered: the world messages in rank order.  All pointers are device memory; work is queued on the
 * handle's stream (order it against the collective's stream with events). */
FMCW_API fmcw_status fmcw_shard_pack(fmcw_handle* h, double* msg_dev);
FMCW_API fmcw_status fmcw_shard_plan(fmcw_handle* h, const double* gathered_dev, uint32_t world, uint32_t rank,
                                     double* local_max_dev);
FMCW_API fmcw_status fmcw_shard_stft(fmcw_handle* h, const double* global_max_dev, const fmcw_stft_out* sout);

/* Peer-memory form of the same path (NVLink/NVSwitch, no collective call between the steps): every rank owns a
 * zero-initialised "mailbox" of fmcw_mailbox_bytes() bytes of device memory that is mapped on all ranks (CUDA IPC /
 * VMM, e.g. torch symmetric memory).  mailboxes[r] is rank r's mailbox as addressable from THIS process (a host
 * array of world device pointers, mailboxes[rank] being the own one).  step is the same strictly increasing number
 * (>= 1) on all ranks for one pass.
 *   fmcw_process_frames -> fmcw_mailbox_post_heads: stores {L_local, first window_length-1 samples} into every
 *     mailbox and raises this rank's flag there;
 *   fmcw_mailbox_plan: waits (on the device) for all world headers, plans, searches the local maximum and posts it;
 *   fmcw_mailbox_stft: waits for all world maxima, takes their maximum and runs the STFT of the own columns.
 * Replaces nothing in the reference (single-process MATLAB); a rank that never posts makes the waiting kernels give
 * up after 20 s and the pass fail with FMCW_ERR_STATE at the next synchronisation. world <= 64. */
FMCW_API uint64_t fmcw_mailbox_bytes(void);
FMCW_API fmcw_status fmcw_mailbox_post_heads(fmcw_handle* h, void* const* mailboxes, uint32_t world, uint32_t rank,
                                             uint64_t step);
FMCW_API fmcw_status fmcw_mailbox_plan(fmcw_handle* h, void* const* mailboxes, uint32_t world, uint32_t rank,
                                       uint64_t step);
FMCW_API fmcw_status fmcw_mailbox_stft(fmcw_handle* h, void* const* mailboxes, uint32_t world, uint32_t rank,
                                       uint64_t step, const fmcw_stft_out* sout);

/* Host-side axes in float64: T (RP:276) for columns [col_begin, col_begin+ncol) and
 * log_freq_bins (RP:293-296).  Either pointer may be NULL. */
FMCW_API fmcw_status fmcw_stft_axes(const fmcw_config* cfg, uint64_t L_total, uint64_t col_begin, uint64_t ncol,
                                    double* time, double* frequency, uint64_t* nfft, uint64_t* ncol_total);

/* Fine-grid PSD band for the spectrogram picture (RP:332-348: surf(T, F, psd), ylim [0 150], clim [-40 0]) of the signal of
 * the last fmcw_run / fmcw_stft_frames / fmcw_stft call: psd = 20*log10(P / max(P)) (RP:283) at the one-sided fine-grid bins
 * F_j = j*fs/nfft inside [f_lo_hz, f_hi_hz], every `*bin_step`-th bin so that at most max_rows rows come back (a 5,000-frame
 * recording has 62,915 fine-grid bins below 150 Hz; the literal psd matrix of RP:283 is 671 GB).  psd: [ncol][n_rows] floats,
 * time-major = MATLAB's column-major psd(rows, cols) restricted to the selected rows; host or device memory with room for
 * capacity_cols columns of max_rows rows.  first_bin / bin_step / n_rows / ncol describe the rows and columns written
 * (F = (first_bin + i*bin_step)*fs/nfft; the time axis is fmcw_stft_axes').  Single-GPU runs only (local columns of a
 * sharded run are returned as they are, normalised by the global maximum). */
FMCW_API fmcw_status fmcw_stft_finegrid(fmcw_handle* h, double f_lo_hz, double f_hi_hz, uint32_t max_rows, float* psd,
                                        uint64_t capacity_cols, uint64_t* first_bin, uint64_t* bin_step, uint64_t* n_rows,
                                        uint64_t* ncol);

/* Full range-Doppler map of one frame in dB (north_star "|X| to dB"): RP:216-219 applied to EVERY range row instead of the
 * detected one only -- per row: mean over chirps removed, 2*chebwin(PN) window, Doppler_fft_size-point FFT over the first
 * min(PN, ND) chirps (fft(.,ND,2) truncates), fftshift -- then 20*log10(|.|).  The reference fills Rx_spectrum(:,:,1) with the
 * target row only and never reads it (RP:221); this is the map its name promises.  out_db: [range_fft_size][Doppler_fft_size]
 * floats (row-major), host or device. */
FMCW_API fmcw_status fmcw_range_doppler_map(fmcw_handle* h, const int16_t* iq, uint64_t n_frames, uint64_t frame, float* out_db);

/* Range spectrum abs(range_fft(:, chirp)) of one frame (RP:410-411). out: [range_fft_size] floats. */
FMCW_API fmcw_status fmcw_range_spectrum(fmcw_handle* h, const int16_t* iq, uint64_t n_frames,
                                         uint64_t frame, uint32_t chirp, float* out);

/* Host-only helpers for the reference's JSON payloads (RP:307-318, 358-367, 386-396, 576-587): append the jsonencode
 * text of a numeric matrix (element (r,c) at a[r*row_stride + c*col_stride], strides in elements) to a file.  Nested
 * row-major like jsonencode, NaN/Inf -> null, shortest round-trip digits; flatten_vectors != 0 writes 1 x N and
 * N x 1 matrices as flat arrays.  Rows are formatted by all host threads. */
FMCW_API fmcw_status fmcw_json_append_f32(const char* path, const float* a, uint64_t rows, uint64_t cols,
                                          int64_t row_stride, int64_t col_stride, int flatten_vectors);
FMCW_API fmcw_status fmcw_json_append_f64(const char* path, const double* a, uint64_t rows, uint64_t cols,
                                          int64_t row_stride, int64_t col_stride, int flatten_vectors);

/* Synthetic scene generator: same bits as fmcw_radar_processing_b200/synth.py.
 * tables: float64 [n_frames][n_scat][4] (A, cycles/sample, cycles/chirp, phase cycles), host or device.
 * iq_out: device or host int16 [n_frames][n_rx][PN][NTS][2]. */
FMCW_API fmcw_status fmcw_synth_frames(fmcw_handle* h, const double* tables, uint32_t n_scat, uint64_t seed,
                                       uint64_t frame0, uint64_t n_frames, double sigma, double dc,
                                       double rx_step, int16_t* iq_out);

#ifdef __cplusplus
}
#endif
#endif /* FMCW_CUDA_H */
