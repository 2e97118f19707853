// C ABI of libfmcw_cuda (include/fmcw_cuda.h): handle, host-side tables, buffer plumbing.
// No C++ type or exception crosses the boundary; every entry point returns an fmcw_status.
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/fmcw_cuda.h"
#include "fmcw_internal.cuh"

using namespace fmcw;

namespace {

// bumped by every (re)allocation of a scratch buffer: recorded run graphs hold the old addresses and are dropped
std::atomic<unsigned long long> g_alloc_epoch{0};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- window functions in float64 (MathWorks blackman / chebwin / kaiser restated) ----------------
std::vector<double> blackman_sym(int N) {
  std::vector<double> w(N, 1.0);
  if (N == 1) return w;
  for (int n = 0; n < N; ++n) {
    const double t = 2.0 * M_PI * n / (N - 1);
    w[n] = 0.42 - 0.5 * std::cos(t) + 0.08 * std::cos(2.0 * t);
  }
  return w;
}

std::vector<double> chebwin_sym(int M, double at) {
  std::vector<double> w(M, 1.0);
  if (M == 1) return w;
  const double order = M - 1.0;
  const double beta = std::cosh(std::acosh(std::pow(10.0, std::fabs(at) / 20.0)) / order);
  std::vector<std::complex<double>> p(M);
  for (int k = 0; k < M; ++k) {
    const double x = beta * std::cos(M_PI * k / M);
    double v;
    if (x > 1) v = std::cosh(order * std::acosh(x));
    else if (x < -1) v = (2 * (M % 2) - 1) * std::cosh(order * std::acosh(-x));
    else v = std::cos(order * std::acos(x));
    p[k] = v;
    if (M % 2 == 0) p[k] *= std::polar(1.0, M_PI / M * k);
  }
  std::vector<double> W(M);
  for (int n = 0; n < M; ++n) {
    long double acc = 0.0L;
    for (int k = 0; k < M; ++k) {
      const long long r = ((long long)n * k) % M;
      const double ang = -2.0 * M_PI * (double)r / M;
      acc += (long double)(p[k].real() * std::cos(ang) - p[k].imag() * std::sin(ang));
    }
    W[n] = (double)acc;
  }
  if (M % 2) {
    const int n = (M + 1) / 2;
    int o = 0;
    for (int i = n - 1; i >= 1; --i) w[o++] = W[i];
    for (int i = 0; i < n; ++i) w[o++] = W[i];
  } else {
    const int n = M / 2 + 1;
    int o = 0;
    for (int i = n - 1; i >= 1; --i) w[o++] = W[i];
    for (int i = 1; i < n; ++i) w[o++] = W[i];
  }
  double mx = w[0];
  for (double v : w) mx = v > mx ? v : mx;
  for (double& v : w) v /= mx;
  return w;
}

double bessel_i0(double x) {
  const double q = x * x / 4.0;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 500; ++k) {
    term *= q / ((double)k * k);
    sum += term;
    if (term < 1e-18 * sum) break;
  }
  return sum;
}

std::vector<double> kaiser_sym(int N, double beta) {
  std::vector<double> w(N, 1.0);
  if (N == 1) return w;
  const double alpha = (N - 1) / 2.0, d = bessel_i0(beta);
  for (int n = 0; n < N; ++n) {
    const double r = (n - alpha) / alpha;
    w[n] = bessel_i0(beta * std::sqrt(std::fmax(0.0, 1.0 - r * r))) / d;
  }
  return w;
}

int nextpow2_u64(unsigned long long L) {
  int lg = 0;
  while ((1ull << lg) < L) ++lg;
  return lg;
}

}  // namespace

struct fmcw_handle {
  fmcw_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;            // look-ahead STFT plan, concurrent with the frame chain
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool lookahead = false;     // a look-ahead plan is in flight on the side stream (ev_join)
  uint32_t mb_world = 0, mb_rank = 0;   // mailbox path: layout of the last pass (assumed again by the look-ahead plan)
  // the five small launches between compaction and the main STFT kernel (plan confirm, operand prepare, colstat,
  // refine, hard) replayed as one CUDA graph: their cost is launch latency, not work
  cudaGraphExec_t mx_exec = nullptr;
  StftTables mx_key_tables{};
  const void* mx_key_xc = nullptr;
  int mx_eligible_calls = 0;
  // mailbox path: the device-side step counter and the whole exchange pass (post heads, plan confirm, operand prepare,
  // colstat, refine, hard, post max, collect max) as one CUDA graph
  DevBuf mb_step;
  unsigned long long mb_step_host = 0, mb_seed = 0;
  cudaGraphExec_t mbx_exec = nullptr;
  struct MbxKey { StftTables t; const void* xc; void* ptr[MAILBOX_MAX_WORLD]; uint32_t world, rank; } mbx_key{};
  int mbx_eligible_calls = 0;
  bool mbx_use_graph = false, mbx_collected = false;
  std::atomic_flag busy = ATOMIC_FLAG_INIT;
  std::string err;
  // chain tables
  DevBuf win_tab, win_tab_d, tw_d, hfft_d, dop_win_d, tw_pair, tw_re, tw_im, dop_tw, dop_win;
  int bin_lo = 0, bin_hi = -1;
  // STFT tables
  StftTables st{};
  StftGeom geom{};
  DevBuf plan, bins, kcb, wdc, qpos, aq, qend, coef, swin, swin_d, hard, derr, gmax;
  // scratch
  DevBuf tcb, colub;
  DevBuf iq_stage, o_rmax, o_det, o_rbin, o_rmag, o_dbin, o_drow, o_slow, o_slow64, f32_stage, xc, det_list, ndet, inten, synth_tab;
  // state
  uint64_t n_frames = 0;
  bool frames_done = false, have_info = false, planned = false;
  uint64_t n_det_host = 0, halo = 0;
  uint64_t plan_L = 0, plan_off = 0, plan_avail = 0;
  StftPlan plan_host{};
  int n_chunks = 12;
  int stft_precise = 0;      // FMCW_OPT_STFT_PRECISION: 1 = float64 STFT kernel
  bool async_host = false;   // FMCW_OPT_ASYNC_HOST: calls with (pinned) host buffers return after enqueueing
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // start, chain, compact, plan+max, main
  bool ev_valid[5] = {false, false, false, false, false};
  // FMCW_OPT_RUN_GRAPH: a whole fmcw_run as one CUDA graph per set of (device) buffers.  A fleet of small recordings is bound by the
  // GPU's launch rate (about 12 launches per recording), not by its work.
  struct RunKey { const void* iq; uint64_t n; const void* fo[7]; const void* inten; uint64_t cap, ld; uint32_t layout; int precise;
                  const void* info_dst; };
  struct RunGraph { RunKey key; cudaGraphExec_t exec; int seen; unsigned long long epoch; };
  std::vector<RunGraph> run_graphs;
  bool run_graph_opt = false, capturing = false;
  fmcw_device_info* info_dst = nullptr;
};

namespace {

struct BusyGuard {
  fmcw_handle* h; bool ok;
  explicit BusyGuard(fmcw_handle* hh) : h(hh), ok(!hh->busy.test_and_set(std::memory_order_acquire)) {}
  ~BusyGuard() { if (ok) h->busy.clear(std::memory_order_release); }
};

fmcw_status fail(fmcw_handle* h, fmcw_status s, const std::string& msg) {
  if (h) h->err = msg;
  return s;
}
fmcw_status cuda_fail(fmcw_handle* h, cudaError_t e, const char* what) {
  cudaGetLastError();
  const fmcw_status s = (e == cudaErrorMemoryAllocation) ? FMCW_ERR_OOM : FMCW_ERR_CUDA;
  return fail(h, s, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CK(call, what)                                          \
  do {                                                          \
    cudaError_t _e = (call);                                    \
    if (_e != cudaSuccess) return cuda_fail(h, _e, what);       \
  } while (0)

template <class T>
cudaError_t upload(DevBuf& b, const std::vector<T>& v, cudaStream_t st) {
  cudaError_t e = b.ensure(v.size() * sizeof(T) + 16);
  if (e != cudaSuccess) return e;
  return cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

fmcw_status validate(const fmcw_config* c, std::string& why) {
  if (!c) { why = "cfg is NULL"; return FMCW_ERR_POINTER; }
  if (c->struct_size != sizeof(fmcw_config)) { why = "fmcw_config.struct_size mismatch"; return FMCW_ERR_CONFIG; }
  if (c->range_fft_size != (uint32_t)NR) { why = "range_fft_size must be 256 (RP:118)"; return FMCW_ERR_CONFIG; }
  const uint32_t nd = c->Doppler_fft_size;
  if (nd < 2 || nd > (uint32_t)MAX_ND || (nd & (nd - 1))) { why = "Doppler_fft_size must be a power of two in [2,64]"; return FMCW_ERR_CONFIG; }
  if (c->max_num_targets != 1) { why = "max_num_targets must be 1 (RP:129)"; return FMCW_ERR_CONFIG; }
  if (c->num_ADC_samples_per_chirp < 1 || c->num_ADC_samples_per_chirp > 4096) { why = "num_ADC_samples_per_chirp out of [1,4096]"; return FMCW_ERR_CONFIG; }
  if (c->num_chirps_per_frame < 1 || c->num_chirps_per_frame > 4096) { why = "num_chirps_per_frame out of [1,4096]"; return FMCW_ERR_CONFIG; }
  if (c->num_Rx_antennas < 1 || c->rx_select >= c->num_Rx_antennas) { why = "rx_select out of range"; return FMCW_ERR_CONFIG; }
  if (c->window_length < 2 || c->window_length > 400) { why = "window_length out of [2,400]"; return FMCW_ERR_CONFIG; }
  if (c->overlap >= c->window_length) { why = "overlap must be < window_length (RP:179)"; return FMCW_ERR_CONFIG; }
  if (c->MAX_FREQ_BINS < 2 || c->MAX_FREQ_BINS > (uint32_t)MAX_NQ) { why = "MAX_FREQ_BINS out of [2,1024]"; return FMCW_ERR_CONFIG; }
  if (c->peak_mode > 1) { why = "peak_mode"; return FMCW_ERR_CONFIG; }
  if (!(c->PRT > 0) || !(c->adc_scale > 0) || !(c->dist_per_bin > 0)) { why = "PRT, adc_scale and dist_per_bin must be positive"; return FMCW_ERR_CONFIG; }
  return FMCW_OK;
}

void fill_tables(fmcw_handle* h) {
  h->st.plan = h->plan.as<StftPlan>();
  h->st.bins = h->bins.as<int>();
  h->st.kcb = h->kcb.as<float>();
  h->st.wdc = h->wdc.as<float>();
  h->st.qpos = h->qpos.as<int>();
  h->st.aq = h->aq.as<float>();
  h->st.qend = h->qend.as<int>();
  h->st.coef = h->coef.as<float>();
  h->st.win = h->swin.as<float>();
  h->st.win_d = h->swin_d.as<double>();
  h->st.hard_list = h->hard.as<unsigned int>();
  h->st.tcB = h->tcb.as<float>();
}

fmcw_status read_info(fmcw_handle* h) {
  CK(cudaStreamSynchronize(h->stream), "synchronize");
  // a look-ahead plan forked by the last call may still be rewriting the plan struct on the side stream
  if (h->lookahead) CK(cudaStreamSynchronize(h->side), "synchronize look-ahead plan");
  unsigned long long nd = 0;
  if (h->frames_done) CK(cudaMemcpy(&nd, h->ndet.p, sizeof(nd), cudaMemcpyDeviceToHost), "read n_det");
  h->n_det_host = nd;
  CK(cudaMemcpy(&h->plan_host, h->plan.p, sizeof(StftPlan), cudaMemcpyDeviceToHost), "read plan");
  int derr = 0;
  CK(cudaMemcpy(&derr, h->derr.p, sizeof(int), cudaMemcpyDeviceToHost), "read device status");
  h->have_info = true;
  if (derr != 0) {
    int zero = 0;
    cudaMemcpy(h->derr.p, &zero, sizeof(int), cudaMemcpyHostToDevice);
    const char* why = derr == -2 ? "more DTFT bins than the plan tables hold" :
                      derr == -3 ? "a query chunk exceeds the shared-memory bin budget" :
                      derr == -4 ? "intensity capacity_cols is smaller than the column count" :
                      derr == -5 ? "too many columns need the exhaustive max search" :
                      derr == -6 ? "a rank never posted to the mailbox (timeout)" : "device-side failure";
    return fail(h, derr == -4 ? FMCW_ERR_SIZE : derr == -6 ? FMCW_ERR_STATE : FMCW_ERR_CUDA, why);
  }
  return FMCW_OK;
}

// frames -> device buffers (internal or caller's), compaction; no synchronisation
struct FrameDev { float* rmax; int32_t* det; int32_t* rbin; float* rmag; int32_t* dbin; float2* drow; float* slow; };

fmcw_status run_frames(fmcw_handle* h, const int16_t* iq, uint64_t n_frames, const fmcw_frame_out* out, FrameDev& d,
                       bool& any_host) {
  const fmcw_config& c = h->cfg;
  const uint32_t NTS = c.num_ADC_samples_per_chirp, PN = c.num_chirps_per_frame, ND = c.Doppler_fft_size;
  if (!iq && n_frames) return fail(h, FMCW_ERR_POINTER, "iq is NULL");
  any_host = false;
  const size_t rx_words = (size_t)PN * NTS;           // uint32 words per (frame, rx)
  uint32_t n_rx = c.num_Rx_antennas, rx_sel = c.rx_select;
  const uint32_t* d_iq = reinterpret_cast<const uint32_t*>(iq);
  if (n_frames && !is_device_ptr(iq)) {
    any_host = true;
    CK(h->iq_stage.ensure(n_frames * rx_words * 4), "alloc iq staging");
    // only the selected RX crosses PCIe (the reference ignores the other antennas, RP:202)
    CK(cudaMemcpy2DAsync(h->iq_stage.p, rx_words * 4, reinterpret_cast<const uint32_t*>(iq) + (size_t)rx_sel * rx_words,
                         (size_t)n_rx * rx_words * 4, rx_words * 4, n_frames, cudaMemcpyHostToDevice, h->stream),
       "H2D iq");
    d_iq = h->iq_stage.as<uint32_t>();
    n_rx = 1; rx_sel = 0;
  }
  auto pick = [&](void* user, DevBuf& b, size_t bytes, bool needed, void*& dst) -> cudaError_t {
    dst = nullptr;
    if (user && is_device_ptr(user)) { dst = user; return cudaSuccess; }
    if (user) any_host = true;
    if (!user && !needed) return cudaSuccess;
    cudaError_t e = b.ensure(bytes ? bytes : 16);
    dst = b.p;
    return e;
  };
  const size_t nf = n_frames ? n_frames : 1;
  void* t;
  CK(pick(out ? out->range_max_abs : nullptr, h->o_rmax, nf * NR * 4, false, t), "alloc"); d.rmax = (float*)t;
  CK(pick(out ? out->detected : nullptr, h->o_det, nf * 4, true, t), "alloc"); d.det = (int32_t*)t;
  CK(pick(out ? out->range_bin : nullptr, h->o_rbin, nf * 4, false, t), "alloc"); d.rbin = (int32_t*)t;
  CK(pick(out ? out->range_mag : nullptr, h->o_rmag, nf * 4, false, t), "alloc"); d.rmag = (float*)t;
  CK(pick(out ? out->doppler_bin : nullptr, h->o_dbin, nf * 4, false, t), "alloc"); d.dbin = (int32_t*)t;
  CK(pick(out ? out->doppler_row : nullptr, h->o_drow, nf * ND * 8, false, t), "alloc"); d.drow = (float2*)t;
  CK(pick(out ? out->slow_time_mag : nullptr, h->o_slow, nf * PN * 4, false, t), "alloc"); d.slow = (float*)t;
  CK(h->o_slow64.ensure(nf * PN * sizeof(sig_t)), "alloc slow-time rows");
  CK(h->det_list.ensure(((nf + 1023) / 1024 * 33 + 8) * 4), "alloc compaction counts");
  CK(h->xc.ensure((nf * PN + c.window_length) * sizeof(sig_t)), "alloc slow-time signal");
  CK(h->colub.ensure((nf * PN + c.window_length) * 4), "alloc column bounds");
  h->st.col_ub = h->colub.as<float>();

  ChainParams p{};
  p.iq = d_iq; p.n_frames = n_frames; p.NTS = NTS; p.PN = PN; p.n_rx = n_rx; p.rx_sel = rx_sel; p.ND = ND;
  p.nts_fft = NTS < (uint32_t)NR ? NTS : (uint32_t)NR;
  p.win_tab = h->win_tab.as<float4>(); p.tw_pair = h->tw_pair.as<float2>();
  p.win_tab_d = h->win_tab_d.as<double>(); p.tw_d = h->tw_d.as<double2>(); p.hfft_d = h->hfft_d.as<double2>();
  p.tw_re = h->tw_re.as<float>(); p.tw_im = h->tw_im.as<float>();
  p.dop_tw = h->dop_tw.as<float2>(); p.dop_win = h->dop_win.as<float>();
  p.bin_lo = h->bin_lo; p.bin_hi = h->bin_hi;
  p.range_thr = (float)c.range_threshold; p.dop_thr = (float)c.Doppler_threshold; p.peak_mode = (int)c.peak_mode;
  p.range_max_abs = d.rmax; p.detected = d.det; p.range_bin = d.rbin; p.range_mag = d.rmag;
  p.doppler_bin = d.dbin; p.doppler_row = d.drow; p.slow_mag = d.slow; p.slow64 = h->o_slow64.as<sig_t>();
  p.spec_out = nullptr;
  for (bool& v : h->ev_valid) v = false;
  const bool timed = !h->capturing;                   // (events recorded into a graph cannot be read back)
  if (timed) { CK(cudaEventRecord(h->ev[0], h->stream), "event"); h->ev_valid[0] = true; }
  CK(launch_frame_chain(p, h->stream), "frame chain kernel");
  if (timed) { CK(cudaEventRecord(h->ev[1], h->stream), "event"); h->ev_valid[1] = true; }
  CompactParams cp{d.det, n_frames, PN, h->o_slow64.as<sig_t>(), h->xc.as<sig_t>(), nullptr,
                   h->ndet.as<unsigned long long>(), h->det_list.as<uint32_t>()};
  CK(launch_compact(cp, h->stream), "compaction kernels");
  if (timed) { CK(cudaEventRecord(h->ev[2], h->stream), "event"); h->ev_valid[2] = true; }
  h->n_frames = n_frames; h->frames_done = true; h->have_info = false; h->planned = false; h->halo = 0;
  return FMCW_OK;
}

fmcw_status copy_frame_outputs(fmcw_handle* h, uint64_t n, const fmcw_frame_out* out, const FrameDev& d) {
  if (!out || !n) return FMCW_OK;
  const fmcw_config& c = h->cfg;
  auto back = [&](void* user, const void* dev, size_t bytes) -> cudaError_t {
    if (!user || user == dev || is_device_ptr(user)) return cudaSuccess;
    return cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, h->stream);
  };
  CK(back(out->range_max_abs, d.rmax, n * NR * 4), "D2H range_max_abs");
  CK(back(out->detected, d.det, n * 4), "D2H detected");
  CK(back(out->range_bin, d.rbin, n * 4), "D2H range_bin");
  CK(back(out->range_mag, d.rmag, n * 4), "D2H range_mag");
  CK(back(out->doppler_bin, d.dbin, n * 4), "D2H doppler_bin");
  CK(back(out->doppler_row, d.drow, n * c.Doppler_fft_size * 8), "D2H doppler_row");
  CK(back(out->slow_time_mag, d.slow, n * c.num_chirps_per_frame * 4), "D2H slow_time_mag");
  return FMCW_OK;
}

// Every plan launch first joins a look-ahead plan that may be in flight; the plan kernel then confirms the
// assumed layout (spec_mode 2) or plans again.
static cudaError_t join_lookahead(fmcw_handle* h, int& spec_mode) {
  spec_mode = 0;
  if (!h->lookahead) return cudaSuccess;
  h->lookahead = false;
  spec_mode = 2;
  return cudaStreamWaitEvent(h->stream, h->ev_join, 0);
}
static fmcw_status fork_lookahead(fmcw_handle* h, uint64_t L_total, uint64_t offset, uint64_t L_local, uint64_t L_avail) {
  CK(cudaEventRecord(h->ev_fork, h->stream), "event");
  CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0), "fork look-ahead plan");
  CK(launch_stft_plan(h->st, h->geom, nullptr, h->cfg.num_chirps_per_frame, L_total, offset, L_local, L_avail, h->n_chunks,
                      h->side, nullptr, 0, 0, nullptr, 1), "look-ahead stft plan");
  CK(cudaEventRecord(h->ev_join, h->side), "event");
  h->lookahead = true;
  return FMCW_OK;
}

// plan (+ max) + main on the signal held in h->xc.  d_ndet != null: sizes come from the device.
fmcw_status run_stft(fmcw_handle* h, bool from_device_count, uint64_t L_total, uint64_t offset, uint64_t L_local,
                     uint64_t L_avail, bool compute_max, double pmax_override, const fmcw_stft_out* sout,
                     uint64_t cols_upper) {
  if (!sout || !sout->intensity) return fail(h, FMCW_ERR_POINTER, "stft output buffer is NULL");
  if (sout->layout > 1) return fail(h, FMCW_ERR_CONFIG, "unknown intensity layout");
  const uint32_t nq = h->cfg.MAX_FREQ_BINS;
  const bool dev_out = is_device_ptr(sout->intensity);
  uint64_t cap = sout->capacity_cols;
  uint64_t ld = sout->ld_cols ? sout->ld_cols : sout->capacity_cols;
  if (sout->layout == FMCW_LAYOUT_FREQ_MAJOR && ld < cap) return fail(h, FMCW_ERR_SIZE, "ld_cols < capacity_cols");
  float* d_out = sout->intensity;
  uint64_t d_ld = ld;
  if (!dev_out) {
    const uint64_t need = cols_upper < cap ? cols_upper : cap;     // columns the staging buffer must hold
    CK(h->inten.ensure((size_t)(need ? need : 1) * nq * 4), "alloc intensity staging");
    d_out = h->inten.as<float>();
    cap = need; d_ld = need;
  }
  if (!h->planned) {
    int spec_mode = 0;
    CK(join_lookahead(h, spec_mode), "join look-ahead plan");
    static const int use_graph = env_int("FMCW_GRAPH", 1) != 0;
    bool graphable = use_graph && from_device_count && compute_max && spec_mode == 2 && ++h->mx_eligible_calls > 1;
    const bool mx_stale = !h->mx_exec || memcmp(&h->mx_key_tables, &h->st, sizeof(StftTables)) != 0 || h->mx_key_xc != h->xc.p;
    if (h->capturing) graphable = false;                   // inside the recording of a whole run: plain launches join the outer graph
    if (graphable) {
      if (mx_stale) {
        if (h->mx_exec) { cudaGraphExecDestroy(h->mx_exec); h->mx_exec = nullptr; }
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal), "begin capture");
        cudaError_t e1 = launch_stft_plan(h->st, h->geom, h->ndet.as<unsigned long long>(), h->cfg.num_chirps_per_frame, 0, 0, 0,
                                          0, h->n_chunks, h->stream, nullptr, 0, 0, nullptr, 2);
        cudaError_t e2 = launch_stft_max(h->st, h->geom, h->xc.as<sig_t>(), h->stream);
        cudaError_t e3 = cudaStreamEndCapture(h->stream, &graph);
        CK(e1, "capture stft plan"); CK(e2, "capture stft max"); CK(e3, "end capture");
        cudaError_t e4 = cudaGraphInstantiate(&h->mx_exec, graph, 0);
        cudaGraphDestroy(graph);
        CK(e4, "instantiate graph");
        h->mx_key_tables = h->st; h->mx_key_xc = h->xc.p;
      }
      CK(cudaGraphLaunch(h->mx_exec, h->stream), "launch plan + max graph");
    } else {
      CK(launch_stft_plan(h->st, h->geom, from_device_count ? h->ndet.as<unsigned long long>() : nullptr,
                          h->cfg.num_chirps_per_frame, L_total, offset, L_local, L_avail, h->n_chunks, h->stream,
                          nullptr, 0, 0, nullptr, spec_mode), "stft plan kernel");
      if (compute_max) CK(launch_stft_max(h->st, h->geom, h->xc.as<sig_t>(), h->stream), "stft max kernels");
    }
    h->planned = true; h->plan_L = L_total; h->plan_off = offset; h->plan_avail = L_avail;
  }
  if (!compute_max) CK(launch_stft_set_max(h->st, pmax_override, h->stream), "stft set max");
  if (!h->capturing) { CK(cudaEventRecord(h->ev[3], h->stream), "event"); h->ev_valid[3] = true; }
  CK(launch_stft_main(h->st, h->geom, h->xc.as<sig_t>(), d_out, cap, d_ld, (int)sout->layout, h->derr.as<int>(), h->stream,
                      nullptr, h->stft_precise), "stft main kernel");
  if (!h->capturing) { CK(cudaEventRecord(h->ev[4], h->stream), "event"); h->ev_valid[4] = true; }
  h->have_info = false;
  if (!dev_out && h->async_host) {
    // no host round trip: copy the upper bound of the column count (columns past the real count are unspecified)
    if (cap) {
      if (sout->layout == FMCW_LAYOUT_TIME_MAJOR)
        CK(cudaMemcpyAsync(sout->intensity, d_out, (size_t)cap * nq * 4, cudaMemcpyDeviceToHost, h->stream), "D2H intensity");
      else
        CK(cudaMemcpy2DAsync(sout->intensity, ld * 4, d_out, d_ld * 4, cap * 4, nq, cudaMemcpyDeviceToHost, h->stream),
           "D2H intensity");
    }
    return FMCW_OK;
  }
  if (!dev_out) {
    fmcw_status s = read_info(h);
    if (s != FMCW_OK) return s;
    const uint64_t ncl = h->plan_host.valid > 0 ? h->plan_host.col_end - h->plan_host.col_begin : 0;
    if (ncl > sout->capacity_cols) return fail(h, FMCW_ERR_SIZE, "intensity capacity_cols is smaller than the column count");
    if (ncl) {
      if (sout->layout == FMCW_LAYOUT_TIME_MAJOR)
        CK(cudaMemcpyAsync(sout->intensity, d_out, (size_t)ncl * nq * 4, cudaMemcpyDeviceToHost, h->stream), "D2H intensity");
      else
        CK(cudaMemcpy2DAsync(sout->intensity, ld * 4, d_out, d_ld * 4, ncl * 4, nq, cudaMemcpyDeviceToHost, h->stream),
           "D2H intensity");
    }
    CK(cudaStreamSynchronize(h->stream), "synchronize");
  }
  return FMCW_OK;
}

// fmcw_set_info_target: the scalars of the run, on the device (the same fields fmcw_get_info derives on the host)
__global__ void info_snapshot_kernel(const StftPlan* __restrict__ P, const unsigned long long* __restrict__ ndet,
                                     const int* __restrict__ derr, uint64_t n_frames, uint32_t PN, fmcw_device_info* dst) {
  fmcw_device_info d;
  memset(&d, 0, sizeof(d));
  d.info.n_frames = n_frames;
  d.info.n_detected = *ndet;
  d.info.L_local = *ndet * PN;
  d.info.L_total = P->L_total; d.info.sample_offset = P->sample_offset; d.info.nfft = P->nfft;
  d.info.ncol_total = P->ncol_total;
  if (P->valid > 0) { d.info.col_begin = P->col_begin; d.info.ncol_local = P->col_end - P->col_begin; }
  d.info.n_dtft_bins = (uint32_t)P->nb; d.info.n_refined = P->n_refined;
  d.info.pmax_raw = P->pmax_raw;
  d.status = *derr != 0 ? *derr : (P->valid < 0 ? P->valid : 0);
  *dst = d;
}

// the body of fmcw_run: look-ahead plan, frames, compaction, STFT (and the info snapshot); no synchronisation with device buffers
fmcw_status run_body(fmcw_handle* h, const int16_t* iq, uint64_t n_frames, const fmcw_frame_out* fout, const fmcw_stft_out* sout,
                     bool& any_host) {
  FrameDev d{};
  const uint64_t L_spec = n_frames * h->cfg.num_chirps_per_frame;
  if (L_spec >= h->cfg.window_length) {
    // look-ahead: plan the STFT for "every frame detects a target" on a side stream while the frame chain runs;
    // the real plan launch confirms it on the device (or plans again if the detection count differs)
    fmcw_status fs = fork_lookahead(h, L_spec, 0, L_spec, L_spec);
    if (fs != FMCW_OK) return fs;
  }
  fmcw_status s = run_frames(h, iq, n_frames, fout, d, any_host);
  if (s != FMCW_OK) return s;
  s = copy_frame_outputs(h, n_frames, fout, d);
  if (s != FMCW_OK) return s;
  const uint64_t L_up = n_frames * h->cfg.num_chirps_per_frame;
  const uint64_t cols_up = L_up >= h->cfg.window_length ? (L_up - h->cfg.overlap) / h->geom.hop : 0;
  s = run_stft(h, true, 0, 0, 0, 0, true, 0.0, sout, cols_up);
  if (s != FMCW_OK) return s;
  if (h->info_dst) {
    info_snapshot_kernel<<<1, 1, 0, h->stream>>>(h->plan.as<StftPlan>(), h->ndet.as<unsigned long long>(), h->derr.as<int>(), n_frames,
                                                 h->cfg.num_chirps_per_frame, h->info_dst);
    CK(cudaGetLastError(), "info snapshot kernel");
  }
  return FMCW_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* fmcw_version(void) { return "libfmcw_cuda 0.1.0 (sm_100a)"; }

const char* fmcw_status_string(fmcw_status s) {
  switch (s) {
    case FMCW_OK: return "ok";
    case FMCW_ERR_CONFIG: return "bad configuration";
    case FMCW_ERR_POINTER: return "bad pointer";
    case FMCW_ERR_CUDA: return "CUDA error";
    case FMCW_ERR_NCCL: return "NCCL error";
    case FMCW_ERR_OOM: return "out of memory";
    case FMCW_ERR_BUSY: return "handle busy";
    case FMCW_ERR_SIZE: return "bad size";
    case FMCW_ERR_NO_DATA: return "not enough slow-time samples";
    case FMCW_ERR_STATE: return "call order violated";
  }
  return "unknown";
}

fmcw_status fmcw_create(const fmcw_config* cfg, const double* calib_data, uint64_t calib_len, int device,
                        fmcw_handle** out) {
  if (!out) return FMCW_ERR_POINTER;
  *out = nullptr;
  std::string why;
  fmcw_status vs = validate(cfg, why);
  if (vs != FMCW_OK) return vs;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return FMCW_ERR_CUDA; }   // no CPU fallback
  if (device < 0 || device >= ndev) return FMCW_ERR_CONFIG;
  fmcw_handle* h = new (std::nothrow) fmcw_handle();
  if (!h) return FMCW_ERR_OOM;
  h->cfg = *cfg;
  h->device = device;
  const fmcw_config& c = h->cfg;
  const uint32_t NTS = c.num_ADC_samples_per_chirp, PN = c.num_chirps_per_frame, ND = c.Doppler_fft_size;
  fmcw_status rc = FMCW_OK;
  auto bail = [&](fmcw_status s) { fmcw_destroy(h); return s; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(FMCW_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(FMCW_ERR_CUDA);
  for (cudaEvent_t& e : h->ev)
    if (cudaEventCreate(&e) != cudaSuccess) return bail(FMCW_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess) return bail(FMCW_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) return bail(FMCW_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) return bail(FMCW_ERR_CUDA);

  // ---- calibration (RP:167-174), window (RP:138) and scale (RP:121, 203) folded into one table ----
  std::vector<std::complex<double>> cal(NTS, {0.0, 0.0});
  if (calib_data && calib_len) {
    const uint64_t N_cal = calib_len / (2ull * c.num_Rx_antennas);
    const uint64_t dec = N_cal / NTS;
    if (dec == 0 || N_cal * 2ull * c.num_Rx_antennas != calib_len) return bail(FMCW_ERR_SIZE);
    const uint64_t base = 2ull * c.rx_select * N_cal;
    for (uint32_t n = 0; n < NTS && (uint64_t)n * dec < N_cal; ++n)
      cal[n] = {calib_data[base + (uint64_t)n * dec], calib_data[base + N_cal + (uint64_t)n * dec]};
  }
  std::complex<double> cmean(0.0, 0.0);
  for (auto& v : cal) cmean += v;
  cmean /= (double)NTS;
  const std::vector<double> wr = blackman_sym((int)NTS);
  const uint32_t nts_fft = NTS < (uint32_t)NR ? NTS : (uint32_t)NR;
  std::vector<float4> wt(NR, make_float4(0.f, 0.f, 0.f, 0.f));
  for (uint32_t n = 0; n < nts_fft; ++n) {
    const double w2 = 2.0 * wr[n];
    const std::complex<double> hh = w2 * c.IF_scale * (cal[n] - cmean);
    wt[n] = make_float4((float)(w2 * c.IF_scale / (c.adc_scale * (double)NTS)), (float)hh.real(), (float)hh.imag(), 0.f);
  }
  std::vector<double> wtd(3 * (size_t)NR, 0.0);
  for (uint32_t n = 0; n < nts_fft; ++n) {
    const double w2 = 2.0 * wr[n];
    const std::complex<double> hh = w2 * c.IF_scale * (cal[n] - cmean);
    wtd[3 * n] = w2 * c.IF_scale / (c.adc_scale * (double)NTS); wtd[3 * n + 1] = hh.real(); wtd[3 * n + 2] = hh.imag();
  }
  std::vector<double2> twd(NR);
  for (int k = 0; k < NR; ++k) {
    const double a = -2.0 * M_PI * (double)k / NR;
    twd[k] = make_double2(std::cos(a), std::sin(a));
  }
  std::vector<double2> hfft(NR);      // H[k] = sum_n h[n] W^(n k): the calibration term of the slow-time row at bin k
  for (int k = 0; k < NR; ++k) {
    double hr = 0.0, hi = 0.0;
    for (uint32_t n = 0; n < nts_fft; ++n) {
      const double2 tw = twd[((uint32_t)n * (uint32_t)k) & (NR - 1)];
      hr += wtd[3 * n + 1] * tw.x - wtd[3 * n + 2] * tw.y;
      hi += wtd[3 * n + 1] * tw.y + wtd[3 * n + 2] * tw.x;
    }
    hfft[k] = make_double2(hr, hi);
  }
  std::vector<float2> twp(256);
  for (int k1 = 0; k1 < 16; ++k1)
    for (int s = 0; s < 16; ++s) {
      const double a = -2.0 * M_PI * (double)(s * k1) / NR;
      twp[k1 * 16 + s] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
  std::vector<float> twre(272, 0.f), twim(272, 0.f);
  for (int k = 0; k < NR; ++k) {
    const double a = -2.0 * M_PI * (double)k / NR;
    twre[k + (k >> 4)] = (float)std::cos(a);
    twim[k + (k >> 4)] = (float)std::sin(a);
  }
  std::vector<float2> dtw(ND);
  for (uint32_t k = 0; k < ND; ++k) {
    const double a = -2.0 * M_PI * (double)k / ND;
    dtw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  const std::vector<double> wd = chebwin_sym((int)PN, 100.0);
  std::vector<float> dwin(ND, 0.f);
  std::vector<double> dwin_d(ND, 0.0);
  for (uint32_t i = 0; i < ND && i < PN; ++i) { dwin[i] = (float)(2.0 * wd[i]); dwin_d[i] = 2.0 * wd[i]; }
  // range gate of f_search_peak: (n-1)*dist_per_bin in [min_distance, max_distance], n = 3..len-2
  h->bin_lo = NR; h->bin_hi = -1;
  for (int n = 3; n <= NR - 2; ++n) {
    const double r = (double)(n - 1) * c.dist_per_bin;
    if (r >= c.min_distance && r <= c.max_distance) { if (n - 1 < h->bin_lo) h->bin_lo = n - 1; h->bin_hi = n - 1; }
  }
  const std::vector<double> wk = kaiser_sym((int)c.window_length, c.kaiser_beta);
  std::vector<float> swin(wk.begin(), wk.end());

  h->geom.win = c.window_length; h->geom.hop = c.window_length - c.overlap; h->geom.nq = c.MAX_FREQ_BINS;
  h->geom.fs = 1.0 / c.PRT;
  {
    // rho = sup over [pi/(win-1), pi] of |W(w)|/W(0): grid maximum plus the Lipschitz slack of the grid spacing.
    // With it, sup |S(w)| <= rho * sum(w x) + sum w |x - xbar| certifies most candidate columns in O(win).
    const int win = (int)c.window_length;
    double W0 = 0.0, Dw = 0.0;
    const double c0 = 0.5 * (win - 1);
    for (int n = 0; n < win; ++n) { W0 += wk[n]; Dw += std::fabs(n - c0) * wk[n]; }
    double rho = 0.0;
    if (win > 2) {
      const int G = 64 * win;
      const double w_lo = M_PI / (win - 1), delta = (M_PI - w_lo) / G;
      for (int gi = 0; gi < G; ++gi) {
        const double w = w_lo + (gi + 0.5) * delta;
        double re = 0.0, im = 0.0;
        for (int n = 0; n < win; ++n) { re += wk[n] * std::cos(w * (n - c0)); im += wk[n] * std::sin(w * (n - c0)); }
        const double mag = std::sqrt(re * re + im * im);
        rho = mag > rho ? mag : rho;
      }
      rho = (rho + Dw * delta * 0.5) / W0;
    }
    h->geom.rho = (float)(rho * (1.0 + 1e-6));
  }
  const int nb_max = 2 * (int)c.MAX_FREQ_BINS + 2;
  const int half = (int)c.window_length / 2;
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  ok(upload(h->win_tab, wt, h->stream)); ok(upload(h->tw_pair, twp, h->stream));
  ok(upload(h->win_tab_d, wtd, h->stream)); ok(upload(h->tw_d, twd, h->stream)); ok(upload(h->hfft_d, hfft, h->stream)); ok(upload(h->dop_win_d, dwin_d, h->stream));
  ok(upload(h->tw_re, twre, h->stream)); ok(upload(h->tw_im, twim, h->stream));
  ok(upload(h->dop_tw, dtw, h->stream)); ok(upload(h->dop_win, dwin, h->stream));
  ok(upload(h->swin, swin, h->stream)); ok(upload(h->swin_d, wk, h->stream));
  ok(h->plan.ensure(sizeof(StftPlan))); ok(h->bins.ensure((size_t)nb_max * 4)); ok(h->kcb.ensure((size_t)nb_max * 4)); ok(h->wdc.ensure((size_t)nb_max * 4));
  ok(h->qpos.ensure(MAX_NQ * 4)); ok(h->aq.ensure(MAX_NQ * 4)); ok(h->qend.ensure((size_t)(nb_max + 2) * 4));
  ok(h->coef.ensure((size_t)nb_max * 2 * half * 4 + 64));
  h->st.hard_cap = 1u << 20;
  ok(h->hard.ensure((size_t)h->st.hard_cap * 4));
  ok(h->tcb.ensure(stft_tc_table_bytes(nb_max) + 256));
  ok(h->derr.ensure(16)); ok(h->ndet.ensure(16)); ok(h->gmax.ensure(16)); ok(h->mb_step.ensure(16));
  if (e == cudaSuccess) {
    ok(cudaMemsetAsync(h->plan.p, 0, sizeof(StftPlan), h->stream));
    ok(cudaMemsetAsync(h->derr.p, 0, 16, h->stream));
    ok(cudaMemsetAsync(h->ndet.p, 0, 16, h->stream));
    ok(cudaMemsetAsync(h->mb_step.p, 0, 16, h->stream));
    ok(cudaStreamSynchronize(h->stream));
  }
  if (e != cudaSuccess) { rc = (e == cudaErrorMemoryAllocation) ? FMCW_ERR_OOM : FMCW_ERR_CUDA; cudaGetLastError(); return bail(rc); }
  h->st.nb_max = nb_max;
  fill_tables(h);
  *out = h;
  return FMCW_OK;
}

void fmcw_destroy(fmcw_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->side) cudaStreamSynchronize(h->side);
  if (h->mx_exec) { cudaGraphExecDestroy(h->mx_exec); h->mx_exec = nullptr; }
  if (h->mbx_exec) { cudaGraphExecDestroy(h->mbx_exec); h->mbx_exec = nullptr; }
  for (auto& e : h->run_graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
  h->run_graphs.clear();
  h->mb_step.release();
  DevBuf* all[] = {&h->win_tab, &h->tw_pair, &h->tw_re, &h->tw_im, &h->dop_tw, &h->dop_win, &h->plan, &h->bins, &h->kcb, &h->wdc,
                   &h->qpos, &h->aq, &h->qend, &h->coef, &h->swin, &h->swin_d, &h->hard, &h->derr, &h->gmax, &h->iq_stage, &h->o_rmax, &h->o_det,
                   &h->o_rbin, &h->o_rmag, &h->o_dbin, &h->o_drow, &h->o_slow, &h->o_slow64, &h->f32_stage, &h->win_tab_d, &h->tw_d, &h->hfft_d, &h->dop_win_d, &h->xc, &h->det_list, &h->ndet, &h->inten,
                   &h->synth_tab, &h->tcb, &h->colub};
  for (DevBuf* b : all) b->release();
  for (cudaEvent_t e : h->ev) if (e) cudaEventDestroy(e);
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* fmcw_last_error(const fmcw_handle* h) { return h ? h->err.c_str() : "NULL handle"; }
void* fmcw_get_stream(fmcw_handle* h) { return h ? (void*)h->stream : nullptr; }

fmcw_status fmcw_synchronize(fmcw_handle* h) {
  if (!h) return FMCW_ERR_POINTER;
  cudaSetDevice(h->device);
  CK(cudaStreamSynchronize(h->stream), "synchronize");
  return FMCW_OK;
}

fmcw_status fmcw_get_info(fmcw_handle* h, fmcw_run_info* info) {
  if (!h || !info) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  fmcw_status s = read_info(h);
  if (s != FMCW_OK) return s;
  const StftPlan& P = h->plan_host;
  std::memset(info, 0, sizeof(*info));
  info->n_frames = h->n_frames;
  info->n_detected = h->n_det_host;
  info->L_local = h->n_det_host * h->cfg.num_chirps_per_frame;
  if (h->planned) {
    info->L_total = P.L_total; info->sample_offset = P.sample_offset; info->nfft = P.nfft;
    info->ncol_total = P.ncol_total;
    if (P.valid > 0) { info->col_begin = P.col_begin; info->ncol_local = P.col_end - P.col_begin; }
    info->n_dtft_bins = (uint32_t)P.nb; info->n_refined = P.n_refined;
    info->pmax_raw = P.pmax_raw;
  }
  return FMCW_OK;
}

fmcw_status fmcw_get_timings(fmcw_handle* h, float* ms) {
  if (!h || !ms) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  CK(cudaStreamSynchronize(h->stream), "synchronize");
  for (int i = 0; i < 4; ++i) {
    ms[i] = 0.f;
    if (h->ev_valid[i] && h->ev_valid[i + 1]) CK(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]), "event elapsed");
  }
  return FMCW_OK;
}

fmcw_status fmcw_process_frames(fmcw_handle* h, const int16_t* iq, uint64_t n_frames, const fmcw_frame_out* out) {
  if (!h) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  FrameDev d{};
  bool any_host = false;
  if (h->mb_world > 0 && !h->lookahead) {
    // mailbox path: plan ahead for "every rank detects a target in each of its (equally many) frames"
    const uint64_t L_loc = n_frames * h->cfg.num_chirps_per_frame, hw = h->cfg.window_length - 1;
    const uint64_t after = (uint64_t)(h->mb_world - 1 - h->mb_rank) * L_loc;
    if (L_loc * h->mb_world >= h->cfg.window_length) {
      fmcw_status fs = fork_lookahead(h, L_loc * h->mb_world, L_loc * h->mb_rank, L_loc, L_loc + (after < hw ? after : hw));
      if (fs != FMCW_OK) return fs;
    }
  }
  fmcw_status s = run_frames(h, iq, n_frames, out, d, any_host);
  if (s != FMCW_OK) return s;
  s = copy_frame_outputs(h, n_frames, out, d);
  if (s != FMCW_OK) return s;
  if (any_host && !h->async_host) CK(cudaStreamSynchronize(h->stream), "synchronize");
  return FMCW_OK;
}

fmcw_status fmcw_set_option(fmcw_handle* h, int option, int64_t value) {
  if (!h) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  switch (option) {
    case FMCW_OPT_ASYNC_HOST: h->async_host = value != 0; return FMCW_OK;
    case FMCW_OPT_STFT_PRECISION:
      if (value != 0 && value != 1) return fail(h, FMCW_ERR_CONFIG, "FMCW_OPT_STFT_PRECISION takes 0 (fast) or 1 (float64)");
      h->stft_precise = (int)value; return FMCW_OK;
    case FMCW_OPT_RUN_GRAPH: h->run_graph_opt = value != 0; return FMCW_OK;
    case FMCW_OPT_STFT_TILES_PER_CTA:
      if (value < 0 || value > 4096) return fail(h, FMCW_ERR_CONFIG, "FMCW_OPT_STFT_TILES_PER_CTA takes 0 .. 4096");
      h->geom.tiles_per_cta = (uint32_t)value;
      g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);     // recorded run graphs carry the old grid
      return FMCW_OK;
    default: return fail(h, FMCW_ERR_CONFIG, "unknown option");
  }
}

fmcw_status fmcw_run(fmcw_handle* h, const int16_t* iq, uint64_t n_frames, const fmcw_frame_out* fout,
                     const fmcw_stft_out* sout) {
  if (!h) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  bool any_host = false;
  // FMCW_OPT_RUN_GRAPH: the whole run as one graph per set of device buffers (first two sightings run plainly: they size the
  // scratch buffers and build the inner plan + max graph; the third is recorded; later ones replay)
  fmcw_handle::RunGraph* rg = nullptr;
  if (h->run_graph_opt && n_frames && sout && sout->intensity && is_device_ptr(iq) && is_device_ptr(sout->intensity)) {
    fmcw_handle::RunKey key;
    std::memset(&key, 0, sizeof(key));
    key.iq = iq; key.n = n_frames;
    const void* fo[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (fout) {
      fo[0] = fout->range_max_abs; fo[1] = fout->detected; fo[2] = fout->range_bin; fo[3] = fout->range_mag;
      fo[4] = fout->doppler_bin; fo[5] = fout->doppler_row; fo[6] = fout->slow_time_mag;
    }
    bool all_dev = true;
    for (int i = 0; i < 7; ++i) { key.fo[i] = fo[i]; if (fo[i] && !is_device_ptr(fo[i])) all_dev = false; }
    key.inten = sout->intensity; key.cap = sout->capacity_cols; key.ld = sout->ld_cols; key.layout = sout->layout;
    key.precise = h->stft_precise; key.info_dst = h->info_dst;
    if (all_dev) {
      for (auto& e : h->run_graphs)
        if (std::memcmp(&e.key, &key, sizeof(key)) == 0) { rg = &e; break; }
      if (!rg && h->run_graphs.size() < 256) {        // (linear search; more buffer sets than this per handle simply run plainly)
        h->run_graphs.push_back(fmcw_handle::RunGraph{key, nullptr, 0, 0});
        rg = &h->run_graphs.back();
      }
    }
  }
  if (rg && rg->exec && rg->epoch != g_alloc_epoch.load(std::memory_order_relaxed)) {
    cudaGraphExecDestroy(rg->exec);              // a scratch buffer moved since the recording
    rg->exec = nullptr; rg->seen = 0;
  }
  if (rg && rg->exec) {
    CK(cudaGraphLaunch(rg->exec, h->stream), "launch run graph");
    for (bool& v : h->ev_valid) v = false;
    h->n_frames = n_frames; h->frames_done = true; h->have_info = false; h->planned = true; h->halo = 0; h->lookahead = false;
    h->plan_L = 0; h->plan_off = 0; h->plan_avail = 0;
    return FMCW_OK;
  }
  const bool capture = rg && ++rg->seen >= 3 && !h->lookahead;
  if (capture) {
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal), "begin capture");
    h->capturing = true;
  }
  fmcw_status s = run_body(h, iq, n_frames, fout, sout, any_host);
  if (capture) {
    h->capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (s != FMCW_OK || e != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      rg->seen = -1000000;                       // this buffer set is not recorded again
      h->lookahead = false; h->frames_done = false; h->planned = false;
      if (s != FMCW_OK) return s;
      return cuda_fail(h, e, "end capture of the run");
    }
    const cudaError_t e2 = cudaGraphInstantiate(&rg->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) { rg->exec = nullptr; rg->seen = -1000000; return cuda_fail(h, e2, "instantiate run graph"); }
    rg->epoch = g_alloc_epoch.load(std::memory_order_relaxed);
    CK(cudaGraphLaunch(rg->exec, h->stream), "launch run graph");
    for (bool& v : h->ev_valid) v = false;
    return FMCW_OK;
  }
  if (s != FMCW_OK) return s;
  if (any_host && !h->async_host) {
    CK(cudaStreamSynchronize(h->stream), "synchronize");
    if (!h->have_info) { s = read_info(h); if (s != FMCW_OK) return s; }
    if (h->plan_host.valid == 0) return fail(h, FMCW_ERR_NO_DATA, "fewer than window_length slow-time samples");
  }
  return FMCW_OK;
}

fmcw_status fmcw_set_info_target(fmcw_handle* h, fmcw_device_info* device_dst) {
  if (!h) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  if (device_dst && !is_device_ptr(device_dst)) return fail(h, FMCW_ERR_POINTER, "the info target must be device memory");
  h->info_dst = device_dst;
  return FMCW_OK;
}

fmcw_status fmcw_stft_frames(fmcw_handle* h, const fmcw_stft_out* sout) {
  if (!h) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done) return fail(h, FMCW_ERR_STATE, "no frames processed");
  h->planned = false; h->halo = 0;
  const uint64_t L_up = h->n_frames * h->cfg.num_chirps_per_frame;
  const uint64_t cols_up = L_up >= h->cfg.window_length ? (L_up - h->cfg.overlap) / h->geom.hop : 0;
  fmcw_status s = run_stft(h, true, 0, 0, 0, 0, true, 0.0, sout, cols_up);
  if (s != FMCW_OK) return s;
  if (sout && sout->intensity && !is_device_ptr(sout->intensity)) {
    if (!h->have_info) { s = read_info(h); if (s != FMCW_OK) return s; }
    if (h->plan_host.valid == 0) return fail(h, FMCW_ERR_NO_DATA, "fewer than window_length slow-time samples");
  }
  return FMCW_OK;
}

fmcw_status fmcw_stft(fmcw_handle* h, const float* x, uint64_t L, const fmcw_stft_out* sout) {
  if (!h || !x) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (L < h->cfg.window_length) return fail(h, FMCW_ERR_NO_DATA, "fewer than window_length samples");
  CK(h->xc.ensure((L + h->cfg.window_length) * sizeof(sig_t)), "alloc signal");
  CK(h->colub.ensure((L + h->cfg.window_length) * 4), "alloc column bounds");
  h->st.col_ub = h->colub.as<float>();
  const float* d_x = x;
  if (!is_device_ptr(x)) {
    CK(h->f32_stage.ensure(L * 4), "alloc signal staging");
    CK(cudaMemcpyAsync(h->f32_stage.p, x, L * 4, cudaMemcpyHostToDevice, h->stream), "copy signal");
    d_x = h->f32_stage.as<float>();
  }
  CK(launch_f32_to_sig(d_x, h->xc.as<sig_t>(), L, h->stream), "widen signal");
  h->frames_done = false; h->planned = false; h->halo = 0; h->n_frames = 0;
  const uint64_t cols = (L - h->cfg.overlap) / h->geom.hop;
  return run_stft(h, false, L, 0, L, L, true, 0.0, sout, cols);
}

fmcw_status fmcw_get_slow_time(fmcw_handle* h, double* dst, uint64_t first, uint64_t count) {
  if (!h || (!dst && count)) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done) return fail(h, FMCW_ERR_STATE, "no frames processed");
  if (!h->have_info) { fmcw_status s = read_info(h); if (s != FMCW_OK) return s; }
  const uint64_t L = h->n_det_host * h->cfg.num_chirps_per_frame;
  if (first + count > L) return fail(h, FMCW_ERR_SIZE, "slow-time range out of bounds");
  if (!count) return FMCW_OK;
  const bool dev = is_device_ptr(dst);
  CK(cudaMemcpyAsync(dst, h->xc.as<sig_t>() + first, count * sizeof(sig_t), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream),
     "copy slow-time samples");
  CK(cudaStreamSynchronize(h->stream), "synchronize");
  return FMCW_OK;
}

fmcw_status fmcw_load_slow_time(fmcw_handle* h, const double* x, uint64_t L_local, uint64_t n_halo) {
  if (!h || (!x && (L_local || n_halo))) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  const uint32_t PN = h->cfg.num_chirps_per_frame;
  if (L_local % PN) return fail(h, FMCW_ERR_SIZE, "L_local must be a multiple of num_chirps_per_frame");
  if (n_halo >= h->cfg.window_length) return fail(h, FMCW_ERR_SIZE, "halo longer than window_length-1");
  CK(h->xc.ensure((L_local + h->cfg.window_length) * sizeof(sig_t)), "alloc slow-time signal");
  CK(h->colub.ensure((L_local + h->cfg.window_length) * 4), "alloc column bounds");
  h->st.col_ub = h->colub.as<float>();
  if (L_local + n_halo)
    CK(cudaMemcpyAsync(h->xc.p, x, (L_local + n_halo) * sizeof(sig_t),
                       is_device_ptr(x) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream), "copy slow-time samples");
  h->n_det_host = L_local / PN;
  h->mb_seed = 0;
  const unsigned long long nd = h->n_det_host;
  CK(cudaMemcpyAsync(h->ndet.p, &nd, sizeof(nd), cudaMemcpyHostToDevice, h->stream), "set detection count");
  CK(cudaStreamSynchronize(h->stream), "synchronize");
  h->n_frames = h->n_det_host; h->frames_done = true; h->have_info = true; h->planned = false; h->halo = n_halo;
  return FMCW_OK;
}

fmcw_status fmcw_set_halo(fmcw_handle* h, const double* src, uint64_t count) {
  if (!h || (!src && count)) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done) return fail(h, FMCW_ERR_STATE, "no frames processed");
  if (count >= h->cfg.window_length) return fail(h, FMCW_ERR_SIZE, "halo longer than window_length-1");
  if (!h->have_info) { fmcw_status s = read_info(h); if (s != FMCW_OK) return s; }
  const uint64_t L = h->n_det_host * h->cfg.num_chirps_per_frame;
  if (count)
    CK(cudaMemcpyAsync(h->xc.as<sig_t>() + L, src, count * sizeof(sig_t),
                       is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream), "copy halo");
  h->halo = count; h->planned = false;
  return FMCW_OK;
}

fmcw_status fmcw_stft_local_max(fmcw_handle* h, uint64_t L_total, uint64_t sample_offset, double* pmax_raw_local) {
  if (!h || !pmax_raw_local) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done) return fail(h, FMCW_ERR_STATE, "no frames processed");
  if (!h->have_info) { fmcw_status s = read_info(h); if (s != FMCW_OK) return s; }
  if (L_total < h->cfg.window_length) return fail(h, FMCW_ERR_NO_DATA, "fewer than window_length slow-time samples");
  const uint64_t L = h->n_det_host * h->cfg.num_chirps_per_frame;
  if (sample_offset + L > L_total) return fail(h, FMCW_ERR_SIZE, "shard exceeds L_total");
  int spec_mode = 0;
  CK(join_lookahead(h, spec_mode), "join look-ahead plan");
  CK(launch_stft_plan(h->st, h->geom, nullptr, h->cfg.num_chirps_per_frame, L_total, sample_offset, L, L + h->halo,
                      h->n_chunks, h->stream, nullptr, 0, 0, nullptr, spec_mode), "stft plan kernel");
  h->planned = true; h->plan_L = L_total; h->plan_off = sample_offset; h->plan_avail = L + h->halo;
  CK(launch_stft_max(h->st, h->geom, h->xc.as<sig_t>(), h->stream), "stft max kernels");
  fmcw_status s = read_info(h);
  if (s != FMCW_OK) return s;
  *pmax_raw_local = h->plan_host.pmax_raw;
  return FMCW_OK;
}

fmcw_status fmcw_stft_sharded(fmcw_handle* h, uint64_t L_total, uint64_t sample_offset, double pmax_raw_global,
                              const fmcw_stft_out* sout) {
  if (!h) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done) return fail(h, FMCW_ERR_STATE, "no frames processed");
  if (!h->have_info) { fmcw_status s = read_info(h); if (s != FMCW_OK) return s; }
  if (!(pmax_raw_global > 0.0)) return fail(h, FMCW_ERR_NO_DATA, "global maximum must be positive");
  const uint64_t L = h->n_det_host * h->cfg.num_chirps_per_frame;
  if (h->planned && (h->plan_L != L_total || h->plan_off != sample_offset || h->plan_avail != L + h->halo)) h->planned = false;
  const uint64_t cols = L / h->geom.hop + 1;
  return run_stft(h, false, L_total, sample_offset, L, L + h->halo, false, pmax_raw_global, sout, cols);
}

// ---- asynchronous sharded path: every hand-off stays on the device --------------------------------
fmcw_status fmcw_shard_pack(fmcw_handle* h, double* msg_dev) {
  if (!h || !msg_dev) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done) return fail(h, FMCW_ERR_STATE, "no frames processed");
  if (!is_device_ptr(msg_dev)) return fail(h, FMCW_ERR_POINTER, "msg must be device memory");
  CK(launch_shard_pack(h->xc.as<sig_t>(), h->ndet.as<unsigned long long>(), h->cfg.num_chirps_per_frame,
                       h->cfg.window_length, msg_dev, h->stream), "shard pack kernel");
  return FMCW_OK;
}

fmcw_status fmcw_shard_plan(fmcw_handle* h, const double* gathered_dev, uint32_t world, uint32_t rank, double* local_max_dev) {
  if (!h || !gathered_dev || !local_max_dev) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done) return fail(h, FMCW_ERR_STATE, "no frames processed");
  if (rank >= world) return fail(h, FMCW_ERR_SIZE, "rank >= world");
  if (!is_device_ptr(gathered_dev) || !is_device_ptr(local_max_dev)) return fail(h, FMCW_ERR_POINTER, "device memory required");
  int spec_mode = 0;
  CK(join_lookahead(h, spec_mode), "join look-ahead plan");
  CK(launch_stft_plan(h->st, h->geom, nullptr, h->cfg.num_chirps_per_frame, 0, 0, 0, 0, h->n_chunks, h->stream,
                      gathered_dev, world, rank, h->xc.as<sig_t>(), spec_mode), "stft plan kernel");
  CK(launch_stft_max(h->st, h->geom, h->xc.as<sig_t>(), h->stream, local_max_dev), "stft max kernels");
  h->planned = true; h->have_info = false;
  return FMCW_OK;
}

fmcw_status fmcw_shard_stft(fmcw_handle* h, const double* global_max_dev, const fmcw_stft_out* sout) {
  if (!h || !global_max_dev || !sout || !sout->intensity) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done || !h->planned) return fail(h, FMCW_ERR_STATE, "fmcw_shard_plan must run first");
  if (!is_device_ptr(global_max_dev) || !is_device_ptr(sout->intensity)) return fail(h, FMCW_ERR_POINTER, "device memory required");
  if (sout->layout > 1) return fail(h, FMCW_ERR_CONFIG, "unknown intensity layout");
  const uint64_t ld = sout->ld_cols ? sout->ld_cols : sout->capacity_cols;
  CK(cudaEventRecord(h->ev[3], h->stream), "event"); h->ev_valid[3] = true;
  CK(launch_stft_main(h->st, h->geom, h->xc.as<sig_t>(), sout->intensity, sout->capacity_cols, ld, (int)sout->layout,
                      h->derr.as<int>(), h->stream, global_max_dev, h->stft_precise), "stft main kernel");
  CK(cudaEventRecord(h->ev[4], h->stream), "event"); h->ev_valid[4] = true;
  h->have_info = false;
  return FMCW_OK;
}

uint64_t fmcw_mailbox_bytes(void) { return (uint64_t)MAILBOX_BYTES; }

static fmcw_status mailbox_args(fmcw_handle* h, void* const* mailboxes, uint32_t world, uint32_t rank, uint64_t step,
                                MailboxSet& mb) {
  if (world == 0 || world > MAILBOX_MAX_WORLD) return fail(h, FMCW_ERR_SIZE, "world must be 1..64");
  if (rank >= world) return fail(h, FMCW_ERR_SIZE, "rank >= world");
  if (step == 0) return fail(h, FMCW_ERR_SIZE, "step numbers start at 1");
  if (h->cfg.window_length > MAILBOX_MAX_WIN) return fail(h, FMCW_ERR_CONFIG, "window_length too large for the mailbox");
  for (uint32_t r = 0; r < MAILBOX_MAX_WORLD; ++r) mb.ptr[r] = r < world ? mailboxes[r] : nullptr;
  for (uint32_t r = 0; r < world; ++r)
    if (!mb.ptr[r]) return fail(h, FMCW_ERR_POINTER, "null mailbox pointer");
  return FMCW_OK;
}

fmcw_status fmcw_mailbox_post_heads(fmcw_handle* h, void* const* mailboxes, uint32_t world, uint32_t rank, uint64_t step) {
  if (!h || !mailboxes) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done) return fail(h, FMCW_ERR_STATE, "no frames processed");
  MailboxSet mb;
  fmcw_status s = mailbox_args(h, mailboxes, world, rank, step, mb);
  if (s != FMCW_OK) return s;
  // the kernels count the passes themselves; the caller's step number re-seeds the counter whenever it does not follow
  if (step != h->mb_step_host + 1) {
    h->mb_seed = step - 1;
    CK(cudaMemcpyAsync(h->mb_step.p, &h->mb_seed, sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream), "seed step counter");
  }
  h->mb_step_host = step;
  h->mbx_collected = false;
  static const int use_graph = env_int("FMCW_GRAPH", 1) != 0;
  h->mbx_use_graph = use_graph && h->lookahead && ++h->mbx_eligible_calls > 1;
  if (h->mbx_use_graph) return FMCW_OK;        // the post kernel is the first node of the graph fmcw_mailbox_plan launches
  CK(launch_mailbox_post_heads(h->xc.as<sig_t>(), h->ndet.as<unsigned long long>(), h->cfg.num_chirps_per_frame,
                               h->cfg.window_length, mb, world, rank, h->mb_step.as<unsigned long long>(), h->stream),
     "mailbox post heads kernel");
  return FMCW_OK;
}

fmcw_status fmcw_mailbox_plan(fmcw_handle* h, void* const* mailboxes, uint32_t world, uint32_t rank, uint64_t step) {
  if (!h || !mailboxes) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done) return fail(h, FMCW_ERR_STATE, "no frames processed");
  MailboxSet mb;
  fmcw_status s = mailbox_args(h, mailboxes, world, rank, step, mb);
  if (s != FMCW_OK) return s;
  void* own = mb.ptr[rank];
  if (step != h->mb_step_host) return fail(h, FMCW_ERR_STATE, "fmcw_mailbox_post_heads of this step must run first");
  unsigned long long* d_step = h->mb_step.as<unsigned long long>();
  int spec_mode = 0;
  CK(join_lookahead(h, spec_mode), "join look-ahead plan");
  h->mb_world = world; h->mb_rank = rank;
  if (h->mbx_use_graph && spec_mode == 2) {
    fmcw_handle::MbxKey key{};
    key.t = h->st; key.xc = h->xc.p; key.world = world; key.rank = rank;
    for (uint32_t r = 0; r < world; ++r) key.ptr[r] = mb.ptr[r];
    if (!h->mbx_exec || memcmp(&key, &h->mbx_key, sizeof(key)) != 0) {
      if (h->mbx_exec) { cudaGraphExecDestroy(h->mbx_exec); h->mbx_exec = nullptr; }
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal), "begin capture");
      cudaError_t e1 = launch_mailbox_post_heads(h->xc.as<sig_t>(), h->ndet.as<unsigned long long>(), h->cfg.num_chirps_per_frame,
                                                 h->cfg.window_length, mb, world, rank, d_step, h->stream);
      cudaError_t e2 = launch_stft_plan(h->st, h->geom, nullptr, h->cfg.num_chirps_per_frame, 0, 0, 0, 0, h->n_chunks, h->stream,
                                        mailbox_heads(own), world, rank, h->xc.as<sig_t>(), 2, mailbox_flag_heads(own), d_step);
      cudaError_t e3 = launch_stft_max(h->st, h->geom, h->xc.as<sig_t>(), h->stream, h->gmax.as<double>());
      cudaError_t e4 = launch_mailbox_post_max(h->gmax.as<double>(), mb, world, rank, d_step, h->stream);
      cudaError_t e5 = launch_mailbox_collect_max(own, world, d_step, h->gmax.as<double>() + 1, h->derr.as<int>(), h->stream);
      cudaError_t e6 = cudaStreamEndCapture(h->stream, &graph);
      CK(e1, "capture post heads"); CK(e2, "capture stft plan"); CK(e3, "capture stft max"); CK(e4, "capture post max");
      CK(e5, "capture collect max"); CK(e6, "end capture");
      cudaError_t e7 = cudaGraphInstantiate(&h->mbx_exec, graph, 0);
      cudaGraphDestroy(graph);
      CK(e7, "instantiate graph");
      h->mbx_key = key;
    }
    CK(cudaGraphLaunch(h->mbx_exec, h->stream), "launch mailbox exchange graph");
    h->mbx_collected = true;
    h->planned = true; h->have_info = false;
    return FMCW_OK;
  }
  if (h->mbx_use_graph) {       // the look-ahead was not joined after all: post now, then the plain sequence
    CK(launch_mailbox_post_heads(h->xc.as<sig_t>(), h->ndet.as<unsigned long long>(), h->cfg.num_chirps_per_frame,
                                 h->cfg.window_length, mb, world, rank, d_step, h->stream), "mailbox post heads kernel");
    h->mbx_use_graph = false;
  }
  CK(launch_stft_plan(h->st, h->geom, nullptr, h->cfg.num_chirps_per_frame, 0, 0, 0, 0, h->n_chunks, h->stream,
                      mailbox_heads(own), world, rank, h->xc.as<sig_t>(), spec_mode, mailbox_flag_heads(own), d_step),
     "stft plan kernel");
  CK(launch_stft_max(h->st, h->geom, h->xc.as<sig_t>(), h->stream, h->gmax.as<double>()), "stft max kernels");
  CK(launch_mailbox_post_max(h->gmax.as<double>(), mb, world, rank, d_step, h->stream), "mailbox post max kernel");
  h->planned = true; h->have_info = false;
  return FMCW_OK;
}

fmcw_status fmcw_mailbox_stft(fmcw_handle* h, void* const* mailboxes, uint32_t world, uint32_t rank, uint64_t step,
                              const fmcw_stft_out* sout) {
  if (!h || !mailboxes || !sout || !sout->intensity) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->frames_done || !h->planned) return fail(h, FMCW_ERR_STATE, "fmcw_mailbox_plan must run first");
  if (!is_device_ptr(sout->intensity)) return fail(h, FMCW_ERR_POINTER, "device memory required");
  if (sout->layout > 1) return fail(h, FMCW_ERR_CONFIG, "unknown intensity layout");
  MailboxSet mb;
  fmcw_status s = mailbox_args(h, mailboxes, world, rank, step, mb);
  if (s != FMCW_OK) return s;
  const uint64_t ld = sout->ld_cols ? sout->ld_cols : sout->capacity_cols;
  if (step != h->mb_step_host) return fail(h, FMCW_ERR_STATE, "fmcw_mailbox_plan of this step must run first");
  if (!h->mbx_collected)
    CK(launch_mailbox_collect_max(mb.ptr[rank], world, h->mb_step.as<unsigned long long>(), h->gmax.as<double>() + 1,
                                  h->derr.as<int>(), h->stream), "mailbox collect max kernel");
  h->mbx_collected = true;
  CK(cudaEventRecord(h->ev[3], h->stream), "event"); h->ev_valid[3] = true;
  CK(launch_stft_main(h->st, h->geom, h->xc.as<sig_t>(), sout->intensity, sout->capacity_cols, ld, (int)sout->layout,
                      h->derr.as<int>(), h->stream, h->gmax.as<double>() + 1, h->stft_precise), "stft main kernel");
  CK(cudaEventRecord(h->ev[4], h->stream), "event"); h->ev_valid[4] = true;
  h->have_info = false;
  return FMCW_OK;
}

fmcw_status fmcw_stft_axes(const fmcw_config* cfg, uint64_t L_total, uint64_t col_begin, uint64_t ncol, double* time,
                           double* frequency, uint64_t* nfft_out, uint64_t* ncol_total) {
  if (!cfg) return FMCW_ERR_POINTER;
  std::string why;
  fmcw_status vs = validate(cfg, why);
  if (vs != FMCW_OK) return vs;
  const uint32_t win = cfg->window_length, hop = win - cfg->overlap;
  const unsigned long long nfft = 1ull << nextpow2_u64(L_total);
  const double fs = 1.0 / cfg->PRT;
  if (nfft_out) *nfft_out = nfft;
  if (ncol_total) *ncol_total = L_total >= win ? (L_total - cfg->overlap) / hop : 0;
  if (time)
    for (uint64_t i = 0; i < ncol; ++i) time[i] = ((double)win / 2.0 + (double)((col_begin + i) * hop)) / fs;   // RP:276
  if (frequency) {
    const uint32_t nq = cfg->MAX_FREQ_BINS;
    const double df = fs / (double)nfft;
    const double d1 = std::log10(df), d2 = std::log10((double)(nfft / 2) * df);               // RP:294-296
    for (uint32_t q = 0; q < nq; ++q) {
      double y = d1 + ((double)q * (d2 - d1)) / (double)(nq - 1);
      if (q == 0) y = d1;
      if (q == nq - 1) y = d2;
      frequency[q] = std::pow(10.0, y);
    }
  }
  return FMCW_OK;
}

fmcw_status fmcw_stft_finegrid(fmcw_handle* h, double f_lo_hz, double f_hi_hz, uint32_t max_rows, float* psd,
                               uint64_t capacity_cols, uint64_t* first_bin, uint64_t* bin_step, uint64_t* n_rows, uint64_t* ncol) {
  if (!h || !psd) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  if (!h->planned) return fail(h, FMCW_ERR_STATE, "no STFT has run on this handle");
  if (!(f_hi_hz >= f_lo_hz) || f_lo_hz < 0.0 || max_rows == 0) return fail(h, FMCW_ERR_CONFIG, "bad band or max_rows");
  if (h->cfg.window_length > 170) return fail(h, FMCW_ERR_CONFIG, "fmcw_stft_finegrid supports window_length <= 170");
  if (!h->have_info) { fmcw_status s = read_info(h); if (s != FMCW_OK) return s; }
  const StftPlan& P = h->plan_host;
  if (P.valid <= 0) return fail(h, FMCW_ERR_NO_DATA, "fewer than window_length slow-time samples");
  const double df = h->geom.fs / (double)P.nfft;
  unsigned long long j0 = (unsigned long long)std::ceil(f_lo_hz / df - 1e-9), j1 = (unsigned long long)std::floor(f_hi_hz / df + 1e-9);
  if (j1 > P.nfft / 2) j1 = P.nfft / 2;
  if (j0 > j1) return fail(h, FMCW_ERR_CONFIG, "no fine-grid bin inside the band");
  const unsigned long long span = j1 - j0 + 1;
  const unsigned long long step = (span + max_rows - 1) / max_rows;
  const unsigned long long rows = (span + step - 1) / step;
  const unsigned long long ncl = P.col_end - P.col_begin;
  if (ncl > capacity_cols) return fail(h, FMCW_ERR_SIZE, "capacity_cols is smaller than the column count");
  const bool dev_out = is_device_ptr(psd);
  float* d_out = psd;
  if (!dev_out) {
    CK(h->inten.ensure((size_t)ncl * rows * 4), "alloc psd staging");
    d_out = h->inten.as<float>();
  }
  CK(launch_stft_finegrid(h->st, h->geom, h->xc.as<sig_t>(), d_out, j0, step, (unsigned)rows, ncl, h->derr.as<int>(), h->stream),
     "stft finegrid kernel");
  if (!dev_out) CK(cudaMemcpyAsync(psd, d_out, (size_t)ncl * rows * 4, cudaMemcpyDeviceToHost, h->stream), "D2H psd");
  CK(cudaStreamSynchronize(h->stream), "synchronize");
  if (first_bin) *first_bin = j0;
  if (bin_step) *bin_step = step;
  if (n_rows) *n_rows = rows;
  if (ncol) *ncol = ncl;
  return FMCW_OK;
}

fmcw_status fmcw_range_spectrum(fmcw_handle* h, const int16_t* iq, uint64_t n_frames, uint64_t frame, uint32_t chirp,
                                float* out) {
  if (!h || !iq || !out) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  const fmcw_config& c = h->cfg;
  if (frame >= n_frames || chirp >= c.num_chirps_per_frame) return fail(h, FMCW_ERR_SIZE, "frame / chirp out of range");
  const size_t frame_words = (size_t)c.num_Rx_antennas * c.num_chirps_per_frame * c.num_ADC_samples_per_chirp;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(iq) + frame * frame_words;
  if (!is_device_ptr(iq)) {
    CK(h->iq_stage.ensure(frame_words * 4), "alloc iq staging");
    CK(cudaMemcpyAsync(h->iq_stage.p, src, frame_words * 4, cudaMemcpyHostToDevice, h->stream), "H2D frame");
    src = h->iq_stage.as<uint32_t>();
  }
  CK(h->o_rmax.ensure(NR * 4), "alloc");
  ChainParams p{};
  p.iq = src; p.n_frames = 1; p.NTS = c.num_ADC_samples_per_chirp; p.PN = c.num_chirps_per_frame;
  p.n_rx = c.num_Rx_antennas; p.rx_sel = c.rx_select; p.ND = c.Doppler_fft_size;
  p.nts_fft = p.NTS < (uint32_t)NR ? p.NTS : (uint32_t)NR;
  p.win_tab = h->win_tab.as<float4>(); p.tw_pair = h->tw_pair.as<float2>();
  p.win_tab_d = h->win_tab_d.as<double>(); p.tw_d = h->tw_d.as<double2>(); p.hfft_d = h->hfft_d.as<double2>();
  p.tw_re = h->tw_re.as<float>(); p.tw_im = h->tw_im.as<float>();
  p.dop_tw = h->dop_tw.as<float2>(); p.dop_win = h->dop_win.as<float>();
  CK(h->o_slow64.ensure((size_t)c.num_chirps_per_frame * sizeof(sig_t)), "alloc");
  p.slow64 = h->o_slow64.as<sig_t>();
  p.bin_lo = NR; p.bin_hi = -1;   // no detection work
  p.range_thr = 0.f; p.dop_thr = 0.f; p.peak_mode = 0;
  const bool dev = is_device_ptr(out);
  float* d_spec = dev ? out : h->o_rmax.as<float>();
  p.spec_out = d_spec; p.spec_frame = 0; p.spec_chirp = chirp;
  CK(launch_frame_chain(p, h->stream), "frame chain kernel");
  if (!dev) CK(cudaMemcpyAsync(out, d_spec, NR * 4, cudaMemcpyDeviceToHost, h->stream), "D2H spectrum");
  CK(cudaStreamSynchronize(h->stream), "synchronize");
  return FMCW_OK;
}

fmcw_status fmcw_range_doppler_map(fmcw_handle* h, const int16_t* iq, uint64_t n_frames, uint64_t frame, float* out_db) {
  if (!h || !iq || !out_db) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  const fmcw_config& c = h->cfg;
  if (frame >= n_frames) return fail(h, FMCW_ERR_SIZE, "frame out of range");
  const size_t frame_words = (size_t)c.num_Rx_antennas * c.num_chirps_per_frame * c.num_ADC_samples_per_chirp;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(iq) + frame * frame_words;
  if (!is_device_ptr(iq)) {
    CK(h->iq_stage.ensure(frame_words * 4), "alloc iq staging");
    CK(cudaMemcpyAsync(h->iq_stage.p, src, frame_words * 4, cudaMemcpyHostToDevice, h->stream), "H2D frame");
    src = h->iq_stage.as<uint32_t>();
  }
  ChainParams p{};
  p.iq = src; p.n_frames = 1; p.NTS = c.num_ADC_samples_per_chirp; p.PN = c.num_chirps_per_frame;
  p.n_rx = c.num_Rx_antennas; p.rx_sel = c.rx_select; p.ND = c.Doppler_fft_size;
  p.nts_fft = p.NTS < (uint32_t)NR ? p.NTS : (uint32_t)NR;
  p.win_tab_d = h->win_tab_d.as<double>(); p.tw_d = h->tw_d.as<double2>(); p.hfft_d = h->hfft_d.as<double2>();
  p.dop_win = h->dop_win.as<float>(); p.dop_win_d = h->dop_win_d.as<double>();
  const size_t bytes = (size_t)NR * c.Doppler_fft_size * 4;
  const bool dev = is_device_ptr(out_db);
  float* d_out = out_db;
  if (!dev) { CK(h->f32_stage.ensure(bytes), "alloc"); d_out = h->f32_stage.as<float>(); }
  CK(launch_range_doppler_map(p, 0, d_out, h->stream), "range-Doppler map kernel");
  if (!dev) CK(cudaMemcpyAsync(out_db, d_out, bytes, cudaMemcpyDeviceToHost, h->stream), "D2H map");
  CK(cudaStreamSynchronize(h->stream), "synchronize");
  return FMCW_OK;
}

fmcw_status fmcw_synth_frames(fmcw_handle* h, const double* tables, uint32_t n_scat, uint64_t seed, uint64_t frame0,
                              uint64_t n_frames, double sigma, double dc, double rx_step, int16_t* iq_out) {
  if (!h || !tables || !iq_out) return FMCW_ERR_POINTER;
  BusyGuard g(h);
  if (!g.ok) return FMCW_ERR_BUSY;
  cudaSetDevice(h->device);
  const fmcw_config& c = h->cfg;
  const double* d_tab = tables;
  if (!is_device_ptr(tables)) {
    CK(h->synth_tab.ensure((size_t)n_frames * n_scat * 4 * sizeof(double) + 16), "alloc scene tables");
    CK(cudaMemcpyAsync(h->synth_tab.p, tables, (size_t)n_frames * n_scat * 4 * sizeof(double), cudaMemcpyHostToDevice, h->stream),
       "H2D scene tables");
    d_tab = h->synth_tab.as<double>();
  }
  const size_t bytes = (size_t)n_frames * c.num_Rx_antennas * c.num_chirps_per_frame * c.num_ADC_samples_per_chirp * 4;
  const bool dev = is_device_ptr(iq_out);
  int16_t* d_out = iq_out;
  if (!dev) { CK(h->iq_stage.ensure(bytes), "alloc iq staging"); d_out = h->iq_stage.as<int16_t>(); }
  CK(launch_synth(d_tab, n_scat, seed, frame0, n_frames, c.num_Rx_antennas, c.num_chirps_per_frame,
                  c.num_ADC_samples_per_chirp, sigma, dc, rx_step, d_out, h->stream), "synth kernel");
  if (!dev) CK(cudaMemcpyAsync(iq_out, d_out, bytes, cudaMemcpyDeviceToHost, h->stream), "D2H iq");
  CK(cudaStreamSynchronize(h->stream), "synchronize");
  return FMCW_OK;
}

}  // extern "C"
