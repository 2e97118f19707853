// MEX gateway of libfmcw_cuda for the reference's radar_processing.m.
//
//   out = fmcw_cuda_mex('run', iq, calib_data, cfg)          % replaces RP:197-261 + RP:265-299
//   out = fmcw_cuda_mex('frames', iq, calib_data, cfg)       % RP:197-261 only ('yes' branch, RP:457-530)
//   out = fmcw_cuda_mex('stft_frames', [], calib_data, cfg)   % RP:538-566 on the signal of the last 'frames' call (float64 inside)
//   out = fmcw_cuda_mex('stft', x, calib_data, cfg)          % RP:270-299 / RP:538-566 on a given signal
//
// iq    int16 [2 x NTS x PN x RX x N]  (MATLAB column-major == C [frame][rx][chirp][sample][I,Q])
// cfg   struct with the fmcw_configurations field names (RP:645-672) plus kaiser_beta, MAX_FREQ_BINS,
//       adc_scale, rx_select (1-based), peak_mode (0 strongest / 1 first = default, the vendor picker's order)
// out   struct: detected, range_idx (1-based), range_mag, doppler_idx (1-based), range_max_abs [256 x N],
//       doppler_row [ND x N] complex, slow_time_mag [PN x N], T [1 x ncol], frequency [1 x 1024],
//       intensity [1024 x ncol] single, nfft, pmax
//
// Build (on a host that has MATLAB): mex -I../include fmcw_cuda_mex.cpp -L../fmcw_radar_processing_b200 -lfmcw_cuda
// This container has no MATLAB: the file is syntax-checked against mex/stub/mex.h only (tests/test_gateways.py).
//
// MATLAB errors unwind with longjmp semantics, so every C++ object is released before mexErrMsgIdAndTxt.
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "mex.h"
#include "fmcw_cuda.h"

namespace {

fmcw_handle* g_handle = nullptr;
fmcw_config g_cfg;
bool g_have_cfg = false;

bool g_locked = false;
std::vector<std::pair<int, int64_t>> g_opts;     // fmcw_cuda_mex('option', id, value): applied to every handle

// Destroys the handle and releases the lock taken when it was created, so that `clear fmcw_cuda_mex` works again once no
// handle is alive (one mexLock per live handle, never more: mexLock counts).
void release_handle() {
  if (g_handle) fmcw_destroy(g_handle);
  g_handle = nullptr;
  g_have_cfg = false;
  if (g_locked) { mexUnlock(); g_locked = false; }
}

void at_exit() { release_handle(); }

double field(const mxArray* s, const char* name, double dflt, bool required, std::string& missing) {
  const mxArray* f = mxGetField(s, 0, name);
  if (!f || mxIsEmpty(f)) {
    if (required && missing.empty()) missing = name;
    return dflt;
  }
  return mxGetScalar(f);
}

bool read_config(const mxArray* s, fmcw_config& c, std::string& missing) {
  std::memset(&c, 0, sizeof(c));
  c.struct_size = sizeof(c);
#define U(name, req, d) c.name = (uint32_t)field(s, #name, d, req, missing)
#define D(name, req, d) c.name = field(s, #name, d, req, missing)
  U(num_Tx_antennas, false, 1); U(num_Rx_antennas, true, 1); U(num_ADC_samples_per_chirp, true, 0);
  U(num_chirps_per_frame, true, 0); U(range_fft_size, true, 256); U(Doppler_fft_size, true, 16);
  U(max_num_targets, true, 1); U(window_length, true, 20); U(overlap, true, 19); U(MAX_FREQ_BINS, false, 1024);
  c.rx_select = (uint32_t)field(s, "rx_select", 1, false, missing) - 1;   // MATLAB 1-based
  U(peak_mode, false, 1);
  D(frame_time, false, 0.15); D(PRT, true, 0); D(Bandwidth, false, 0); D(carrier_frequency, false, 0);
  D(sampling_frequency, false, 0); D(IF_scale, true, 0); D(range_threshold, true, 200); D(Doppler_threshold, true, 50);
  D(min_distance, true, 0.9); D(max_distance, true, 25.0); c.lambda = field(s, "lambda", 0, false, missing);
  D(Hz_to_mps_constant, false, 0); D(R_max, false, 0); D(dist_per_bin, true, 0); D(fD_max, false, 0); D(fD_per_bin, false, 0);
  D(kaiser_beta, false, 3.0); D(adc_scale, false, 4095.0);
#undef U
#undef D
  return missing.empty();
}

}  // namespace

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  std::string err_id, err_msg;   // filled instead of throwing; reported after all C++ state is consistent
  do {
    if (nrhs < 1 || !mxIsChar(prhs[0])) { err_id = "fmcw:usage"; err_msg = "out = fmcw_cuda_mex(cmd, data, calib_data, cfg)"; break; }
    char cmd[16] = {0};
    mxGetString(prhs[0], cmd, sizeof(cmd));
    if (std::strcmp(cmd, "release") == 0) { release_handle(); plhs[0] = mxCreateDoubleScalar(0.0); break; }   // fmcw_cuda_mex('release'): frees the GPU state, unlocks the MEX file
    if (std::strcmp(cmd, "option") == 0) {
      // fmcw_cuda_mex('option', id, value): fmcw_set_option on the current handle and on every handle created later
      // (id 2 = FMCW_OPT_STFT_PRECISION: 0 tensor-core kernel, 1 float64 kernel; see include/fmcw_cuda.h)
      if (nrhs != 3 || !mxIsNumeric(prhs[1]) || !mxIsNumeric(prhs[2]) || mxGetNumberOfElements(prhs[1]) != 1 || mxGetNumberOfElements(prhs[2]) != 1) {
        err_id = "fmcw:usage"; err_msg = "fmcw_cuda_mex('option', id, value)"; break;
      }
      const int id = (int)mxGetScalar(prhs[1]);
      const int64_t value = (int64_t)mxGetScalar(prhs[2]);
      if (g_handle) {
        const fmcw_status st = fmcw_set_option(g_handle, id, value);
        if (st != FMCW_OK) { err_id = "fmcw:option"; err_msg = fmcw_last_error(g_handle); break; }
      }
      bool found = false;
      for (auto& o : g_opts) if (o.first == id) { o.second = value; found = true; }
      if (!found) g_opts.push_back(std::make_pair(id, value));
      plhs[0] = mxCreateDoubleScalar(0.0);
      break;
    }
    if (nrhs != 4 || !mxIsStruct(prhs[3])) { err_id = "fmcw:usage"; err_msg = "out = fmcw_cuda_mex(cmd, data, calib_data, cfg)"; break; }
    fmcw_config c;
    std::string missing;
    if (!read_config(prhs[3], c, missing)) { err_id = "fmcw:config"; err_msg = "cfg is missing field " + missing; break; }
    if (!g_handle || !g_have_cfg || std::memcmp(&c, &g_cfg, sizeof(c)) != 0) {
      release_handle();
      const double* cal = mxIsDouble(prhs[2]) ? mxGetPr(prhs[2]) : nullptr;
      fmcw_status st = fmcw_create(&c, cal, cal ? (uint64_t)mxGetNumberOfElements(prhs[2]) : 0, 0, &g_handle);
      if (st != FMCW_OK) { err_id = "fmcw:create"; err_msg = fmcw_status_string(st); g_handle = nullptr; break; }
      g_cfg = c; g_have_cfg = true;
      bool opt_ok = true;
      for (const auto& o : g_opts) if (fmcw_set_option(g_handle, o.first, o.second) != FMCW_OK) opt_ok = false;
      if (!opt_ok) { err_id = "fmcw:option"; err_msg = fmcw_last_error(g_handle); break; }
      if (!g_locked) { mexLock(); g_locked = true; }
      mexAtExit(at_exit);
    }
    const uint32_t NTS = c.num_ADC_samples_per_chirp, PN = c.num_chirps_per_frame, ND = c.Doppler_fft_size, NQ = c.MAX_FREQ_BINS;
    const bool is_stft = std::strcmp(cmd, "stft") == 0, is_run = std::strcmp(cmd, "run") == 0;
    const bool is_stft_frames = std::strcmp(cmd, "stft_frames") == 0;
    if (!is_stft && !is_run && !is_stft_frames && std::strcmp(cmd, "frames") != 0) { err_id = "fmcw:usage"; err_msg = "cmd must be 'run', 'frames', 'stft_frames' or 'stft'"; break; }
    const char* names[] = {"detected", "range_idx", "range_mag", "doppler_idx", "range_max_abs", "doppler_row", "slow_time_mag",
                           "T", "frequency", "intensity", "nfft", "pmax"};
    plhs[0] = mxCreateStructMatrix(1, 1, 12, names);
    uint64_t L = 0, n = 0;
    fmcw_status st = FMCW_OK;
    mxArray* inten = nullptr;
    if (is_stft_frames) {
      fmcw_run_info fi;
      st = fmcw_get_info(g_handle, &fi);
      if (st != FMCW_OK) { err_id = "fmcw:info"; err_msg = fmcw_last_error(g_handle); break; }
      L = fi.L_local;
      const uint64_t cap = L > c.overlap ? (L - c.overlap) / (c.window_length - c.overlap) : 1;
      inten = mxCreateNumericMatrix(NQ, cap ? cap : 1, mxSINGLE_CLASS, mxREAL);
      fmcw_stft_out so = {(float*)mxGetData(inten), cap ? cap : 1, 0, FMCW_LAYOUT_TIME_MAJOR, 0};
      st = fmcw_stft_frames(g_handle, &so);
    } else if (is_stft) {
      if (!mxIsSingle(prhs[1])) { err_id = "fmcw:type"; err_msg = "stft input must be single"; break; }
      L = (uint64_t)mxGetNumberOfElements(prhs[1]);
      const uint64_t cap = L > c.overlap ? (L - c.overlap) / (c.window_length - c.overlap) : 1;
      inten = mxCreateNumericMatrix(NQ, cap ? cap : 1, mxSINGLE_CLASS, mxREAL);
      fmcw_stft_out so = {(float*)mxGetData(inten), cap ? cap : 1, 0, FMCW_LAYOUT_TIME_MAJOR, 0};
      st = fmcw_stft(g_handle, (const float*)mxGetData(prhs[1]), L, &so);
    } else {
      if (!mxIsInt16(prhs[1])) { err_id = "fmcw:type"; err_msg = "iq must be int16 [2 x NTS x PN x RX x N]"; break; }
      const uint64_t per_frame = 2ull * NTS * PN * c.num_Rx_antennas;
      n = (uint64_t)mxGetNumberOfElements(prhs[1]) / per_frame;
      mxArray* det = mxCreateNumericMatrix(1, n, mxINT32_CLASS, mxREAL);
      mxArray* rbin = mxCreateNumericMatrix(1, n, mxINT32_CLASS, mxREAL);
      mxArray* rmag = mxCreateNumericMatrix(1, n, mxSINGLE_CLASS, mxREAL);
      mxArray* dbin = mxCreateNumericMatrix(1, n, mxINT32_CLASS, mxREAL);
      mxArray* rmax = mxCreateNumericMatrix(c.range_fft_size, n, mxSINGLE_CLASS, mxREAL);       // 256 x N column-major == [frame][256]
      mxArray* drow = mxCreateNumericMatrix(2 * ND, n, mxSINGLE_CLASS, mxREAL);                   // interleaved re/im, split in the .m wrapper
      mxArray* slow = mxCreateNumericMatrix(PN, n, mxSINGLE_CLASS, mxREAL);
      fmcw_frame_out fo = {(float*)mxGetData(rmax), (int32_t*)mxGetData(det), (int32_t*)mxGetData(rbin), (float*)mxGetData(rmag),
                           (int32_t*)mxGetData(dbin), (float*)mxGetData(drow), (float*)mxGetData(slow)};
      if (is_run) {
        const uint64_t Lmax = n * PN;
        const uint64_t cap = Lmax > c.overlap ? (Lmax - c.overlap) / (c.window_length - c.overlap) : 1;
        inten = mxCreateNumericMatrix(NQ, cap ? cap : 1, mxSINGLE_CLASS, mxREAL);
        fmcw_stft_out so = {(float*)mxGetData(inten), cap ? cap : 1, 0, FMCW_LAYOUT_TIME_MAJOR, 0};
        st = fmcw_run(g_handle, (const int16_t*)mxGetData(prhs[1]), n, &fo, &so);
      } else {
        st = fmcw_process_frames(g_handle, (const int16_t*)mxGetData(prhs[1]), n, &fo);
      }
      // MATLAB indices are 1-based (RP:211, 233)
      int32_t* rb = (int32_t*)mxGetData(rbin); int32_t* db = (int32_t*)mxGetData(dbin);
      for (uint64_t i = 0; i < n; ++i) { rb[i] += 1; db[i] += 1; }
      mxSetField(plhs[0], 0, "detected", det); mxSetField(plhs[0], 0, "range_idx", rbin); mxSetField(plhs[0], 0, "range_mag", rmag);
      mxSetField(plhs[0], 0, "doppler_idx", dbin); mxSetField(plhs[0], 0, "range_max_abs", rmax);
      mxSetField(plhs[0], 0, "doppler_row", drow); mxSetField(plhs[0], 0, "slow_time_mag", slow);
    }
    if (st != FMCW_OK) { err_id = "fmcw:run"; err_msg = std::string(fmcw_status_string(st)) + ": " + fmcw_last_error(g_handle); break; }
    if (inten) {
      fmcw_run_info info;
      st = fmcw_get_info(g_handle, &info);
      if (st != FMCW_OK) { err_id = "fmcw:info"; err_msg = fmcw_last_error(g_handle); break; }
      const uint64_t ncol = info.ncol_local;
      mxSetN(inten, ncol);                                   // 1024 x ncol, no copy: time-major == column-major
      mxArray* T = mxCreateDoubleMatrix(1, ncol, mxREAL);
      mxArray* F = mxCreateDoubleMatrix(1, NQ, mxREAL);
      uint64_t nfft = 0, nct = 0;
      fmcw_stft_axes(&c, info.L_total, 0, ncol, mxGetPr(T), mxGetPr(F), &nfft, &nct);
      mxSetField(plhs[0], 0, "T", T); mxSetField(plhs[0], 0, "frequency", F); mxSetField(plhs[0], 0, "intensity", inten);
      mxSetField(plhs[0], 0, "nfft", mxCreateDoubleScalar((double)nfft));
      mxSetField(plhs[0], 0, "pmax", mxCreateDoubleScalar(info.pmax_raw));
    }
  } while (false);
  if (!err_id.empty()) {
    static char id[64], msg[512];                            // static: survive the longjmp of mexErrMsgIdAndTxt
    std::strncpy(id, err_id.c_str(), sizeof(id) - 1);
    std::strncpy(msg, err_msg.c_str(), sizeof(msg) - 1);
    err_id.clear(); err_id.shrink_to_fit(); err_msg.clear(); err_msg.shrink_to_fit();
    mexErrMsgIdAndTxt(id, "%s", msg);
  }
  (void)nlhs;
}
