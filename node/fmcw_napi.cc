// Node-API addon over libfmcw_cuda for the JS side (dashboard / ingestion service).
//
//   const fmcw = require('./build/Release/fmcw_napi');
//   const h = fmcw.create(cfgObject, calibFloat64Array);                 // fmcw_create
//   const out = await fmcw.run(h, iqInt16Array, nFrames);                // fmcw_run on a libuv pool thread
//   // out: { detected:Int32Array, rangeBin:Int32Array, rangeMag:Float32Array, dopplerBin:Int32Array,
//   //        rangeMaxAbs:Float32Array(256*n), time:Float64Array, frequency:Float64Array, intensity:Float32Array(1024*ncol), ncol }
//   fmcw.destroy(h);
//
// The GPU call runs in napi_create_async_work's execute callback (off the JS thread; handles are usable from
// any thread, one call at a time); results are library-filled ArrayBuffers; failures reject the promise with
// the status string.  No Node toolchain exists in this container: the file is syntax-checked against
// node/stub/node_api.h only (tests/test_gateways.py).  The reference branch has no JS caller of its own
// (src/app/page.js is the create-next-app template).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "node_api.h"
#include "fmcw_cuda.h"

namespace {

struct Job {
  fmcw_handle* h; fmcw_config cfg;
  const int16_t* iq; uint64_t n;
  std::vector<int32_t> det, rbin, dbin; std::vector<float> rmag, rmax, inten; std::vector<double> T, F;
  uint64_t ncol = 0; fmcw_status st = FMCW_OK; std::string err;
  napi_deferred deferred; napi_async_work work; napi_ref iq_ref;
};

double num(napi_env env, napi_value obj, const char* key, double dflt) {
  napi_value v; bool has = false;
  if (napi_has_named_property(env, obj, key, &has) != napi_ok || !has) return dflt;
  double d = dflt;
  napi_get_named_property(env, obj, key, &v);
  napi_get_value_double(env, v, &d);
  return d;
}

void fill_config(napi_env env, napi_value o, fmcw_config& c) {
  std::memset(&c, 0, sizeof(c));
  c.struct_size = sizeof(c);
#define U(n, d) c.n = (uint32_t)num(env, o, #n, d)
#define D(n, d) c.n = num(env, o, #n, d)
  U(num_Tx_antennas, 1); U(num_Rx_antennas, 1); U(num_ADC_samples_per_chirp, 0); U(num_chirps_per_frame, 0);
  U(range_fft_size, 256); U(Doppler_fft_size, 16); U(max_num_targets, 1); U(window_length, 20); U(overlap, 19);
  U(MAX_FREQ_BINS, 1024); U(peak_mode, 1);
  c.rx_select = (uint32_t)num(env, o, "rx_select", 1) - 1;
  D(frame_time, 0.15); D(PRT, 0); D(Bandwidth, 0); D(carrier_frequency, 0); D(sampling_frequency, 0); D(IF_scale, 0);
  D(range_threshold, 200); D(Doppler_threshold, 50); D(min_distance, 0.9); D(max_distance, 25.0);
  c.lambda = num(env, o, "lambda", 0);
  D(Hz_to_mps_constant, 0); D(R_max, 0); D(dist_per_bin, 0); D(fD_max, 0); D(fD_per_bin, 0); D(kaiser_beta, 3.0); D(adc_scale, 4095.0);
#undef U
#undef D
}

struct Wrapped { fmcw_handle* h; fmcw_config cfg; };

napi_value Create(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  if (argc < 1) { napi_throw_error(env, "FMCW_USAGE", "create(cfg [, calibFloat64Array])"); return nullptr; }
  Wrapped* w = new Wrapped();
  fill_config(env, argv[0], w->cfg);
  void* cal = nullptr; size_t cal_len = 0; napi_typedarray_type ty; napi_value ab; size_t off;
  if (argc > 1) {
    // the calibration vector is read as doubles: anything but a Float64Array is rejected, not reinterpreted
    if (napi_get_typedarray_info(env, argv[1], &ty, &cal_len, &cal, &ab, &off) != napi_ok || ty != napi_float64_array) {
      delete w;
      napi_throw_error(env, "FMCW_TYPE", "calibration data must be a Float64Array");
      return nullptr;
    }
  }
  fmcw_status st = fmcw_create(&w->cfg, (const double*)cal, cal_len, 0, &w->h);
  if (st != FMCW_OK) { delete w; napi_throw_error(env, "FMCW_CREATE", fmcw_status_string(st)); return nullptr; }
  napi_value ext;
  napi_create_external(env, w, [](napi_env, void* d, void*) { Wrapped* x = (Wrapped*)d; if (x->h) fmcw_destroy(x->h); delete x; }, nullptr, &ext);
  return ext;
}

napi_value Destroy(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  Wrapped* w = nullptr;
  napi_get_value_external(env, argv[0], (void**)&w);
  if (w && w->h) { fmcw_destroy(w->h); w->h = nullptr; }
  return nullptr;
}

// setOption(handle, id, value): fmcw_set_option (id 2 = FMCW_OPT_STFT_PRECISION: 0 tensor-core kernel, 1 float64 kernel)
napi_value SetOption(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  if (argc < 3) { napi_throw_error(env, "FMCW_USAGE", "setOption(handle, id, value)"); return nullptr; }
  Wrapped* w = nullptr;
  if (napi_get_value_external(env, argv[0], (void**)&w) != napi_ok || !w || !w->h) {
    napi_throw_error(env, "FMCW_HANDLE", "setOption: not a live handle"); return nullptr;
  }
  int32_t id = 0; int64_t value = 0;
  if (napi_get_value_int32(env, argv[1], &id) != napi_ok || napi_get_value_int64(env, argv[2], &value) != napi_ok) {
    napi_throw_error(env, "FMCW_TYPE", "setOption: id and value must be numbers"); return nullptr;
  }
  const fmcw_status st = fmcw_set_option(w->h, id, value);
  if (st != FMCW_OK) { napi_throw_error(env, "FMCW_OPTION", fmcw_last_error(w->h)); return nullptr; }
  return nullptr;
}

void Execute(napi_env, void* data) {   // libuv pool thread: no N-API calls here
  Job* j = (Job*)data;
  const fmcw_config& c = j->cfg;
  const uint64_t n = j->n, PN = c.num_chirps_per_frame;
  j->det.resize(n); j->rbin.resize(n); j->dbin.resize(n); j->rmag.resize(n); j->rmax.resize(n * c.range_fft_size);
  const uint64_t Lmax = n * PN;
  const uint64_t cap = Lmax > c.overlap ? (Lmax - c.overlap) / (c.window_length - c.overlap) : 1;
  j->inten.resize((cap ? cap : 1) * (uint64_t)c.MAX_FREQ_BINS);
  fmcw_frame_out fo = {j->rmax.data(), j->det.data(), j->rbin.data(), j->rmag.data(), j->dbin.data(), nullptr, nullptr};
  fmcw_stft_out so = {j->inten.data(), cap ? cap : 1, 0, FMCW_LAYOUT_TIME_MAJOR, 0};
  j->st = fmcw_run(j->h, j->iq, n, &fo, &so);
  if (j->st != FMCW_OK) { j->err = std::string(fmcw_status_string(j->st)) + ": " + fmcw_last_error(j->h); return; }
  fmcw_run_info info;
  j->st = fmcw_get_info(j->h, &info);
  if (j->st != FMCW_OK) { j->err = fmcw_last_error(j->h); return; }
  j->ncol = info.ncol_local;
  j->T.resize(j->ncol); j->F.resize(c.MAX_FREQ_BINS);
  uint64_t nfft, nct;
  fmcw_stft_axes(&c, info.L_total, 0, j->ncol, j->T.data(), j->F.data(), &nfft, &nct);
}

template <class T>
napi_value typed(napi_env env, napi_typedarray_type ty, const T* src, size_t count) {
  void* dst; napi_value ab, ta;
  napi_create_arraybuffer(env, count * sizeof(T), &dst, &ab);
  if (count) std::memcpy(dst, src, count * sizeof(T));
  napi_create_typedarray(env, ty, count, ab, 0, &ta);
  return ta;
}

void Complete(napi_env env, napi_status, void* data) {   // back on the JS thread
  Job* j = (Job*)data;
  if (j->st != FMCW_OK) {
    napi_value msg, err;
    napi_create_string_utf8(env, j->err.c_str(), j->err.size(), &msg);
    napi_create_error(env, nullptr, msg, &err);
    napi_reject_deferred(env, j->deferred, err);
  } else {
    napi_value o, v;
    napi_create_object(env, &o);
    napi_set_named_property(env, o, "detected", typed(env, napi_int32_array, j->det.data(), j->det.size()));
    napi_set_named_property(env, o, "rangeBin", typed(env, napi_int32_array, j->rbin.data(), j->rbin.size()));
    napi_set_named_property(env, o, "dopplerBin", typed(env, napi_int32_array, j->dbin.data(), j->dbin.size()));
    napi_set_named_property(env, o, "rangeMag", typed(env, napi_float32_array, j->rmag.data(), j->rmag.size()));
    napi_set_named_property(env, o, "rangeMaxAbs", typed(env, napi_float32_array, j->rmax.data(), j->rmax.size()));
    napi_set_named_property(env, o, "time", typed(env, napi_float64_array, j->T.data(), j->T.size()));
    napi_set_named_property(env, o, "frequency", typed(env, napi_float64_array, j->F.data(), j->F.size()));
    napi_set_named_property(env, o, "intensity", typed(env, napi_float32_array, j->inten.data(), j->ncol * j->cfg.MAX_FREQ_BINS));
    napi_create_double(env, (double)j->ncol, &v);
    napi_set_named_property(env, o, "ncol", v);
    napi_resolve_deferred(env, j->deferred, o);
  }
  napi_delete_reference(env, j->iq_ref);
  napi_delete_async_work(env, j->work);
  delete j;
}

napi_value Run(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  if (argc < 3) { napi_throw_error(env, "FMCW_USAGE", "run(handle, iqInt16Array, nFrames)"); return nullptr; }
  Wrapped* w = nullptr;
  if (napi_get_value_external(env, argv[0], (void**)&w) != napi_ok || !w || !w->h) {
    napi_throw_error(env, "FMCW_HANDLE", "invalid or destroyed handle");
    return nullptr;
  }
  void* data = nullptr; size_t len = 0; napi_typedarray_type ty; napi_value ab; size_t off;
  if (napi_get_typedarray_info(env, argv[1], &ty, &len, &data, &ab, &off) != napi_ok || ty != napi_int16_array) {
    napi_throw_error(env, "FMCW_TYPE", "iq must be an Int16Array [frame][rx][chirp][sample][I,Q]");
    return nullptr;
  }
  double n = 0;
  if (napi_get_value_double(env, argv[2], &n) != napi_ok || !(n >= 0) || n != (double)(uint64_t)n) {
    napi_throw_error(env, "FMCW_SIZE", "nFrames must be a non-negative integer");
    return nullptr;
  }
  // fmcw_run reads n * RX * PN * NTS (I,Q) pairs: never past the ArrayBuffer
  const fmcw_config& wc = w->cfg;
  const double per_frame = 2.0 * wc.num_Rx_antennas * wc.num_chirps_per_frame * wc.num_ADC_samples_per_chirp;
  if (n * per_frame > (double)len) {
    napi_throw_error(env, "FMCW_SIZE", "iq holds fewer than nFrames frames of the configured shape");
    return nullptr;
  }
  Job* j = new Job();
  j->h = w->h; j->cfg = w->cfg;
  j->iq = (const int16_t*)data;
  j->n = (uint64_t)n;
  napi_create_reference(env, argv[1], 1, &j->iq_ref);     // keep the input alive while the pool thread reads it
  napi_value promise, name;
  napi_create_promise(env, &j->deferred, &promise);
  napi_create_string_utf8(env, "fmcw_run", NAPI_AUTO_LENGTH, &name);
  napi_create_async_work(env, nullptr, name, Execute, Complete, j, &j->work);
  napi_queue_async_work(env, j->work);
  return promise;
}

napi_value Init(napi_env env, napi_value exports) {
  napi_value f;
  napi_create_function(env, "create", NAPI_AUTO_LENGTH, Create, nullptr, &f); napi_set_named_property(env, exports, "create", f);
  napi_create_function(env, "destroy", NAPI_AUTO_LENGTH, Destroy, nullptr, &f); napi_set_named_property(env, exports, "destroy", f);
  napi_create_function(env, "run", NAPI_AUTO_LENGTH, Run, nullptr, &f); napi_set_named_property(env, exports, "run", f);
  napi_create_function(env, "setOption", NAPI_AUTO_LENGTH, SetOption, nullptr, &f); napi_set_named_property(env, exports, "setOption", f);
  return exports;
}

}  // namespace

NAPI_MODULE(fmcw_napi, Init)
