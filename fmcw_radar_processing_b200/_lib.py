"""ctypes binding of libfmcw_cuda.so (include/fmcw_cuda.h).  No CPU fallback: if the shared
library is missing, or the GPU is, the product path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfmcw_cuda.so")

FMCW_OK = 0
STATUS_NAMES = {0: "FMCW_OK", 1: "FMCW_ERR_CONFIG", 2: "FMCW_ERR_POINTER", 3: "FMCW_ERR_CUDA", 4: "FMCW_ERR_NCCL",
                5: "FMCW_ERR_OOM", 6: "FMCW_ERR_BUSY", 7: "FMCW_ERR_SIZE", 8: "FMCW_ERR_NO_DATA", 9: "FMCW_ERR_STATE"}
PEAK_STRONGEST, PEAK_FIRST = 0, 1
LAYOUT_TIME_MAJOR, LAYOUT_FREQ_MAJOR = 0, 1
OPT_ASYNC_HOST = 1
OPT_STFT_PRECISION = 2      # 0: tensor-core TF32 x 2 kernel (default), 1: float64 kernel
OPT_STFT_TILES_PER_CTA = 4  # n > 1: one CTA per n column tiles in the tensor-core STFT kernel (fleets of small recordings)
OPT_RUN_GRAPH = 3           # 1: whole fmcw_run calls on device buffers are recorded and replayed as one CUDA graph


class FmcwError(RuntimeError):
    def __init__(self, status, message=""):
        self.status = status
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")


class fmcw_config(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "struct_size", "num_Tx_antennas", "num_Rx_antennas", "num_ADC_samples_per_chirp", "num_chirps_per_frame",
        "range_fft_size", "Doppler_fft_size", "max_num_targets", "window_length", "overlap", "MAX_FREQ_BINS",
        "rx_select", "peak_mode", "reserved0")] + [(n, C.c_double) for n in (
        "frame_time", "PRT", "Bandwidth", "carrier_frequency", "sampling_frequency", "IF_scale", "range_threshold",
        "Doppler_threshold", "min_distance", "max_distance", "lambda_", "Hz_to_mps_constant", "R_max", "dist_per_bin",
        "fD_max", "fD_per_bin", "kaiser_beta", "adc_scale")]


class fmcw_frame_out(C.Structure):
    _fields_ = [("range_max_abs", C.c_void_p), ("detected", C.c_void_p), ("range_bin", C.c_void_p),
                ("range_mag", C.c_void_p), ("doppler_bin", C.c_void_p), ("doppler_row", C.c_void_p),
                ("slow_time_mag", C.c_void_p)]


class fmcw_stft_out(C.Structure):
    _fields_ = [("intensity", C.c_void_p), ("capacity_cols", C.c_uint64), ("ld_cols", C.c_uint64),
                ("layout", C.c_uint32), ("reserved0", C.c_uint32)]


class fmcw_run_info(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_frames", "n_detected", "L_local", "L_total", "sample_offset", "nfft",
                                          "ncol_total", "col_begin", "ncol_local")] + \
               [("n_dtft_bins", C.c_uint32), ("n_refined", C.c_uint32), ("pmax_raw", C.c_double)]


class fmcw_device_info(C.Structure):
    _fields_ = [("info", fmcw_run_info), ("status", C.c_int32), ("reserved0", C.c_int32)]


EXPORTS = {
    "fmcw_version": (C.c_char_p, []),
    "fmcw_status_string": (C.c_char_p, [C.c_int]),
    "fmcw_create": (C.c_int, [C.POINTER(fmcw_config), C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "fmcw_destroy": (None, [C.c_void_p]),
    "fmcw_last_error": (C.c_char_p, [C.c_void_p]),
    "fmcw_get_stream": (C.c_void_p, [C.c_void_p]),
    "fmcw_synchronize": (C.c_int, [C.c_void_p]),
    "fmcw_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "fmcw_get_info": (C.c_int, [C.c_void_p, C.POINTER(fmcw_run_info)]),
    "fmcw_set_info_target": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fmcw_get_timings": (C.c_int, [C.c_void_p, C.POINTER(C.c_float * 4)]),
    "fmcw_process_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(fmcw_frame_out)]),
    "fmcw_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(fmcw_frame_out), C.POINTER(fmcw_stft_out)]),
    "fmcw_stft_frames": (C.c_int, [C.c_void_p, C.POINTER(fmcw_stft_out)]),
    "fmcw_stft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(fmcw_stft_out)]),
    "fmcw_get_slow_time": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "fmcw_load_slow_time": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "fmcw_set_halo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "fmcw_stft_local_max": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]),
    "fmcw_stft_sharded": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_double, C.POINTER(fmcw_stft_out)]),
    "fmcw_shard_pack": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fmcw_shard_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "fmcw_shard_stft": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(fmcw_stft_out)]),
    "fmcw_mailbox_bytes": (C.c_uint64, []),
    "fmcw_mailbox_post_heads": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_uint64]),
    "fmcw_mailbox_plan": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_uint64]),
    "fmcw_mailbox_stft": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_uint64,
                                    C.POINTER(fmcw_stft_out)]),
    "fmcw_stft_axes": (C.c_int, [C.POINTER(fmcw_config), C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fmcw_stft_finegrid": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_uint32, C.c_void_p, C.c_uint64,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fmcw_range_doppler_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "fmcw_range_spectrum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]),
    "fmcw_json_append_f32": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int]),
    "fmcw_json_append_f64": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int]),
    "fmcw_synth_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double,
                                    C.c_double, C.c_double, C.c_void_p]),
}

_lib = None


def load():
    """Loads the in-tree shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FmcwError(3, f"{LIB_PATH} is missing: run `python -m fmcw_radar_processing_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
