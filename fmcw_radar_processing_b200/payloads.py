"""JSON payloads of the reference (the wire format the Next.js dashboard consumes), RP:302-436 and
RP:576-596, with MATLAB ``jsonencode`` conventions: matrices nest row-major, vectors (row or column)
flatten, 1x1 values are scalars, NaN/Inf become ``null``, key order is field order.

Writers stream to the file: at the reference's hop 1 the spectrogram payload is hundreds of MB.
"""
from __future__ import annotations

import io
import json
import math

import numpy as np

FMT64 = "%.15g"     # jsonencode prints doubles with 15 significant digits
FMT32 = "%.9g"      # values that are float32 in the library (round-trips a float32)


def _fmt_for(a: np.ndarray) -> str:
    return FMT32 if a.dtype == np.float32 else FMT64


def _num(v, fmt) -> str:
    v = float(v)
    if math.isnan(v) or math.isinf(v):
        return "null"
    if v == int(v) and abs(v) < 1e15:
        return str(int(v))
    return fmt % v


def _row(a: np.ndarray, fmt: str) -> str:
    if a.size == 0:
        return "[]"
    if np.isfinite(a).all():
        txt = np.char.mod(fmt, a)
        return "[" + ",".join(txt.tolist()) + "]"
    return "[" + ",".join(_num(v, fmt) for v in a) + "]"


def write_value(f: io.TextIOBase, v):
    """jsonencode(v) for str, scalars, vectors and 2-D matrices."""
    if isinstance(v, str):
        f.write(json.dumps(v))
        return
    a = np.asarray(v)
    if a.dtype.kind in "iub":
        a = a.astype(np.float64)
    fmt = _fmt_for(a)
    if a.ndim == 0 or a.size == 1 and a.ndim <= 2:
        f.write(_num(a.reshape(-1)[0], fmt))          # 1x1 -> scalar
    elif a.ndim == 1 or 1 in a.shape:
        f.write(_row(a.reshape(-1), fmt))             # row or column vector -> flat array
    else:
        f.write("[")
        for i in range(a.shape[0]):                   # row-major nesting
            if i:
                f.write(",")
            f.write(_row(a[i], fmt))
        f.write("]")


NATIVE_MIN_ELEMS = 1 << 16      # larger float matrices go through the native multi-threaded writer


def _native_append(path: str, a: np.ndarray) -> bool:
    """Appends jsonencode(a) with libfmcw_cuda's host-side writer (no GPU involved); False if unavailable."""
    try:
        from . import _lib
        lib = _lib.load()
    except Exception:
        return False
    if a.dtype not in (np.float32, np.float64) or a.ndim != 2:
        return False
    fn = lib.fmcw_json_append_f32 if a.dtype == np.float32 else lib.fmcw_json_append_f64
    es = a.itemsize
    if a.strides[0] % es or a.strides[1] % es:
        return False
    return fn(path.encode(), a.ctypes.data, a.shape[0], a.shape[1], a.strides[0] // es, a.strides[1] // es, 1) == 0


def write_struct(path: str, fields) -> None:
    """jsonencode(struct) with ``fields`` an ordered list of (name, value).  Large float matrices (the spectrogram,
    the range-FFT heat map) are streamed by the native writer, everything else by Python."""
    f = open(path, "w")
    try:
        f.write("{")
        for i, (k, v) in enumerate(fields):
            if i:
                f.write(",")
            f.write(json.dumps(k) + ":")
            a = v if isinstance(v, np.ndarray) else None
            if a is not None and a.ndim == 2 and a.size >= NATIVE_MIN_ELEMS and 1 not in a.shape:
                f.close()
                ok = _native_append(path, a)
                f = open(path, "a")
                if ok:
                    continue
            write_value(f, v)
        f.write("}")
    finally:
        f.close()


def matlab_growing_matrix(values: np.ndarray, detected: np.ndarray) -> np.ndarray:
    """The shape quirk of RP:157-159 + RP:245-250: arrays are allocated ``max_num_targets x frame_count``
    (1 x N) but written at ``(fr_idx, j)``, so MATLAB grows them to ``lastDetectedFrame x N`` with the data
    in column 1 and zeros elsewhere (frames without a target stay 0)."""
    N = len(values)
    det = np.flatnonzero(detected)
    rows = max(1, int(det[-1]) + 1) if det.size else 1
    M = np.zeros((rows, N), dtype=np.float64)
    if det.size:
        M[det, 0] = np.asarray(values, dtype=np.float64)[det]
    return M


def spectrogram_payload(T, frequency, intensity):
    """RP:307-312; intensity is 1024 x ncol."""
    return [("time", T), ("frequency", frequency), ("intensity", intensity),
            ("title", "All Frames - Log-Scaled Spectrogram"), ("xLabel", "Time (s)"), ("yLabel", "Frequency (Hz)")]


def range_fft_payload(frame_count, array_bin_range, range_tx1rx1_max_abs, filename):
    """RP:355-361; range_tx1rx1_max_abs is 256 x N."""
    return [("time_axis", np.arange(frame_count) * 0.15), ("array_bin_range", array_bin_range),
            ("range_tx1rx1_max_abs", range_tx1rx1_max_abs), ("filename", filename)]


def range_speed_payload(frame_count, rng, speed, detected, filename):
    """RP:379-389 with the growing-matrix shape of RP:245-250."""
    return [("time_axis", np.arange(frame_count) * 0.15), ("range", matlab_growing_matrix(rng, detected)),
            ("speed", matlab_growing_matrix(speed, detected)), ("filename", filename)]


def fft_payload(magnitude, filename, fr_idx=100):
    """RP:410-422."""
    return [("range_bins", np.arange(len(magnitude))), ("magnitude", magnitude), ("frame_index", fr_idx), ("filename", filename)]


def batch_spectrogram_payload(T, frequency, intensity, batch, start_frame, end_frame, filename):
    """RP:576-584."""
    return [("time", T), ("frequency", frequency), ("intensity", intensity),
            ("title", f"Spectrogram - Batch {batch}"), ("xLabel", "Time (s) (relative to detected activity)"),
            ("yLabel", "Frequency (Hz)"), ("start_frame", start_frame), ("end_frame", end_frame), ("filename_base", filename)]
