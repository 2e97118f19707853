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

# ONE number format everywhere: the shortest text that round-trips the stored type (float64, or float32 for the arrays the
# library keeps in float32), which is what std::to_chars writes in the native writer and what current MATLAB releases'
# jsonencode prints for doubles (0.1 + 0.2 -> 0.30000000000000004).  Every numeric array goes through the native writer
# (csrc/json_writer.cu, host-only, multi-threaded); the Python formatter below is used for 1 x 1 values and only stands in
# for arrays when libfmcw_cuda.so is not built.


def _num(v) -> str:
    """A scalar: integers print without a fraction, everything else in shortest round-trip form."""
    f32 = isinstance(v, np.float32)
    x = float(v)
    if math.isnan(x) or math.isinf(x):
        return "null"
    if x == int(x) and abs(x) < 1e15:
        return str(int(x))
    return str(v) if f32 else repr(x)


def _row(a: np.ndarray) -> str:
    return "[" + ",".join(_num(v) for v in a) + "]"


def write_value(f: io.TextIOBase, v):
    """jsonencode(v) for str, scalars, vectors and 2-D matrices, in Python."""
    if isinstance(v, str):
        f.write(json.dumps(v))
        return
    a = np.asarray(v)
    if a.dtype.kind in "iub":
        a = a.astype(np.float64)
    if a.ndim == 0 or a.size == 1 and a.ndim <= 2:
        f.write(_num(a.reshape(-1)[0]))               # 1x1 -> scalar
    elif a.ndim == 1 or 1 in a.shape:
        f.write(_row(a.reshape(-1)))                  # row or column vector -> flat array
    else:
        f.write("[" + ",".join(_row(a[i]) for i in range(a.shape[0])) + "]")     # row-major nesting


USE_NATIVE = True               # tests switch it off to compare the two writers


def _native_append(path: str, a: np.ndarray) -> bool:
    """Appends jsonencode(a) with libfmcw_cuda's host-side writer (no GPU involved); False if unavailable."""
    if not USE_NATIVE:
        return False
    try:
        from . import _lib
        lib = _lib.load()
    except Exception:
        return False
    if a.dtype.kind in "iub":
        a = a.astype(np.float64)
    if a.dtype not in (np.float32, np.float64) or a.ndim > 2 or a.size <= 1:
        return False
    if a.ndim == 1:
        a = a[None, :]
    fn = lib.fmcw_json_append_f32 if a.dtype == np.float32 else lib.fmcw_json_append_f64
    es = a.itemsize
    if a.strides[0] % es or a.strides[1] % es:
        a = np.ascontiguousarray(a)
    return fn(path.encode(), a.ctypes.data, a.shape[0], a.shape[1], a.strides[0] // es, a.strides[1] // es, 1) == 0


def write_struct(path: str, fields) -> None:
    """jsonencode(struct) with ``fields`` an ordered list of (name, value); numeric arrays are streamed by the native
    writer (the spectrogram payload is gigabytes of text at the reference's hop 1)."""
    f = open(path, "w")
    try:
        f.write("{")
        for i, (k, v) in enumerate(fields):
            if i:
                f.write(",")
            f.write(json.dumps(k) + ":")
            if isinstance(v, (np.ndarray, list, tuple)) and not isinstance(v, str):
                a = np.asarray(v)
                if a.dtype.kind in "fiub" and a.size > 1:
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


def _jet(x):
    """MATLAB's jet colormap at x in [0, 1] -> uint8 RGB."""
    r = np.clip(np.minimum(4 * x - 1.5, -4 * x + 4.5), 0, 1)
    g = np.clip(np.minimum(4 * x - 0.5, -4 * x + 3.5), 0, 1)
    b = np.clip(np.minimum(4 * x + 0.5, -4 * x + 2.5), 0, 1)
    return (np.stack([r, g, b], axis=-1) * 255 + 0.5).astype(np.uint8)


def write_spectrogram_png(path, psd_band, clim=(-40.0, 0.0), max_width=3600):
    """spectrogram.png (RP:332-348): ``psd_band`` [ncol][n_rows] is psd on the fine-grid rows between ylim [0 150] Hz
    (fmcw_stft_finegrid); view(0, 90), axis off, colormap(jet), clim [-40 0].  Frequency runs upwards, time to the right;
    at most ``max_width`` columns (6 in x 600 dpi, RP:344).  A plain 8-bit RGB PNG written with zlib -- the rendering of
    MATLAB's surf / exportgraphics is not reproduced pixel for pixel (out of scope)."""
    import struct
    import zlib
    a = np.asarray(psd_band, dtype=np.float32)
    step = max(1, -(-a.shape[0] // max_width))
    img = a[::step].T[::-1]                                    # rows = frequency (top = f_hi), columns = time
    x = (np.nan_to_num(img, nan=clim[0], neginf=clim[0], posinf=clim[1]) - clim[0]) / (clim[1] - clim[0])
    rgb = _jet(np.clip(x, 0, 1))
    h, w, _ = rgb.shape
    raw = np.concatenate([np.zeros((h, 1), dtype=np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
    return w, h
