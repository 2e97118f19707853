"""CPU tests of the host-side mirror: configuration (RP:89-179), recording container, payload writers
(jsonencode conventions) and main()'s error behaviour (RPA:25-66)."""
import json
import os

import numpy as np
import pytest

from fmcw_radar_processing_b200 import payloads as P
from fmcw_radar_processing_b200 import synth
from fmcw_radar_processing_b200.config import fmcw_configurations, to_c_config
from fmcw_radar_processing_b200.parse import f_parse_data2, make_sxml, write_recording
from oracle import fmcw_oracle as O


def test_configuration_matches_oracle_and_reference_names():
    sx = make_sxml(numSamplesPerChirp=64, numChirpsPerFrame=16, numAntennasRx=2)
    cfg = fmcw_configurations(sx)
    o = O.configure(sx)
    names = ["frame_time", "PRT", "Bandwidth", "num_Tx_antennas", "num_Rx_antennas", "carrier_frequency",
             "num_ADC_samples_per_chirp", "num_chirps_per_frame", "sampling_frequency", "range_fft_size",
             "Doppler_fft_size", "IF_scale", "range_threshold", "Doppler_threshold", "min_distance", "max_distance",
             "max_num_targets", "lambda", "Hz_to_mps_constant", "R_max", "dist_per_bin", "fD_max", "fD_per_bin",
             "window_length", "max_slider_index", "overlap"]                       # RP:645-672, in order
    assert list(cfg)[:len(names)] == names
    for n in names:
        if n == "max_slider_index":
            continue
        assert cfg[n] == getattr(o, "lambda_" if n == "lambda" else n), n
    c = to_c_config(cfg)
    assert c.rx_select == 0 and c.window_length == 20 and c.overlap == 19 and c.MAX_FREQ_BINS == 1024


def test_recording_round_trip(tmp_path):
    sx = make_sxml(numSamplesPerChirp=64, numChirpsPerFrame=16, numAntennasRx=2)
    iq = np.random.default_rng(0).integers(0, 4096, size=(5, 2, 16, 64, 2)).astype(np.int16)
    calib = synth.default_calib(2, 64)
    base = str(tmp_path / "radar_data")
    write_recording(base, iq, calib, sx)
    frame, n, cal, sx2 = f_parse_data2(base)
    assert n == 5 and np.array_equal(np.asarray(frame), iq)
    assert np.allclose(cal, calib / 4095.0)
    assert fmcw_configurations(sx2) == fmcw_configurations(sx)


def test_jsonencode_conventions(tmp_path):
    path = str(tmp_path / "x.json")
    P.write_struct(path, [("row", np.arange(3.0)[None, :]), ("col", np.arange(3.0)[:, None]),
                          ("mat", np.array([[1.0, np.nan], [np.inf, 0.1 + 0.2]])), ("one", np.array([[5.0]])),
                          ("s", "radar_data"), ("f32", np.float32([0.1, 0.25]))])
    txt = open(path).read()
    d = json.loads(txt)
    assert list(d) == ["row", "col", "mat", "one", "s", "f32"]                   # field order
    assert d["row"] == [0, 1, 2] and d["col"] == [0, 1, 2]                       # vectors flatten
    assert d["mat"] == [[1, None], [None, 0.1 + 0.2]]                            # row-major nesting, NaN/Inf -> null, shortest round trip
    assert '0.30000000000000004' in txt                                          # what MATLAB's jsonencode prints for 0.1 + 0.2
    assert d["one"] == 5 and d["s"] == "radar_data"
    assert '"f32":[0.1,0.25]' in txt                                             # float32 values in THEIR shortest round-trip form


def test_range_speed_growing_matrix_quirk():
    """RP:157-159 allocate 1 x N, RP:245-250 write (fr_idx, 1): lastDetectedFrame x N, data in column 1."""
    vals = np.array([3.0, 0.0, 4.5, 0.0, 0.0])
    det = np.array([1, 0, 1, 0, 0], dtype=bool)
    M = P.matlab_growing_matrix(vals, det)
    assert M.shape == (3, 5) and M[0, 0] == 3.0 and M[2, 0] == 4.5 and M.sum() == 7.5
    assert P.matlab_growing_matrix(vals, np.zeros(5, bool)).shape == (1, 5)
    fields = dict(P.range_speed_payload(5, vals, vals, det, "radar_data"))
    assert np.allclose(fields["time_axis"], np.arange(5) * 0.15) and fields["filename"] == "radar_data"


def test_payload_keys_match_reference():
    keys = lambda f: [k for k, _ in f]
    assert keys(P.spectrogram_payload([0], [1], np.zeros((2, 2)))) == ["time", "frequency", "intensity", "title", "xLabel", "yLabel"]
    assert dict(P.spectrogram_payload([0], [1], np.zeros((2, 2))))["title"] == "All Frames - Log-Scaled Spectrogram"
    assert keys(P.range_fft_payload(2, [0], np.zeros((2, 2)), "f")) == ["time_axis", "array_bin_range", "range_tx1rx1_max_abs", "filename"]
    assert keys(P.fft_payload(np.zeros(4), "f")) == ["range_bins", "magnitude", "frame_index", "filename"]
    b = dict(P.batch_spectrogram_payload([0], [1], np.zeros((2, 2)), 3, 201, 300, "radar_data"))
    assert b["title"] == "Spectrogram - Batch 3" and b["start_frame"] == 201 and b["end_frame"] == 300
    assert b["xLabel"] == "Time (s) (relative to detected activity)" and b["filename_base"] == "radar_data"


def test_main_reports_missing_files_like_the_reference(tmp_path):
    from fmcw_radar_processing_b200.radar_processing import main
    r = main({"processAnimalActivity": "no", "workdir": str(tmp_path)})
    assert r["status"] == "error" and r["message"] == "Failed at reading files from blob storage."
    assert r["steps"][0]["step"] == "Read Files" and r["steps"][0]["status"] == "error"


def test_main_reports_processing_failure_without_gpu(tmp_path):
    """Without a CUDA device the library refuses to run (no CPU fallback) and main() reports a failed step
    (RPA:56-66) instead of raising."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fmcw_radar_processing_b200.radar_processing import main
    sx = make_sxml(numSamplesPerChirp=64, numChirpsPerFrame=16)
    write_recording(str(tmp_path / "radar_data"), np.full((2, 1, 16, 64, 2), 2048, np.int16), synth.default_calib(1, 64), sx)
    r = main({"processAnimalActivity": "no", "workdir": str(tmp_path)})
    assert r["status"] == "error" and r["message"] == "Failed at radar processing step."
    assert [s["step"] for s in r["steps"]] == ["Read Files", "Radar Processing"]


def test_native_json_writer_matches_python_writer(tmp_path):
    """fmcw_json_append_f32/f64 (host-only, multi-threaded) against the Python writer: same nesting, nulls, values."""
    rng = np.random.default_rng(3)
    a = (rng.standard_normal((300, 257)) * 40).astype(np.float32)
    a[2, 3] = np.nan
    a[5, 7] = np.inf
    view = np.ascontiguousarray(a.T).T                      # strided like intensity[:ncol].T
    d = rng.standard_normal((260, 256))
    fields = [("time", np.arange(5) * 0.15), ("intensity", view), ("m64", d), ("title", "x")]
    p_native, p_py = str(tmp_path / "n.json"), str(tmp_path / "p.json")
    try:
        P.write_struct(p_native, fields)
        P.USE_NATIVE = False
        P.write_struct(p_py, fields)
    finally:
        P.USE_NATIVE = True
    n, p = json.load(open(p_native)), json.load(open(p_py))
    assert list(n) == list(p) == ["time", "intensity", "m64", "title"]
    assert n["intensity"][2][3] is None and n["intensity"][5][7] is None
    gn = np.array([[np.nan if v is None else v for v in r] for r in n["intensity"]], dtype=np.float32)
    m = np.isfinite(a)
    assert np.array_equal(gn[m], a[m])                      # shortest round-trip digits reproduce every float32
    assert np.allclose(np.array(n["m64"]), d, rtol=0, atol=0)
    assert np.array_equal(np.array(p["m64"]), d)            # the Python stand-in round-trips too
    gp = np.array([[np.nan if v is None else v for v in r] for r in p["intensity"]], dtype=np.float32)
    assert np.array_equal(gp[m], a[m])


def test_spectrogram_png_writer(tmp_path):
    """payloads.write_spectrogram_png (RP:332-348): jet, clim [-40 0], frequency upwards, decimated to max_width columns."""
    import struct
    import zlib
    band = np.full((500, 40), -60.0, dtype=np.float32)         # [ncol][n_rows]
    band[:, 0] = 0.0                                            # lowest frequency row at the maximum
    band[:, 20] = -20.0
    band[3, 5] = -np.inf
    w, h = P.write_spectrogram_png(str(tmp_path / "s.png"), band, max_width=250)
    assert (w, h) == (250, 40)
    png = open(tmp_path / "s.png", "rb").read()
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    ilen = struct.unpack(">I", png[33:37])[0]
    px = np.frombuffer(zlib.decompress(png[41:41 + ilen]), dtype=np.uint8).reshape(h, 1 + 3 * w)[:, 1:].reshape(h, w, 3)
    assert tuple(px[-1, 0]) == (128, 0, 0)                      # 0 dB -> top of jet (dark red), bottom image row = lowest frequency
    assert tuple(px[0, 0]) == (0, 0, 128)                       # below clim -> bottom of jet (dark blue)
    assert px[h - 1 - 20, 0, 1] > 200                           # -20 dB = middle of the map: green channel saturated


def test_device_info_records_decode_like_the_c_struct():
    """fleet.py reads one fmcw_device_info per recording from a single device copy: the ctypes mirror has the C layout
    (9 x u64, 2 x u32, f64, 2 x i32 = 96 bytes) and parse_device_infos walks consecutive records."""
    import ctypes as C
    from fmcw_radar_processing_b200 import _lib
    from fmcw_radar_processing_b200.api import FmcwCuda
    assert C.sizeof(_lib.fmcw_run_info) == 88 and C.sizeof(_lib.fmcw_device_info) == 96
    recs = (_lib.fmcw_device_info * 3)()
    for i, r in enumerate(recs):
        r.info.n_frames = 500 + i; r.info.n_detected = 400 + i; r.info.L_local = (400 + i) * 64; r.info.nfft = 32768
        r.info.ncol_local = (400 + i) * 64 - 19; r.info.pmax_raw = 1.5e9 * (i + 1); r.status = -4 * (i == 2)
    raw = np.frombuffer(bytes(recs), dtype=np.uint8)
    out = FmcwCuda.parse_device_infos(raw)
    assert len(out) == 3
    for i, (info, status) in enumerate(out):
        assert info["n_frames"] == 500 + i and info["n_detected"] == 400 + i and info["nfft"] == 32768
        assert info["ncol_local"] == (400 + i) * 64 - 19 and info["pmax_raw"] == 1.5e9 * (i + 1)
        assert status == (-4 if i == 2 else 0)
