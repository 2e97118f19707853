"""Pins the oracle (and on a GPU box the CUDA library) to OUTPUTS OF THE REFERENCE ITSELF, when they exist.

The reference is MATLAB; neither MATLAB nor GNU Octave exists in the build image or on the GPU boxes, so these files
cannot be produced here.  On any host that has one:

    python tests/golden/make_ref_cases.py                                  # (already committed) the input recordings
    octave --eval "addpath('matlab'); make_reference_golden('<reference>/radar-etl-pipeline')"
    git add tests/golden/ref_out && git commit

runs the untouched radar_processing_with_azure.m -> radar_processing.m on the committed recordings and leaves the JSON
files it wrote under tests/golden/ref_out/<case>/<no|yes>/.  Every test here SKIPS while those files are absent (the
oracle then stays "parity unpinned") and compares against them when they are present:
float64 oracle vs. the reference's jsonencode output (15 significant digits) at 1e-9 relative; the CUDA library at the
north_star tolerances.
"""
import json
import os

import numpy as np
import pytest

from fmcw_radar_processing_b200 import parse
from oracle import fmcw_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = os.path.join(HERE, "golden", "ref_cases")
OUT = os.path.join(HERE, "golden", "ref_out")


def _cases():
    return sorted(d for d in os.listdir(CASES) if os.path.isdir(os.path.join(CASES, d))) if os.path.isdir(CASES) else []


def _ref(case, branch, name):
    p = os.path.join(OUT, case, branch, name)
    if not os.path.exists(p):
        pytest.skip(f"{os.path.relpath(p, HERE)} not generated yet (needs MATLAB / Octave: matlab/make_reference_golden.m)")
    with open(p) as f:
        return json.load(f)


def _arr(v):
    return np.array([[np.nan if x is None else x for x in row] if isinstance(row, list) else (np.nan if row is None else row)
                     for row in v], dtype=np.float64)


def _oracle_no(case):
    frame, n, calib, sx = parse.f_parse_data2(os.path.join(CASES, case, "radar_data"))
    frames, _, cal, _ = O.f_parse_data2(np.asarray(frame), calib * 4095.0, sx)
    return O.radar_processing_no(frames, cal, sx, stft="literal"), n


def test_recordings_are_committed_and_readable():
    assert len(_cases()) >= 2
    for c in _cases():
        frame, n, calib, sx = parse.f_parse_data2(os.path.join(CASES, c, "radar_data"))
        assert frame.shape[0] == n and frame.shape[1] * frame.shape[2] * 0 == 0
        assert n * frame.shape[2] >= 100                        # RP:410: linear index 100 must exist
        assert "Text" in sx["Device"]["BaseEndpoint"]["chirpDuration_ns"]


@pytest.mark.parametrize("case", _cases())
def test_oracle_matches_reference_no_branch(case):
    spec = _ref(case, "no", "spectrogram_data.json")
    rfft = _ref(case, "no", "radar_data_range_fft_data.json")
    rs = _ref(case, "no", "radar_data_range_speed_data.json")
    ref, n = _oracle_no(case)
    st = ref["stft"]
    assert np.allclose(_arr(spec["time"]), st["T"], rtol=1e-12)
    assert np.allclose(_arr(spec["frequency"]), st["frequency"], rtol=1e-12)
    g = _arr(spec["intensity"])
    assert g.shape == st["intensity"].shape
    fin = np.isfinite(st["intensity"])
    assert np.array_equal(np.isfinite(g), fin)
    assert np.abs(g[fin] - st["intensity"][fin]).max() < 1e-8                  # dB; jsonencode keeps 15 digits
    assert np.allclose(_arr(rfft["range_tx1rx1_max_abs"]), ref["range_tx1rx1_max_abs"], rtol=1e-9, atol=1e-9)
    # range / speed: the growing-matrix quirk of RP:245-250 puts the track in column 1
    rng = _arr(rs["range"])
    rng = rng[:, 0] if rng.ndim == 2 else rng
    last = int(np.flatnonzero(ref["detected"]).max()) + 1
    assert np.allclose(rng[:last], ref["range"][:last], rtol=1e-12, atol=1e-12)     # identical range bins (peak search)
    spd = _arr(rs["speed"])
    spd = spd[:, 0] if spd.ndim == 2 else spd
    assert np.allclose(spd[:last], ref["speed"][:last], rtol=1e-12, atol=1e-12)     # identical Doppler bins


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases())
def test_library_matches_reference_no_branch(case):
    from fmcw_radar_processing_b200.api import FmcwCuda
    from fmcw_radar_processing_b200.config import fmcw_configurations
    from tests import helpers as H
    spec = _ref(case, "no", "spectrogram_data.json")
    rfft = _ref(case, "no", "radar_data_range_fft_data.json")
    frame, n, calib, sx = parse.f_parse_data2(os.path.join(CASES, case, "radar_data"))
    cfg = fmcw_configurations(sx)
    h = FmcwCuda(cfg, calib)
    out, inten = h.run(np.ascontiguousarray(frame))
    nc = h.info()["ncol_local"]
    H.assert_spectrogram_contract(inten[:nc].T, _arr(spec["intensity"]))
    e_db, e_rel = H.db_errors(out["range_max_abs"], _arr(rfft["range_tx1rx1_max_abs"]).T)
    assert e_db < 1e-3 and e_rel < 2e-4
    h.close()


@pytest.mark.parametrize("case", _cases())
def test_oracle_matches_reference_yes_branch(case):
    """RP:444-607: one spectrogram JSON per 100-frame batch (at most four), title / axis labels / frame range included."""
    first = _ref(case, "yes", "radar_data_spectrogram_batch_1.json")
    frame, n, calib, sx = parse.f_parse_data2(os.path.join(CASES, case, "radar_data"))
    frames, _, cal, _ = O.f_parse_data2(np.asarray(frame), calib * 4095.0, sx)
    ref = O.radar_processing_yes(frames, cal, sx, stft="literal")
    assert len(ref["batches"]) >= 1
    for b in ref["batches"]:
        js = first if b["batch"] == 1 else _ref(case, "yes", f"radar_data_spectrogram_batch_{b['batch']}.json")
        assert js["title"] == f"Spectrogram - Batch {b['batch']}" and js["start_frame"] == b["start_frame"] and js["end_frame"] == b["end_frame"]
        g = _arr(js["intensity"])
        fin = np.isfinite(b["intensity"])
        assert g.shape == b["intensity"].shape and np.array_equal(np.isfinite(g), fin)
        assert np.abs(g[fin] - b["intensity"][fin]).max() < 1e-8
        assert np.allclose(_arr(js["time"]), b["T"], rtol=1e-12) and np.allclose(_arr(js["frequency"]), b["frequency"], rtol=1e-12)
