"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them (CPU), and the CUDA path must match them (GPU)."""
import glob
import os

import numpy as np
import pytest

from fmcw_radar_processing_b200 import synth
from tests import helpers as H

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _case(g):
    n, NTS, PN, n_rx, seed = (int(v) for v in g["params"])
    case = H.make_case(n_frames=n, NTS=NTS, PN=PN, n_rx=n_rx, seed=seed)
    return case


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_and_generator_reproduce_golden(path):
    g = np.load(path)
    case = _case(g)
    assert np.array_equal(case["iq"], g["iq"])                       # the counter-based generator is stable
    ref = H.oracle_no(case)
    assert np.array_equal(ref["range_idx"], g["range_idx"]) and np.array_equal(ref["doppler_idx"], g["doppler_idx"])
    assert np.allclose(ref["range_tx1rx1_max_abs"], g["range_tx1rx1_max_abs"], rtol=1e-12, atol=1e-12)
    assert np.allclose(ref["stft"]["intensity"], g["intensity"], rtol=0, atol=1e-8)
    assert np.allclose(ref["stft"]["frequency"], g["frequency"], rtol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_path_matches_golden(path):
    from fmcw_radar_processing_b200.api import FmcwCuda
    g = np.load(path)
    case = _case(g)
    h = FmcwCuda(case["cfg"], g["calib_codes"] / 4095.0)
    out, inten = h.run(np.ascontiguousarray(g["iq"]))
    info = h.info()
    d = g["detected"]
    assert np.array_equal(out["detected"].astype(bool), d)
    assert np.array_equal(out["range_bin"][d], g["range_idx"][d] - 1)
    assert np.array_equal(out["doppler_bin"][d], g["doppler_idx"][d] - 1)
    e_db, _ = H.db_errors(out["range_max_abs"], g["range_tx1rx1_max_abs"].T)
    assert e_db < 1e-3
    nc = info["ncol_local"]
    assert nc == g["intensity"].shape[1] and info["nfft"] == int(g["nfft"])
    s_db, _ = H.spectrogram_errors(inten[:nc].T, g["intensity"])
    assert s_db < 1e-3
    h.close()
