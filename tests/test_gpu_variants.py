"""The kernel variants the library can fall back to (environment switches read once per process) give the same results as the
default path: the STFT epilogue without TMA stores / with 2-D boxes, the CUDA-core STFT, the round-1 frame-chain kernel.
Each variant runs in its own process through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from tests import helpers as H
from fmcw_radar_processing_b200.api import FmcwCuda
case = H.make_case(n_frames=300, NTS=128, PN=64)
h = FmcwCuda(case["cfg"], case["calib"])
out, inten = h.run(case["iq"])
info = h.info()
np.savez(sys.argv[1], inten=inten[:info["ncol_total"]], range_max_abs=out["range_max_abs"], range_bin=out["range_bin"],
         doppler_bin=out["doppler_bin"], detected=out["detected"], pmax=info["pmax_raw"])
h.close()
"""


def _run(tmp_path, name, env):
    path = str(tmp_path / (name + ".npz"))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", _CHILD % {"root": ROOT}, path], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(path)


@pytest.fixture(scope="module")
def default_result(tmp_path_factory):
    return _run(tmp_path_factory.mktemp("variants"), "default", {})


@pytest.mark.parametrize("env", [{"FMCW_TC_TMA": "0"}, {"FMCW_TC_TMA3D": "0"}])
def test_stft_epilogue_variants_are_bit_identical(tmp_path, default_result, env):
    """Same arithmetic, different way out of shared memory: every float of the spectrogram is identical."""
    got = _run(tmp_path, "v", env)
    assert got["inten"].shape == default_result["inten"].shape and got["inten"].shape[0] > 128 * 148   # more tiles than SMs
    assert np.array_equal(got["inten"], default_result["inten"])


def test_cuda_core_stft_matches_the_tensor_core_kernel(tmp_path, default_result):
    """FMCW_STFT_VARIANT=3: float32 FMA evaluation of the same bracket-bin sums; within the spectrogram contract of each other."""
    got = _run(tmp_path, "v", {"FMCW_STFT_VARIANT": "3"})
    a, b = got["inten"].astype(np.float64), default_result["inten"].astype(np.float64)
    strong = b > -60.0
    assert np.abs(a[strong] - b[strong]).max() < 2e-3
    assert float(got["pmax"]) == pytest.approx(float(default_result["pmax"]), rel=1e-6)


def test_round1_frame_chain_kernel_matches_the_warp_kernel(tmp_path, default_result):
    """FMCW_CHAIN_VARIANT=0: same bins and detections; magnitudes to float32 FFT rounding."""
    got = _run(tmp_path, "v", {"FMCW_CHAIN_VARIANT": "0"})
    for k in ("range_bin", "doppler_bin", "detected"):
        assert np.array_equal(got[k], default_result[k]), k
    a, b = got["range_max_abs"].astype(np.float64), default_result["range_max_abs"].astype(np.float64)
    peak = b.max(axis=1, keepdims=True)
    strong = b > peak * 1e-3
    assert np.abs(20 * np.log10(a[strong] / b[strong])).max() < 1e-3
