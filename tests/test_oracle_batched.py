"""The vectorised oracle (CPU baseline of bench.py --impl reference) equals the serial oracle."""
import numpy as np

from oracle import fmcw_oracle_batched as OB
from tests import helpers as H


def test_batched_equals_serial():
    case = H.make_case(n_frames=24, NTS=128, PN=64, n_rx=2)
    ref = H.oracle_no(case)
    got = OB.run_no_branch(case["iq"], case["calib_codes"], case["sxml"], workers=2)
    assert np.array_equal(got["detected"], ref["detected"])
    assert np.array_equal(got["range_idx"], ref["range_idx"])
    assert np.array_equal(got["doppler_idx"], ref["doppler_idx"])
    assert np.allclose(got["range_tx1rx1_max_abs"], ref["range_tx1rx1_max_abs"], rtol=1e-12, atol=1e-12)
    assert np.allclose(got["slow_time_signal_all_frames"], ref["slow_time_signal_all_frames"], rtol=1e-12, atol=1e-9)
    assert np.allclose(got["doppler_rows"], ref["doppler_rows"], rtol=1e-10, atol=1e-9)
    assert np.allclose(got["stft"]["intensity"], ref["stft"]["intensity"], rtol=0, atol=1e-7)


def test_batched_handles_no_detection():
    case = H.make_case(n_frames=6, NTS=64, PN=16)
    case["iq"][2:4] = 2048          # DC only: nothing after calibration + mean removal
    ref = H.oracle_no(case)
    got = OB.run_no_branch(case["iq"], case["calib_codes"], case["sxml"])
    assert list(got["detected"]) == list(ref["detected"])
    assert not got["detected"][2] and not got["detected"][3]
    assert np.allclose(got["stft"]["intensity"], ref["stft"]["intensity"], atol=1e-7)
