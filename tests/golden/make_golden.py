#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ from the float64 oracle (the reference ships no
vectors of its own and cannot be run here; see oracle/fmcw_oracle.py "PARITY UNPINNED").

    python tests/golden/make_golden.py

Each .npz holds the seeded int16 input, the calibration codes, and the oracle's outputs for the 'no' branch.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import helpers as H  # noqa: E402

CASES = {"field_64x16": dict(n_frames=8, NTS=64, PN=16, n_rx=2, seed=11),
         "c1_128x64": dict(n_frames=3, NTS=128, PN=64, n_rx=1, seed=1)}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name, kw in CASES.items():
        case = H.make_case(**kw)
        ref = H.oracle_no(case)
        st = ref["stft"]
        np.savez_compressed(os.path.join(here, name + ".npz"), iq=case["iq"], calib_codes=case["calib_codes"],
                            params=np.array([kw["n_frames"], kw["NTS"], kw["PN"], kw["n_rx"], kw["seed"]]),
                            range_tx1rx1_max_abs=ref["range_tx1rx1_max_abs"], detected=ref["detected"],
                            range_idx=ref["range_idx"], range_mag=ref["range_mag"], doppler_idx=ref["doppler_idx"],
                            doppler_rows=ref["doppler_rows"], slow_time=ref["slow_time_signal_all_frames"],
                            T=st["T"], frequency=st["frequency"], intensity=st["intensity"].astype(np.float64),
                            nfft=st["nfft"], pmax_raw=st["pmax_raw"])
        print(name, os.path.getsize(os.path.join(here, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()
