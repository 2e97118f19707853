#!/usr/bin/env python
"""Writes the input recordings of the reference-run harness: tests/golden/ref_cases/<case>/radar_data.{xml,raw.bin}.

    python tests/golden/make_ref_cases.py

matlab/make_reference_golden.m then runs the UNTOUCHED radar_processing_with_azure.m / radar_processing.m on them (MATLAB
or GNU Octave, on any host that has one) and leaves the JSON files the reference wrote under tests/golden/ref_out/;
tests/test_reference_golden.py compares the oracle and the CUDA library with those files.  The recordings are small
(the literal STFT of the reference is O(L^2)) and seeded; the generating code is fmcw_radar_processing_b200/synth.py.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fmcw_radar_processing_b200 import parse  # noqa: E402
from tests import helpers as H  # noqa: E402

# name -> make_case arguments.  PN * n_frames >= 100: the reference's fft_data block reads linear index 100 (RP:410-411).
CASES = {"c1_128x64_12f": dict(n_frames=12, NTS=128, PN=64, n_rx=1, seed=1),
         "field_64x16_2rx_40f": dict(n_frames=40, NTS=64, PN=16, n_rx=2, seed=11),
         "field_64x16_210f_yes": dict(n_frames=210, NTS=64, PN=16, n_rx=1, seed=5)}    # three batches of the 'yes' branch


def main():
    base = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cases")
    for name, kw in CASES.items():
        d = os.path.join(base, name)
        os.makedirs(d, exist_ok=True)
        case = H.make_case(**kw)
        parse.write_recording(os.path.join(d, "radar_data"), case["iq"], case["calib_codes"], case["sxml"])
        print(name, sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d)), "bytes")


if __name__ == "__main__":
    main()
