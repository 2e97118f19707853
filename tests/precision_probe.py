"""Precision probe (not a test): prints parity metrics of the CUDA path against the float64 oracle, including the
spectrogram error by level band quoted in DESIGN.md.  python tests/precision_probe.py > profiles/precision_<round>.txt"""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from tests import helpers as H
from oracle import fmcw_oracle as O
from fmcw_radar_processing_b200.api import FmcwCuda
from fmcw_radar_processing_b200 import synth

for (NTS, PN, nf) in [(128, 64, 60), (64, 16, 80), (256, 256, 6)]:
    case = H.make_case(n_frames=nf, NTS=NTS, PN=PN)
    ref = H.oracle_no(case)
    h = FmcwCuda(case["cfg"], case["calib"])
    out, inten = h.run(case["iq"])
    info = h.info()
    print("shape", NTS, PN, nf, info)
    print(" det equal", np.array_equal(out["detected"].astype(bool), ref["detected"]),
          "range_bin equal", np.array_equal(out["range_bin"][ref["detected"]], ref["range_idx"][ref["detected"]] - 1),
          "doppler_bin equal", np.array_equal(out["doppler_bin"][ref["detected"]], ref["doppler_idx"][ref["detected"]] - 1))
    print(" range_max_abs err", H.db_errors(out["range_max_abs"], ref["range_tx1rx1_max_abs"].T))
    d = ref["detected"]
    print(" range_mag rel", np.abs(out["range_mag"][d] / ref["range_mag"][d] - 1).max())
    gd = out["doppler_row"][..., 0] + 1j * out["doppler_row"][..., 1]
    print(" doppler_row err", H.db_errors(np.abs(gd[d]), np.abs(ref["doppler_rows"][d])))
    slow_ref = np.abs(ref["slow_time_signal_all_frames"]).reshape(-1, PN)
    print(" slow mag rel", np.abs(out["slow_time_mag"][d] / slow_ref - 1).max())
    st = ref["stft"]
    nc = info["ncol_local"]
    print(" ncol", nc, st["intensity"].shape, "nfft", info["nfft"], st["nfft"], "pmax rel", info["pmax_raw"] / st["pmax_raw"] - 1)
    print(" spectrogram e2e err", H.spectrogram_errors(inten[:nc].T, st["intensity"]))
    # STFT in isolation on the float32-rounded oracle signal
    x32 = np.abs(ref["slow_time_signal_all_frames"]).astype(np.float32)
    ref2 = O.stft_restated(x32.astype(np.float64), case["ocfg"])
    g2 = h.stft(x32)
    print(" spectrogram isolated err", H.spectrogram_errors(g2[:nc].T, ref2["intensity"]), "min dB", ref2["intensity"].min())
    g3 = h.stft(x32, layout=1)
    print(" layouts agree", np.array_equal(g3[:, :nc], g2[:nc].T))
    T, F, nfft, nct = h.stft_axes(info["L_total"])
    print(" axes", np.abs(T - st["T"]).max(), np.abs(F / st["frequency"] - 1).max())
    # generator parity
    g = h.synth_frames(case["tables"], case["scene"].seed, 0)
    print(" synth mismatches", int((g != case["iq"]).sum()))
    h.close()

# spectrogram error by level band (end to end), default (tensor-core TF32 x 2) and FMCW_OPT_STFT_PRECISION = 1 (float64)
from fmcw_radar_processing_b200 import _lib as L
BANDS = [(-60, 1), (-100, -60), (-140, -100), (-180, -140), (-220, -180), (-260, -220), (-400, -260)]
for scene_name, scene in (("C1 scene (single target)", None), ("C2 scene (torso + limbs)", synth.scene_c2(2))):
    case = H.make_case(n_frames=60, NTS=128, PN=64, scene=scene)
    ref = H.oracle_no(case)
    r = ref["stft"]["intensity"]
    for mode in (0, 1):
        h = FmcwCuda(case["cfg"], case["calib"])
        h.set_option(L.OPT_STFT_PRECISION, mode)
        out, inten = h.run(case["iq"])
        nc = h.info()["ncol_local"]
        g = inten[:nc].T.astype(np.float64)
        print(f"{scene_name}, STFT precision mode {mode}: non-finite agree {np.array_equal(np.isfinite(g), np.isfinite(r))}")
        fin = np.isfinite(r) & np.isfinite(g)
        for lo, hi in BANDS:
            m = fin & (r > lo) & (r <= hi)
            if m.any():
                rel = np.abs(10 ** ((g[m] - r[m]) / 20) - 1)
                print(f" band ({lo:4d},{hi:4d}] dB: n={m.sum():8d}  max |ddB|={np.abs(g[m]-r[m]).max():.2e}  max rel={rel.max():.2e}  p99.9 rel={np.quantile(rel,0.999):.2e}")
        h.close()

# cost of the float64 kernel on the C2 workload (5,000 frames)
import torch
from bench import DeviceRun, time_device
for mode in (0, 1):
    rr = DeviceRun("c2", 5000, 0, 1, 0)
    rr.h.set_option(L.OPT_STFT_PRECISION, mode)
    ms, st = time_device(rr, lambda: rr.h.run(rr.iq, rr.out, rr.inten), 10, 3, lambda: torch.cuda.synchronize())
    print(f"C2 5,000 frames, STFT precision mode {mode}: step {ms:.4f} ms, stft_main {st[3]:.4f} ms")
    rr.close()
