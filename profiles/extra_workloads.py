#!/usr/bin/env python
"""Device-resident throughput of the fused chain on the other BASELINE.json configs (supplementary to bench.py,
which measures configs[1]).  python profiles/extra_workloads.py > profiles/extra_workloads_<round>.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fmcw_radar_processing_b200 import synth
from fmcw_radar_processing_b200.api import FmcwCuda
from fmcw_radar_processing_b200.config import fmcw_configurations
from fmcw_radar_processing_b200.parse import make_sxml

CASES = [("C1 shape, 500 frames, 1 RX 64x128", 500, 1, 64, 128, 20, 19, synth.scene_c1(1)),
         ("C2, 5,000 frames, 3 RX 64x128", 5000, 3, 64, 128, 20, 19, synth.scene_c2(2)),
         ("C3, 200,000 frames, 1 RX 64x128", 200000, 1, 64, 128, 20, 19, synth.scene_c1(3)),
         ("C4 shape, 20,000 frames, 4 RX 256x256", 20000, 4, 256, 256, 20, 19, synth.scene_c1(4)),
         ("field shape, 20,000 frames, 2 RX 16x64", 20000, 2, 16, 64, 20, 19, synth.scene_c1(5)),
         ("C5 point, 5,000 frames, window 64 / 75 % overlap", 5000, 1, 64, 128, 64, 48, synth.scene_c1(1000)),
         ("C5 point, 5,000 frames, window 256 / 90 % overlap", 5000, 1, 64, 128, 256, 230, synth.scene_c1(1001))]

dev = torch.device("cuda", 0)
print(f"{'workload':55s} {'frames/s':>12s} {'ms/step':>9s} {'chain':>8s} {'stft':>8s} {'GB/s (alg.)':>12s}")
for name, n, n_rx, PN, NTS, win, ov, scene in CASES:
    sx = make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=n_rx)
    cfg = fmcw_configurations(sx, window_length=win, overlap=ov)
    h = FmcwCuda(cfg, synth.default_calib(n_rx, NTS) / 4095.0, torch_stream_sync=False)
    tab = synth.scene_tables(scene, cfg["dist_per_bin"], 256, cfg["PRT"], cfg["lambda"], 0, n)
    iq = torch.empty((n, n_rx, PN, NTS, 2), dtype=torch.int16, device=dev)
    h.synth_frames(tab, scene.seed, 0, out=iq)
    out = h.alloc_frame_out(n, device=dev)
    inten = torch.empty((max(1, h.max_cols(n)), 1024), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    for _ in range(3):
        h.run(iq, out, inten)
    h.synchronize()
    st = torch.cuda.ExternalStream(h.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10 if n <= 20000 else 3
    e0.record(st)
    for _ in range(K):
        h.run(iq, out, inten)
    e1.record(st)
    h.synchronize()
    ms = e0.elapsed_time(e1) / K
    tm = h.timings()
    info = h.info()
    bytes_alg = n * (NTS * PN * 4 + 256 * 4 + 16 + 64 + PN * 4) + info["L_total"] * 8 + info["ncol_local"] * 1024 * 4
    print(f"{name:55s} {n / ms * 1e3:12.0f} {ms:9.3f} {tm["chain_ms"]:8.3f} {tm["stft_main_ms"]:8.3f} pm {tm["plan_max_ms"]:6.3f} {bytes_alg / ms / 1e6:12.0f}"
          f"   det {info['n_detected']} cols {info['ncol_local']} nfft 2^{int(np.log2(info['nfft']))} bins {info['n_dtft_bins']}")
    h.close()
    del iq, out, inten
    torch.cuda.empty_cache()

# ---- C5 fleet: many independent radars (C1 shape, 500 frames each) on one GPU ----
import time
from fmcw_radar_processing_b200.fleet import Fleet

n_radars, n, n_rx, PN, NTS = 64, 500, 1, 64, 128
sx = make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=n_rx)
cfg = fmcw_configurations(sx)
calib = synth.default_calib(n_rx, NTS) / 4095.0
gen = FmcwCuda(cfg, calib, torch_stream_sync=False)
recs = []
for r in range(n_radars):
    scene = synth.scene_c1(1000 + r)
    tab = synth.scene_tables(scene, cfg["dist_per_bin"], 256, cfg["PRT"], cfg["lambda"], 0, n)
    iq = torch.empty((n, n_rx, PN, NTS, 2), dtype=torch.int16, device=dev)
    gen.synth_frames(tab, scene.seed, 0, out=iq)
    recs.append(iq)
gen.synchronize()
gen.close()
for nh in (1, 4, 8):
    fleet = Fleet(cfg, calib, n_handles=nh)
    fleet.run(recs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 3
    for _ in range(K):
        res = fleet.run(recs)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    print(f"C5 fleet, {n_radars} radars x {n} frames, {nh} handle(s) in flight: {n_radars * n / dt:12.0f} frames/s  "
          f"{dt * 1e3:8.3f} ms per fleet pass (host wall clock)")
    fleet.close()
