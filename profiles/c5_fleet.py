#!/usr/bin/env python
"""BASELINE configs[4] at its stated size: a fleet of 1,024 concurrent field radars (C1 shape: 1 RX x 64 chirps x 128
samples, 500 frames each, seeds 1000..2023; every radar has its own slow-time history, nfft and normalisation maximum),
STFT window / overlap sweep {32, 64, 128, 256} x {50, 75, 90 %} plus the reference's window 20 / overlap 19, whole radars per
GPU (no exchange between radars or GPUs).

    python profiles/c5_fleet.py [--radars 1024]                                            # one GPU takes all radars
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 profiles/c5_fleet.py

Per GPU the recordings go through fleet.Fleet: a pool of handles (own stream, tables, scratch each), recordings dealt round
robin, every hand-off on the device, so the small kernels of different radars overlap.  (A kernel that takes several radars
in one launch does not exist yet: DESIGN.md section 9.)
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fmcw_radar_processing_b200 import synth  # noqa: E402
from fmcw_radar_processing_b200.api import FmcwCuda  # noqa: E402
from fmcw_radar_processing_b200.config import fmcw_configurations  # noqa: E402
from fmcw_radar_processing_b200.fleet import Fleet  # noqa: E402
from fmcw_radar_processing_b200.parse import make_sxml  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--radars", type=int, default=1024)
ap.add_argument("--frames", type=int, default=500)
ap.add_argument("--handles", type=int, default=8)
ap.add_argument("--passes", type=int, default=3)
ap.add_argument("--only", type=int, default=0, help="first N entries of the sweep only")
a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NTS, PN = 128, 64
sx = make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=1)
calib = synth.default_calib(1, NTS) / 4095.0
mine = [r for r in range(a.radars) if r % world == rank]
# recordings, generated on the device: radar r = C1 scene with seed 1000 + r and its own start range
cfg0 = fmcw_configurations(sx)
g = FmcwCuda(cfg0, calib, device=lr)
recs = []
for r in mine:
    sc = synth.scene_c1(seed=1000 + r)
    sc.scatterers[0].R0 = 4.0 + (r % 17)
    tab = synth.scene_tables(sc, cfg0["dist_per_bin"], 256, cfg0["PRT"], cfg0["lambda"], 0, a.frames)
    iq = torch.empty((a.frames, 1, PN, NTS, 2), dtype=torch.int16, device=dev)
    g.synth_frames(tab, sc.seed, 0, sigma=sc.sigma, dc=sc.dc, rx_step=sc.rx_step, out=iq)
    recs.append(iq)
g.close()
sweep = [(20, 19)] + [(w, int(w * ov)) for w in (32, 64, 128, 256) for ov in (0.5, 0.75, 0.9)]
if a.only:
    sweep = sweep[:a.only]
rows = []
for win, ov in sweep:
    cfg = fmcw_configurations(sx, window_length=win, overlap=ov)
    fleet = Fleet(cfg, calib, n_handles=a.handles, device=lr)
    for _ in range(3):                                      # warm-up: buffers, plans; the third pass records the run graphs
        fleet.run(recs)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.passes):
        res = fleet.run(recs)
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / a.passes
    cols = sum(x["ncol"] for x in res)
    det = sum(x["info"]["n_detected"] for x in res)
    v = torch.tensor([dt, float(cols), float(det)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = v.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = v.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = v
    rows.append({"window": win, "overlap": ov, "hop": win - ov, "seconds_per_pass_max_over_ranks": float(mx[0]),
                 "frames_per_s": a.radars * a.frames / float(mx[0]), "spectrogram_columns": int(sm[1]), "detected_frames": int(sm[2]),
                 "spectrogram_gbs_all_gpus": float(sm[1]) * 4096 / float(mx[0]) / 1e9})
    fleet.close()
    del res
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps({"workload": "C5 (BASELINE configs[4]): fleet of independent radars, C1 shape, own history / nfft / maximum each",
                      "radars": a.radars, "frames_per_radar": a.frames, "n_gpus": world, "radars_per_gpu": len(mine),
                      "handles_per_gpu": a.handles, "passes_timed": a.passes,
                      "timing": "host wall clock around whole fleet passes, device synchronised on both sides; inputs resident in HBM",
                      "sweep": rows}, indent=1))
if world > 1:
    dist.destroy_process_group()
