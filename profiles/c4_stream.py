#!/usr/bin/env python
"""BASELINE configs[3] at its stated size: 4 RX x 256 chirps x 256 samples, 2,000,000 frames, frame-sharded over the GPUs of
one node.  2.1 TB of samples and 2.1 TB of hop-1 spectrogram never exist at once: frames are generated ON THE DEVICE chunk by
chunk by the counter-based generator (same bits as the NumPy generator, tests/test_gpu_pipeline.py), processed and dropped;
the spectrogram is written piece by piece into a reusable buffer (fmcw_radar_processing_b200/streaming.py).

    python profiles/c4_stream.py [--frames-total N]                                       # one GPU
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 profiles/c4_stream.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fmcw_radar_processing_b200 import synth  # noqa: E402
from fmcw_radar_processing_b200.config import fmcw_configurations  # noqa: E402
from fmcw_radar_processing_b200.parse import make_sxml  # noqa: E402
from fmcw_radar_processing_b200.streaming import StreamingRecording  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames-total", type=int, default=2000000)
ap.add_argument("--chunk", type=int, default=4096, help="frames generated + processed per step (1 MiB each)")
ap.add_argument("--piece-cols", type=int, default=4 * 1024 * 1024, help="spectrogram columns per piece (4 KiB each)")
a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NTS, PN, RX = 256, 256, 4
sx = make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=RX)
cfg = fmcw_configurations(sx)
scene = synth.scene_c1(seed=4)
per = -(-a.frames_total // world)
f_lo, f_hi = rank * per, min(a.frames_total, (rank + 1) * per)
n_local = max(0, f_hi - f_lo)
s = StreamingRecording(cfg, synth.default_calib(RX, NTS) / 4095.0, device=lr, distributed=world > 1)
s.reserve(n_local)
iq = torch.empty((a.chunk, RX, PN, NTS, 2), dtype=torch.int16, device=dev)
out = s.h.alloc_frame_out(a.chunk, device=dev)
buf = torch.empty((a.piece_cols + 32, 1024), dtype=torch.float32, device=dev)


def sync():
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)


sync()
t_gen = t_chain = 0.0
n_det = 0
t0 = time.perf_counter()
for f0 in range(f_lo, f_hi, a.chunk):
    m = min(a.chunk, f_hi - f0)
    tab = synth.scene_tables(scene, cfg["dist_per_bin"], 256, cfg["PRT"], cfg["lambda"], f0, m)
    ta = time.perf_counter()
    s.h.synth_frames(tab, scene.seed, f0, sigma=scene.sigma, dc=scene.dc, rx_step=scene.rx_step, out=iq[:m])   # synchronises
    tb = time.perf_counter()
    o = s.push_frames(iq[:m], {k: v[:m] for k, v in out.items()}, keep_track=False)                             # info(): synchronises
    tc = time.perf_counter()
    t_gen += tb - ta
    t_chain += tc - tb
sync()
t_pass1 = time.perf_counter() - t0
checks = {"sum": 0.0, "cols": 0}


def consumer(c0, n, b):
    # stand-in for the real consumer (host copy, file, classifier): one strided read so that every piece is observed
    checks["sum"] += float(b[:n:4096, ::64].double().sum().item())
    checks["cols"] += n


consumer(0, 8192, buf)           # warm the consumer's own kernels up outside the timed region
checks = {"sum": 0.0, "cols": 0}
sync()
t0 = time.perf_counter()
r = s.stft(buf, piece_cols=a.piece_cols, consumer=consumer)
sync()
t_stft_wall = time.perf_counter() - t0
t_stft = r["seconds"]["max_pass"] + r["seconds"]["spectrogram_pass"]      # without the stand-in consumer
vals = torch.tensor([t_gen, t_chain, t_pass1, t_stft, float(s.L // PN), float(checks["cols"]), t_stft_wall], dtype=torch.float64, device=dev)
if world > 1:
    mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
else:
    mx = sm = vals
if rank == 0:
    t_gen, t_chain, t_pass1, t_stft = (float(v) for v in mx[:4])
    n = a.frames_total
    print(json.dumps({
        "workload": "C4 (BASELINE configs[3]): 4 RX x 256 chirps x 256 samples, RX 1 processed (RP:202), hop-1 STFT (window 20)",
        "frames_total": n, "n_gpus": world, "frames_per_gpu": per, "chunk_frames": a.chunk, "piece_cols": a.piece_cols,
        "detected_frames": int(sm[4]), "spectrogram_columns": int(sm[5]), "L_total": r["L_total"], "nfft": r["nfft"], "pieces_per_gpu": r["pieces"],
        "seconds_max_over_ranks": {"generate_on_device": t_gen, "frame_chain_incl_slow_time_append": t_chain, "pass1_wall": t_pass1,
                                   "stft_max_pass_plus_spectrogram_pass": t_stft},
        "frames_per_s": {"chain_only": n / t_chain, "chain_plus_stft": n / (t_chain + t_stft), "incl_generation": n / (t_pass1 + t_stft)},
        "bytes": {"samples_generated": n * RX * PN * NTS * 4, "samples_read_by_the_chain": n * PN * NTS * 4,
                  "spectrogram_written": int(sm[5]) * 4096},
        "gbs_per_gpu": {"chain_input": n * PN * NTS * 4 / world / t_chain / 1e9, "stft_output": int(sm[5]) * 4096 / world / t_stft / 1e9},
        "stft_seconds_rank0": r["seconds"], "stft_wall_incl_consumer_and_exchanges": float(mx[6]),
        "timing": "host wall clock, device synchronised at every chunk / piece boundary (the streaming driver is host-synchronous)",
    }, indent=1))
if world > 1:
    dist.destroy_process_group()
