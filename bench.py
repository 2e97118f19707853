#!/usr/bin/env python
"""Benchmark of the fused range-Doppler-STFT chain (BASELINE.json metric: radar frames/s and achieved
HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2] [--scaling weak|strong] [--impl reference]

A step = one pass of the whole hot path (frame chain -> compaction -> STFT plan + max -> STFT) over one
recording.  The headline workload is C3 (BASELINE.json configs[2]: 200,000 frames of the reference shape, the
largest configuration that fits one GPU: 6.6 GB of samples in, 52 GB of spectrogram out per step); ``--workload
c2`` runs configs[1] (5,000 frames, 3 RX, walking-animal scene), which is also measured after the headline and
reported under the key ``c2``.  ``value`` is measured with the inputs resident in HBM; ``e2e`` through the same
C-ABI call with pinned HOST buffers (H2D of the frames and D2H of every output inside the timed region).  With
N > 1 every rank owns a contiguous shard of ONE recording (weak: ``frames`` per GPU; strong: ``frames`` in
total); the STFT runs over the whole concatenated signal with a halo (fmcw_radar_processing_b200/distributed.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "radar_frames_per_second_range_doppler_stft"
UNIT = "frames/s"
PN, NTS = 64, 128
# algorithmic bytes per frame of the C1/C2/C3 shape (1 processed RX, 64 x 128, hop 1)
CHAIN_BYTES_PER_FRAME = NTS * PN * 4 + 256 * 4 + 16 + 16 * 4 + PN * 4          # 34,128: samples in, range_fft column, track, Doppler row, slow-time row out
STFT_BYTES_PER_FRAME = PN * 4 + PN * 1024 * 4                                  # 262,400: slow-time row in, 64 spectrogram columns out
WORKLOADS = {
    "c3": dict(frames=200000, n_rx=1, scene="c1", seed=3,
               text="C3 (BASELINE.json configs[2]): long streaming batch, single moving point target, 1 RX x 64 chirps x 128 samples"),
    "c2": dict(frames=5000, n_rx=3, scene="c2", seed=2,
               text="C2 (BASELINE.json configs[1]): multi-target walking-animal scene, 3 RX x 64 chirps x 128 samples, RX 1 processed as in the reference (RP:202)"),
}


def build_workload(name):
    from fmcw_radar_processing_b200 import synth
    from fmcw_radar_processing_b200.config import fmcw_configurations
    from fmcw_radar_processing_b200.parse import make_sxml
    w = WORKLOADS[name]
    sx = make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=w["n_rx"])
    cfg = fmcw_configurations(sx)
    scene = synth.scene_c1(seed=w["seed"]) if w["scene"] == "c1" else synth.scene_c2(seed=w["seed"])
    return sx, cfg, scene


def scene_tables(scene, cfg, frame0, n):
    from fmcw_radar_processing_b200 import synth
    return synth.scene_tables(scene, cfg["dist_per_bin"], cfg["range_fft_size"], cfg["PRT"], cfg["lambda"], frame0, n)


def kernels_per_step(frames, world, mailbox):
    """Launches of OUR kernels per step: look-ahead stft_plan + stft_tc_prepare (side stream), frame_chain_warp,
    compaction (one fused kernel up to 16,384 frames, else scan + gather), stft_plan (confirm), stft_tc_prepare,
    colstat, refine, hard, stft_tc; the sharded paths add their exchange kernels."""
    k = 9 + (1 if frames <= 16384 else 2)
    if world > 1:
        k = k + 3 if mailbox else k - 2 + 1          # mailbox: post_heads, post_max, collect_max; NCCL: no look-ahead, + shard_pack
    return k


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.active = False

    def run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            names = {getattr(N, k): k for k in dir(N) if k.startswith("nvmlClocksThrottleReason") and isinstance(getattr(N, k), int)}
            while not self.stop_flag:
                if self.active:
                    self.samples.append(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                    r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if bit and (r & bit) and "None" not in name and "All" not in name:
                            self.reasons.add(name.replace("nvmlClocksThrottleReason", ""))
                time.sleep(0.002)
        except Exception as e:   # NVML missing: report that, never fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


def cpu_baseline(sx, cfg, scene, n_sample, L_total_full, threads):
    """The oracle (port of the reference lines) timed on the host cores on a bounded sample of the same
    workload: the first n_sample frames through the serial frame loop (RP:197-261) and the restated STFT
    of their slow-time signal on the fine grid of the FULL recording (same bins per column)."""
    from threadpoolctl import threadpool_limits
    from fmcw_radar_processing_b200 import synth
    from oracle import fmcw_oracle as O
    tab = scene_tables(scene, cfg, 0, n_sample)
    iq = synth.synth_frames(tab, scene.seed, 0, cfg["num_Rx_antennas"], PN, NTS, sigma=scene.sigma, dc=scene.dc,
                            rx_step=scene.rx_step)
    calib = synth.default_calib(cfg["num_Rx_antennas"], NTS)
    frames, n, cal, _ = O.f_parse_data2(iq, calib, sx)
    ocfg = O.configure(sx)
    with threadpool_limits(limits=threads):
        t0 = time.perf_counter()
        r = O.radar_processing_no(frames, cal, sx, stft=None)
        x = np.abs(r["slow_time_signal_all_frames"])
        if len(x) >= ocfg.window_length:
            ncol_s = len(x) - ocfg.overlap
            O.stft_restated(x, ocfg, pmax_raw=None, col_range=(0, ncol_s), L_total=max(L_total_full, len(x)))
        dt = time.perf_counter() - t0
    return n_sample / dt, dt


def workload_config(args, world):
    w = WORKLOADS[args.workload]
    total = args.frames if args.scaling == "strong" else args.frames * world
    per_gpu_in = args.frames_per_gpu * PN * NTS * 4
    per_gpu_out = args.frames_per_gpu * PN * 1024 * 4
    return {"workload": f"{w['text']}; {total} frames in one recording ({args.frames_per_gpu} per GPU), range-Doppler + hop-1 "
                        f"STFT (window 20, 1024 log-frequency bins)",
            "name": args.workload, "frames_total": total, "frames_per_gpu": args.frames_per_gpu, "rx": w["n_rx"], "chirps": PN,
            "samples": NTS, "stft_window": 20, "stft_hop": 1,
            "intensity_layout": "time-major [col][1024] (MATLAB memory order)",
            "l2": f"inputs ({per_gpu_in / 1e6:.0f} MB of RX-1 samples) and outputs ({per_gpu_out / 1e9:.2f} GB) per step and GPU exceed "
                  f"the 126 MB L2; no explicit flush",
            "parallelism": f"frame-sharded x{world} ({args.scaling} scaling)" if world > 1 else "single GPU"}


def run_reference(args):
    """--impl reference: the reference's own (CPU) implementation of the path.  MATLAB/Octave are absent, so
    this is the oracle port, vectorised and with all host threads (oracle/fmcw_oracle_batched.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fmcw_radar_processing_b200 import synth
    from oracle import fmcw_oracle_batched as OB
    sx, cfg, scene = build_workload(args.workload)
    cores = os.cpu_count() or 1
    n_sample = args.ref_sample
    total = args.frames if args.scaling == "strong" else args.frames * args.gpus
    L_full = total * PN
    tab = scene_tables(scene, cfg, 0, n_sample)
    iq = synth.synth_frames(tab, scene.seed, 0, cfg["num_Rx_antennas"], PN, NTS, sigma=scene.sigma, dc=scene.dc,
                            rx_step=scene.rx_step)
    calib = synth.default_calib(cfg["num_Rx_antennas"], NTS)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        OB.run_no_branch(iq, calib, sx, L_total=L_full, workers=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    v = n_sample / (ms / 1e3)
    sample = (f"first {n_sample} frames of the workload per step: vectorised frame chain + restated STFT of their columns on the "
              f"full recording's fine grid (nfft = 2^nextpow2({L_full}))")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def host_mem_available():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable"):
                    return int(ln.split()[1]) * 1024
    except OSError:
        pass
    return None


class DeviceRun:
    """One workload resident in HBM on this rank: handle, synthetic frames, output buffers."""

    def __init__(self, name, n, rank, world, local_rank, frame0=None, miss_every=0):
        import torch
        from fmcw_radar_processing_b200 import synth
        from fmcw_radar_processing_b200.api import FmcwCuda
        self.torch = torch
        self.sx, self.cfg, self.scene = build_workload(name)
        self.n, self.n_rx = n, WORKLOADS[name]["n_rx"]
        self.dev = torch.device("cuda", local_rank)
        self.calib = synth.default_calib(self.n_rx, NTS) / 4095.0
        self.h = FmcwCuda(self.cfg, self.calib, device=local_rank, torch_stream_sync=False)
        frame0 = rank * n if frame0 is None else frame0
        self.iq = torch.empty((n, self.n_rx, PN, NTS, 2), dtype=torch.int16, device=self.dev)
        # synthetic input, generated on the device by the counter-based generator, in chunks of 20,000 frames
        for f0 in range(0, n, 20000):
            m = min(20000, n - f0)
            tab = scene_tables(self.scene, self.cfg, frame0 + f0, m)
            self.h.synth_frames(tab, self.scene.seed, frame0 + f0, sigma=self.scene.sigma, dc=self.scene.dc,
                                rx_step=self.scene.rx_step, out=self.iq[f0:f0 + m])
        if miss_every:      # frames without a target: DC only (no detection, RP:242), every miss_every-th frame
            self.iq[::miss_every] = 2048
        self.out = self.h.alloc_frame_out(n, device=self.dev)
        self.cols_cap = self.h.max_cols(n) + (20 if world > 1 else 0)
        self.inten = torch.empty((self.cols_cap, 1024), dtype=torch.float32, device=self.dev)
        self.stream = torch.cuda.ExternalStream(self.h.stream, device=self.dev)

    def close(self):
        self.h.close()
        del self.iq, self.inten, self.out
        self.torch.cuda.empty_cache()


def time_device(run, step, steps, warmup, barrier, sampler=None, stage_steps=None):
    """W untimed + K timed steps with CUDA events on the library's stream; then per-stage times from the events the
    library records around each stage."""
    torch = run.torch
    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.active = True
    run.stream.wait_stream(torch.cuda.current_stream(run.dev))
    e0.record(run.stream)
    for _ in range(steps):
        step()
    run.stream.wait_stream(torch.cuda.current_stream(run.dev))   # NCCL work of the sharded path runs on torch's stream
    e1.record(run.stream)
    barrier()
    if sampler:
        sampler.active = False
    dev_ms = e0.elapsed_time(e1) / steps
    stage_ms = np.zeros(4)
    k = stage_steps or min(steps, 5)
    for _ in range(k):
        step()
        tm = run.h.timings()
        stage_ms += np.array([tm["chain_ms"], tm["compact_ms"], tm["plan_max_ms"], tm["stft_main_ms"]])
    return dev_ms, stage_ms / k


def stage_dict(stage_ms):
    return {"frame_chain": float(stage_ms[0]), "compaction": float(stage_ms[1]), "stft_plan_and_max": float(stage_ms[2]),
            "stft_main": float(stage_ms[3])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-mailbox", action="store_true", help="N>1: NCCL collectives instead of the peer-memory mailboxes")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU (weak) or in total (strong); default: the workload's")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=4000, help="frames of the bounded cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=2000, help="frames per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary points (c2, 10 % non-detecting frames)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.frames is None:
        args.frames = WORKLOADS[args.workload]["frames"]
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    args.frames_per_gpu = args.frames if args.scaling == "weak" else -(-args.frames // max(world_env if args.impl == "b200" else args.gpus, 1))
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fmcw_radar_processing_b200 import synth
    from fmcw_radar_processing_b200.api import FmcwCuda
    from fmcw_radar_processing_b200.distributed import ShardedRun

    world, rank = world_env, int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    n = args.frames_per_gpu
    if args.scaling == "strong":      # contiguous ranges of one recording of args.frames frames; the last rank may hold fewer
        n = max(0, min(n, args.frames - rank * n))
    counts = [n] * world if args.scaling == "weak" else [max(0, min(args.frames_per_gpu, args.frames - r * args.frames_per_gpu)) for r in range(world)]
    run = DeviceRun(args.workload, n, rank, world, local_rank, frame0=sum(counts[:rank]))
    h, iq, out, inten = run.h, run.iq, run.out, run.inten
    sampler = ClockSampler(local_rank)
    sampler.start()

    sharded = ShardedRun(h, frame_counts=counts) if world > 1 else None
    peer_mailbox = bool(sharded.use_peer_mailbox()) if (sharded is not None and not args.no_mailbox) else False
    shard_check = None
    if sharded is not None:
        shard_check = check_sharded_equals_single(args, rank, world, local_rank, dev, peer_mailbox, barrier)

    def step_device():
        if sharded is None:
            h.run(iq, out, inten)
        else:
            sharded.step_async(iq, out, inten)

    # ---- device-resident timing: `value` ----
    dev_ms, stage_ms = time_device(run, step_device, args.steps, args.warmup, barrier, sampler)
    info = h.info()
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    total_frames = sum(counts)
    value = total_frames / (dev_ms / 1e3)

    # ---- end to end through the C ABI with pinned host buffers: `e2e` ----
    e2e = None
    if not args.no_e2e:
        # two pinned output sets + the pinned input per rank must fit the host; if they do not, the e2e arm runs the same
        # workload on fewer frames per GPU (it is PCIe bound at a constant number of bytes per frame) and says so
        per_frame = 2 * (PN * 4096 + 256 * 4 + 16 * 8 + PN * 4 + 16) + PN * NTS * 4
        avail = host_mem_available()
        m = n
        if avail is not None and per_frame * n * world > 0.55 * avail:
            m = max(1000, int(0.55 * avail / world / per_frame) // 1000 * 1000)
        if world > 1:
            t = torch.tensor([m], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            m = int(t.item())
        if m < n:
            run_e = DeviceRun(args.workload, m, rank, world, local_rank, frame0=rank * m)
            sh_e = None
            if world > 1:
                sh_e = ShardedRun(run_e.h, frame_counts=[m] * world)
                if peer_mailbox:
                    sh_e.use_peer_mailbox()
                sh_e.step_async(run_e.iq, run_e.out, run_e.inten)
            else:
                run_e.h.run(run_e.iq, run_e.out, run_e.inten)
            barrier()
            e2e = measure_e2e(args, run_e, sh_e, world, barrier, sampler, m * world, run_e.h.info())
            e2e["frames_per_gpu"] = m
            e2e["note"] = (f"host memory ({avail / 1e9:.0f} GB available) cannot pin two output sets of {n} frames per GPU x {world}: "
                           f"the e2e arm runs {m} frames per GPU of the same workload")
            run_e.close()
        else:
            e2e = measure_e2e(args, run, sharded, world, barrier, sampler, total_frames, info)
            e2e["frames_per_gpu"] = n

    # ---- secondary points (single GPU only): C2, and the headline with 10 % of the frames without a detection ----
    extra = {}
    if world == 1 and not args.no_extra:
        run.close()
        del h, iq, out, inten
        if args.workload != "c2":
            r2 = DeviceRun("c2", WORKLOADS["c2"]["frames"], 0, 1, local_rank)
            ms2, st2 = time_device(r2, lambda: r2.h.run(r2.iq, r2.out, r2.inten), 20, 3, barrier)
            extra["c2"] = {"workload": WORKLOADS["c2"]["text"] + f"; {r2.n} frames", "value": r2.n / (ms2 / 1e3), "unit": UNIT,
                           "ms_per_step": ms2, "steps": 20, "stage_ms": stage_dict(st2),
                           "chain_frac_of_peak": None}
            extra["c2"]["chain_gbs"] = r2.n * (CHAIN_BYTES_PER_FRAME + STFT_BYTES_PER_FRAME) / (ms2 * 1e-3) / 1e9
            r2.close()
        nm = min(n, 50000)
        r3 = DeviceRun(args.workload, nm, 0, 1, local_rank, miss_every=10)
        ms3, st3 = time_device(r3, lambda: r3.h.run(r3.iq, r3.out, r3.inten), 10, 3, barrier)
        i3 = r3.h.info()
        r3.close()
        r4 = DeviceRun(args.workload, nm, 0, 1, local_rank)
        ms4, st4 = time_device(r4, lambda: r4.h.run(r4.iq, r4.out, r4.inten), 10, 3, barrier)
        r4.close()
        extra["miss10"] = {"what": f"{nm} frames of the headline workload with every 10th frame DC-only (no detection): the STFT plan "
                                   f"made ahead for 'every frame detects' is discarded and re-made on the device after the compaction",
                           "frames": nm, "n_detected": i3["n_detected"], "ms_per_step": ms3, "stage_ms": stage_dict(st3),
                           "frames_per_s": nm / (ms3 / 1e3),
                           "all_detect_same_size": {"ms_per_step": ms4, "stage_ms": stage_dict(st4), "frames_per_s": nm / (ms4 / 1e3)}}

    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines (SURVEY 8(d) algorithmic bytes): the dominant kernel, and the whole chain ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ncl, L_local = info["ncol_local"], info["L_local"]
    stft_bytes = L_local * 4 + ncl * 1024 * 4
    achieved = stft_bytes / (stage_ms[3] * 1e-3) / 1e9 if stage_ms[3] > 0 else None
    chain_gbs = (max(counts) * (CHAIN_BYTES_PER_FRAME + STFT_BYTES_PER_FRAME)) / (dev_ms * 1e-3) / 1e9     # per GPU
    roofline = {"kernel": "stft_tc_kernel (tcgen05 STFT main)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(stft_bytes), "avg_launch_ms": float(stage_ms[3]),
                "stage_ms": stage_dict(stage_ms),
                "chain": {"what": "whole step (frame chain + compaction + plan/max + STFT) per GPU", "bound": "hbm",
                          "algorithmic_bytes_per_frame": CHAIN_BYTES_PER_FRAME + STFT_BYTES_PER_FRAME,
                          "achieved": chain_gbs, "peak": peak, "unit": "GB/s", "frac": chain_gbs / peak,
                          "frac_of_nominal_8tbs": chain_gbs / 8000.0},
                "frame_chain_kernel": {"kernel": "frame_chain_warp_kernel", "algorithmic_bytes_per_frame": CHAIN_BYTES_PER_FRAME,
                                       "achieved": max(counts) * CHAIN_BYTES_PER_FRAME / (stage_ms[0] * 1e-3) / 1e9 if stage_ms[0] > 0 else None,
                                       "peak": peak, "unit": "GB/s", "note": "instruction-issue / FMA-pipe bound, not HBM bound (DESIGN.md 3.1)"},
                "chain_gbs": chain_gbs, "chain_frac_of_peak": chain_gbs / peak}
    # the write stream this kernel produces, issued alone by every SM (tests/cuda/tma_store_probe.cu, profiles/tma_store_probe_r2.txt)
    WRITE_CEILING_GBS = 6060.0
    if achieved:
        roofline["write_ceiling"] = {"value": WRITE_CEILING_GBS, "unit": "GB/s", "frac": achieved / WRITE_CEILING_GBS,
                                     "source": "profiles/tma_store_probe_r2.txt: the same 3-D box stores with idle SMs"}
    if roofline["frame_chain_kernel"]["achieved"]:
        roofline["frame_chain_kernel"]["frac"] = roofline["frame_chain_kernel"]["achieved"] / peak
    traffic_path = os.path.join(ROOT, "profiles", "stft_tc_traffic.json")
    if os.path.exists(traffic_path):
        # ncu --set full captures the kernel at C2 size (1.31 GB per launch); dram bytes scale with the output, so the measured
        # ratio to the algorithmic bytes carries over to this workload's launch
        tr = json.load(open(traffic_path))
        ratio = tr.get("dram_bytes_per_launch", 0) / max(1, tr.get("algorithmic_bytes_per_launch", 1))
        roofline["traffic"] = int(ratio * stft_bytes) if ratio > 0 else None
        roofline["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum = %.3f x algorithmic bytes in the ncu --set full capture at "
                                      "C2 size (profiles/ncu_stft_r2.txt: %d B for %d B), scaled to this launch"
                                      % (ratio, tr.get("dram_bytes_per_launch", 0), tr.get("algorithmic_bytes_per_launch", 0)))
    if "c2" in extra:
        extra["c2"]["chain_frac_of_peak"] = extra["c2"]["chain_gbs"] / peak

    cb = None
    if not args.no_cpu_baseline:
        v, dt = cpu_baseline(run.sx, run.cfg, run.scene, args.cpu_sample, info["L_total"], threads=1)
        cb = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
              "sample": f"first {args.cpu_sample} frames: serial float64 frame loop (RP:197-261) + restated STFT of their "
                        f"columns on the full recording's fine grid; NumPy/SciPy oracle, 1 thread"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": cb,
            "e2e": e2e, "gpu_launches": kernels_per_step(max(counts), world, peer_mailbox) * args.steps,
            "clocks": sampler.summary(),
            "info": {k: info[k] for k in ("n_detected", "L_total", "nfft", "ncol_total", "ncol_local", "n_dtft_bins", "n_refined")}}
    line.update(extra)
    if world > 1:
        line["config"]["shard_exchange"] = ("peer-memory mailboxes over NVLink (headers + max inside the kernels); NCCL for the "
                                            "track gather" if peer_mailbox else "NCCL all-gather + all-reduce + track gather")
        line["shard_check"] = shard_check
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def check_sharded_equals_single(args, rank, world, local_rank, dev, peer_mailbox, barrier):
    """Once per multi-GPU run: a small recording (world x 300 frames of the workload's scene) through the sharded path must
    reproduce, bit for bit, the spectrogram columns and the track that ONE GPU computes for the whole recording."""
    import torch
    import torch.distributed as dist
    from fmcw_radar_processing_b200.distributed import ShardedRun
    m = 300
    part = DeviceRun(args.workload, m, rank, world, local_rank)
    sh = ShardedRun(part.h, frame_counts=[m] * world)
    if peer_mailbox:
        sh.use_peer_mailbox()
    sh.step_async(part.iq, part.out, part.inten)
    barrier()
    inf = part.h.info()
    whole = DeviceRun(args.workload, m * world, 0, 1, local_rank, frame0=0)
    whole.h.run(whole.iq, whole.out, whole.inten)
    torch.cuda.synchronize(dev)
    b, k = inf["col_begin"], inf["ncol_local"]
    same = bool(torch.equal(part.inten[:k], whole.inten[b:b + k]))
    same_track = bool(torch.equal(part.out["range_bin"], whole.out["range_bin"][rank * m:(rank + 1) * m]) and
                      torch.equal(part.out["doppler_bin"], whole.out["doppler_bin"][rank * m:(rank + 1) * m]))
    t = torch.tensor([1 if (same and same_track) else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    part.close()
    whole.close()
    ok = bool(int(t.item()))
    if not ok:
        raise SystemExit(f"rank {rank}: sharded spectrogram / track differs from the single-GPU run (columns equal: {same}, track equal: {same_track})")
    return {"frames": m * world, "columns_checked_rank0": int(k), "bit_identical_to_single_gpu": ok}


def measure_e2e(args, run, sharded, world, barrier, sampler, total_frames, info):
    """The same call with pinned HOST buffers.  Single GPU: blocking calls, then the streaming form (two handles,
    FMCW_OPT_ASYNC_HOST: the H2D of recording i+1 overlaps the D2H of recording i).  Sharded: two handle / buffer sets per
    rank in the same alternation."""
    import torch
    import torch.distributed as dist
    from fmcw_radar_processing_b200 import _lib as L
    from fmcw_radar_processing_b200.api import FmcwCuda
    h, iq, out, inten, dev, n = run.h, run.iq, run.out, run.inten, run.dev, run.n
    out_bytes = sum(int(np.prod(v.shape)) * v.element_size() for v in out.values())
    ncl = info["ncol_local"]
    iq_h = torch.empty(iq.shape, dtype=torch.int16, pin_memory=True)
    iq_h.copy_(iq)
    iq_np = iq_h.numpy()
    sets = []
    for _ in range(2):
        o_h = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in out.items()}
        i_h = torch.empty((run.cols_cap, 1024), dtype=torch.float32, pin_memory=True)
        sets.append((o_h, i_h))
    n_e2e = max(3, min(args.steps, 8 if n <= 20000 else 4))
    e2e_mode = None
    blocking_ms = None
    if sharded is None:
        out_np = {k: v.numpy() for k, v in sets[0][0].items()}
        inten_np = sets[0][1].numpy()
        for _ in range(2):
            h.run(iq_np, out_np, inten_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            h.run(iq_np, out_np, inten_np)
        barrier()
        blocking_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
        h2 = FmcwCuda(run.cfg, run.calib, device=dev.index, torch_stream_sync=False)
        hs = [h, h2]
        for hh in hs:
            hh.set_option(L.OPT_ASYNC_HOST, 1)
        np_sets = [(iq_np, {k: v.numpy() for k, v in s[0].items()}, s[1].numpy()) for s in sets]

        def run_pipelined(k_steps):
            for i in range(k_steps):
                j = i & 1
                hs[j].synchronize()              # the previous recording on this handle is complete
                hs[j].run(*np_sets[j])
            for hh in hs:
                hh.synchronize()

        run_pipelined(2)
        barrier()
        sampler.active = True
        t0 = time.perf_counter()
        run_pipelined(n_e2e)
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
        sampler.active = False
        assert np.array_equal(np_sets[1][2][:1000], np_sets[0][2][:1000])      # both handles produced the same spectrogram
        h.set_option(L.OPT_ASYNC_HOST, 0)
        h2.close()
        e2e_mode = ("streaming: two handles alternate recordings (FMCW_OPT_ASYNC_HOST), H2D of step i+1 overlaps D2H of step i; "
                    "every step copies its own inputs and outputs")
    else:
        # sharded streaming: copy stream in, library stream compute, copy stream out; two device/host buffer sets alternate
        cp_in, cp_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        dsets = [(iq, out, inten), (torch.empty_like(iq), {k: torch.empty_like(v) for k, v in out.items()}, torch.empty_like(inten))]
        cur = torch.cuda.current_stream(dev)
        ncl_ = max(1, ncl)
        ev_done = [None, None]

        dbg = os.environ.get("FMCW_E2E_DEBUG", "")

        def one(i):
            j = i & 1
            d_iq, d_out, d_int = dsets[j]
            if ev_done[j] is not None:
                cp_in.wait_event(ev_done[j])      # the D2H of the step that last used this set is complete
            with torch.cuda.stream(cp_in):
                if "noh2d" not in dbg:
                    d_iq.copy_(iq_h, non_blocking=True)
            cur.wait_stream(cp_in)
            run.stream.wait_stream(cp_in)
            sharded.step_async(d_iq, d_out, d_int)
            cp_out.wait_stream(run.stream)
            cp_out.wait_stream(cur)
            with torch.cuda.stream(cp_out):
                if "nod2h" not in dbg:
                    sets[j][1][:ncl_].copy_(d_int[:ncl_], non_blocking=True)
                for k in d_out:
                    sets[j][0][k].copy_(d_out[k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cp_out)
            ev_done[j] = ev

        for i in range(2):
            one(i)
        barrier()
        sampler.active = True
        t0 = time.perf_counter()
        for i in range(n_e2e):
            one(i)
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
        sampler.active = False
        e2e_mode = ("streaming per rank: H2D on a copy stream, sharded step on the library stream, D2H on a second copy stream, two "
                    "buffer sets alternate; every step copies its own inputs and outputs")
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = n * PN * NTS * 4          # only the processed RX crosses PCIe (cudaMemcpy2D in the library)
    d2h = ncl * 1024 * 4 + out_bytes
    res = {"value": total_frames / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": n_e2e,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "mode": e2e_mode,
           "pcie_gbs_per_gpu": (h2d + d2h) / (e2e_ms * 1e-3) / 1e9,
           "timing": "host wall clock around K calls with a device synchronise on both sides"}
    if blocking_ms is not None:
        res["blocking_ms_per_step"] = blocking_ms
        res["blocking_value"] = total_frames / (blocking_ms / 1e3)
    return res


if __name__ == "__main__":
    main()
