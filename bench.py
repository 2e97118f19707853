#!/usr/bin/env python
"""Benchmark of the fused range-Doppler-STFT chain (BASELINE.json metric: radar frames/s and achieved
HBM GB/s), workload C2 of SURVEY.md 8d.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the whole hot path (frame chain -> compaction -> STFT) over one recording of
``--frames`` frames (default 5,000: configs[1]).  ``value`` is measured with the inputs resident in HBM,
``e2e`` through the same C-ABI call with pinned HOST buffers (H2D of the frames and D2H of every output
inside the timed region).  With N > 1 every rank owns a contiguous 5,000-frame shard of one N x 5,000
frame recording (weak scaling); the STFT runs over the whole concatenated signal with a halo
(fmcw_radar_processing_b200/distributed.py).
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
# algorithmic bytes per frame of the C1/C2 shape (1 processed RX, 64 x 128, hop 1)
CHAIN_BYTES_PER_FRAME = 128 * 64 * 4 + 256 * 4 + 16 + 16 * 4 + 64 * 4          # 34,128
STFT_BYTES_PER_FRAME = 64 * 4 + 64 * 1024 * 4                                  # 262,400
KERNELS_PER_STEP = 10  # look-ahead stft_plan + stft_tc_prepare (side stream), frame_chain, compact_fused, stft_plan, stft_tc_prepare (confirm), colstat, refine, hard, stft_tc
KERNELS_PER_STEP_MAILBOX = 13  # the same + mailbox_post_heads, mailbox_post_max, mailbox_collect_max (N > 1, peer-memory path)
KERNELS_PER_STEP_NCCL = 9  # no look-ahead: frame_chain, compact_fused, shard_pack, stft_plan, stft_tc_prepare, colstat, refine, hard, stft_tc


def build_workload(n_frames, n_rx=3):
    from fmcw_radar_processing_b200 import synth
    from fmcw_radar_processing_b200.config import fmcw_configurations
    from fmcw_radar_processing_b200.parse import make_sxml
    sx = make_sxml(numSamplesPerChirp=128, numChirpsPerFrame=64, numAntennasRx=n_rx)
    cfg = fmcw_configurations(sx)
    scene = synth.scene_c2(seed=2)
    return sx, cfg, scene


def scene_tables(scene, cfg, frame0, n):
    from fmcw_radar_processing_b200 import synth
    return synth.scene_tables(scene, cfg["dist_per_bin"], cfg["range_fft_size"], cfg["PRT"], cfg["lambda"], frame0, n)


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
    iq = synth.synth_frames(tab, scene.seed, 0, cfg["num_Rx_antennas"], 64, 128, sigma=scene.sigma, dc=scene.dc,
                            rx_step=scene.rx_step)
    calib = synth.default_calib(cfg["num_Rx_antennas"], 128)
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


def run_reference(args):
    """--impl reference: the reference's own (CPU) implementation of the path.  MATLAB/Octave are absent, so
    this is the oracle port, vectorised and with all host threads (oracle/fmcw_oracle_batched.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fmcw_radar_processing_b200 import synth
    from oracle import fmcw_oracle_batched as OB
    sx, cfg, scene = build_workload(args.frames)
    cores = os.cpu_count() or 1
    n_sample = args.ref_sample
    L_full = args.frames * args.gpus * 64
    tab = scene_tables(scene, cfg, 0, n_sample)
    iq = synth.synth_frames(tab, scene.seed, 0, cfg["num_Rx_antennas"], 64, 128, sigma=scene.sigma, dc=scene.dc,
                            rx_step=scene.rx_step)
    calib = synth.default_calib(cfg["num_Rx_antennas"], 128)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        OB.run_no_branch(iq, calib, sx, L_total=L_full, workers=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    v = n_sample / (ms / 1e3)
    sample = f"first {n_sample} frames of the workload per step: vectorised frame chain + restated STFT of their columns on the full recording's fine grid"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, cfg):
    return {"workload": f"C2 (BASELINE.json configs[1]): multi-target walking-animal scene, 3 RX x 64 chirps x 128 samples, "
                        f"{args.frames} frames per GPU, RX 1 processed as in the reference (RP:202), range-Doppler + hop-1 STFT "
                        f"(window 20, 1024 log-frequency bins)",
            "frames_per_gpu": args.frames, "rx": 3, "chirps": 64, "samples": 128, "stft_window": 20, "stft_hop": 1,
            "intensity_layout": "time-major [col][1024] (MATLAB memory order)",
            "l2": "inputs (164 MB RX-1 samples) and outputs (1.3 GB) per step exceed the 126 MB L2; no explicit flush",
            "parallelism": f"frame-sharded x{args.gpus}" if args.gpus > 1 else "single GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-mailbox", action="store_true", help="N>1: NCCL collectives instead of the peer-memory mailboxes")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=5000)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=4000, help="frames of the bounded cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=2000, help="frames per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fmcw_radar_processing_b200 import synth
    from fmcw_radar_processing_b200.api import FmcwCuda
    from fmcw_radar_processing_b200.distributed import ShardedRun

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sx, cfg, scene = build_workload(args.frames)
    n, PN, NTS, n_rx = args.frames, 64, 128, 3
    h = FmcwCuda(cfg, synth.default_calib(n_rx, NTS) / 4095.0, device=local_rank, torch_stream_sync=False)

    # ---- synthetic input, generated on the device by the counter-based generator ----
    frame0 = rank * n
    tab = scene_tables(scene, cfg, frame0, n)
    iq = torch.empty((n, n_rx, PN, NTS, 2), dtype=torch.int16, device=dev)
    h.synth_frames(tab, scene.seed, frame0, sigma=scene.sigma, dc=scene.dc, rx_step=scene.rx_step, out=iq)
    out = h.alloc_frame_out(n, device=dev)
    cols_cap = h.max_cols(n) + (20 if world > 1 else 0)
    inten = torch.empty((cols_cap, 1024), dtype=torch.float32, device=dev)
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sharded = ShardedRun(h, frame_counts=[n] * world) if world > 1 else None
    peer_mailbox = bool(sharded.use_peer_mailbox()) if (sharded is not None and not args.no_mailbox) else False

    def step_device():
        if sharded is None:
            h.run(iq, out, inten)
        else:
            sharded.step_async(iq, out, inten)

    # ---- device-resident timing: `value` ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    stage_ms = np.zeros(4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active = True
    t_wall0 = time.perf_counter()
    stream.wait_stream(torch.cuda.current_stream(dev))
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    stream.wait_stream(torch.cuda.current_stream(dev))   # NCCL work of the sharded path runs on torch's stream
    e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.active = False
    dev_ms = e0.elapsed_time(e1) / args.steps             # CUDA events on the library's stream
    # per-kernel durations: CUDA events recorded by the library on its own stream around each stage
    for _ in range(args.steps):
        step_device()
        tm = h.timings()
        stage_ms += np.array([tm["chain_ms"], tm["compact_ms"], tm["plan_max_ms"], tm["stft_main_ms"]])
    stage_ms /= args.steps
    info = h.info()
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    total_frames = n * world
    value = total_frames / (dev_ms / 1e3)

    # ---- end to end through the C ABI with pinned host buffers: `e2e` ----
    e2e = None
    if not args.no_e2e:
        iq_h = torch.empty(iq.shape, dtype=torch.int16, pin_memory=True)
        iq_h.copy_(iq)
        out_h = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in out.items()}
        out_np = {k: v.numpy() for k, v in out_h.items()}
        inten_h = torch.empty((cols_cap, 1024), dtype=torch.float32, pin_memory=True)
        iq_np, inten_np = iq_h.numpy(), inten_h.numpy()

        def step_host():
            if sharded is None:
                h.run(iq_np, out_np, inten_np)
            else:   # sharded: frames from host, per-rank spectrogram columns back to host
                iq.copy_(iq_h, non_blocking=True)
                sharded.step_async(iq, out, inten)
                ncl_ = h.info()["ncol_local"]
                inten_h[:max(1, ncl_)].copy_(inten[:max(1, ncl_)], non_blocking=True)
                for k in out:
                    out_h[k].copy_(out[k], non_blocking=True)
                torch.cuda.synchronize(dev)

        n_e2e = max(4, min(args.steps, 10))
        for _ in range(2):
            step_host()
        barrier()
        sampler.active = True
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_host()
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
        sampler.active = False
        blocking_ms = e2e_ms
        pipelined = False
        if sharded is None:
            # streaming form of the same call: two handles (FMCW_OPT_ASYNC_HOST), so that recording i+1's H2D overlaps
            # recording i's D2H on the full-duplex link; every step still moves its own inputs and outputs
            from fmcw_radar_processing_b200 import _lib as L
            h2 = FmcwCuda(cfg, synth.default_calib(n_rx, NTS) / 4095.0, device=local_rank, torch_stream_sync=False)
            hs = [h, h2]
            for hh in hs:
                hh.set_option(L.OPT_ASYNC_HOST, 1)
            out_h2 = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in out.items()}
            inten_h2 = torch.empty((cols_cap, 1024), dtype=torch.float32, pin_memory=True)
            sets = [(iq_np, out_np, inten_np), (iq_np, {k: v.numpy() for k, v in out_h2.items()}, inten_h2.numpy())]

            def run_pipelined(k_steps):
                for i in range(k_steps):
                    j = i & 1
                    hs[j].synchronize()              # the previous recording on this handle is complete
                    hs[j].run(*sets[j])
                for hh in hs:
                    hh.synchronize()

            run_pipelined(4)
            barrier()
            sampler.active = True
            t0 = time.perf_counter()
            run_pipelined(n_e2e)
            barrier()
            e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
            sampler.active = False
            pipelined = True
            assert np.array_equal(sets[1][2][:1000], sets[0][2][:1000])      # both handles produced the same spectrogram
            h.set_option(L.OPT_ASYNC_HOST, 0)
            h2.close()
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        ncl = info["ncol_local"]
        h2d = n * PN * NTS * 4          # only the processed RX crosses PCIe (cudaMemcpy2D in the library)
        d2h = ncl * 1024 * 4 + sum(int(np.prod(v.shape)) * v.element_size() for v in out.values())
        e2e = {"value": total_frames / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": n_e2e,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "mode": ("streaming: two handles alternate recordings (FMCW_OPT_ASYNC_HOST), H2D of step i+1 overlaps D2H of step i; "
                        "every step copies its own inputs and outputs") if pipelined else "blocking calls",
               "blocking_ms_per_step": blocking_ms, "blocking_value": total_frames / (blocking_ms / 1e3),
               "timing": "host wall clock around K calls with a device synchronise on both sides"}

    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (stft_main_kernel), SURVEY 8(d) algorithmic bytes ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ncl, L_local = info["ncol_local"], info["L_local"]
    stft_bytes = L_local * 4 + ncl * 1024 * 4
    achieved = stft_bytes / (stage_ms[3] * 1e-3) / 1e9 if stage_ms[3] > 0 else None
    roofline = {"kernel": "stft_tc_kernel (tcgen05 STFT main)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(stft_bytes), "avg_launch_ms": float(stage_ms[3]),
                "stage_ms": {"frame_chain": float(stage_ms[0]), "compaction": float(stage_ms[1]),
                             "stft_plan_and_max": float(stage_ms[2]), "stft_main": float(stage_ms[3])},
                "chain_gbs": (n * (CHAIN_BYTES_PER_FRAME + STFT_BYTES_PER_FRAME) / (dev_ms / world * 1e-3) / 1e9) if world == 1 else
                             (total_frames * (CHAIN_BYTES_PER_FRAME + STFT_BYTES_PER_FRAME) / world / (dev_ms * 1e-3) / 1e9),
                "chain_frac_of_peak": None}
    roofline["chain_frac_of_peak"] = roofline["chain_gbs"] / peak
    traffic_path = os.path.join(ROOT, "profiles", "stft_tc_traffic.json")
    if os.path.exists(traffic_path):
        roofline["traffic"] = json.load(open(traffic_path)).get("dram_bytes_per_launch")

    cb = None
    if not args.no_cpu_baseline:
        v, dt = cpu_baseline(sx, cfg, scene, args.cpu_sample, info["L_total"], threads=1)
        cb = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
              "sample": f"first {args.cpu_sample} frames: serial float64 frame loop (RP:197-261) + restated STFT of their "
                        f"columns on the full recording's fine grid; NumPy/SciPy oracle, 1 thread"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, cfg), "roofline": roofline, "cpu_baseline": cb,
            "e2e": e2e, "gpu_launches": (KERNELS_PER_STEP if world == 1 else KERNELS_PER_STEP_MAILBOX if peer_mailbox else KERNELS_PER_STEP_NCCL) * args.steps, "clocks": sampler.summary(),
            "info": {k: info[k] for k in ("n_detected", "L_total", "nfft", "ncol_total", "ncol_local", "n_dtft_bins", "n_refined")}}
    if world > 1:
        line["config"]["shard_exchange"] = ("peer-memory mailboxes over NVLink (headers + max inside the kernels); NCCL for the "
                                            "track gather" if peer_mailbox else "NCCL all-gather + all-reduce + track gather")
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
