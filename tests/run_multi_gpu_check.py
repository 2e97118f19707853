"""Multi-GPU check of the frame-sharded path (run under torchrun, one rank per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tests/run_multi_gpu_check.py [--frames 600]

Every rank processes its contiguous frame range of ONE synthetic recording.  The spectrogram of the peer-memory
mailbox path must equal the NCCL-collective path bit for bit, both must agree with a single-GPU run of the whole
recording (rank 0 computes it), and the gathered track must be the concatenation of the per-rank tracks.
Prints one line per rank and exits non-zero on a mismatch.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload, scene_tables          # noqa: E402
from fmcw_radar_processing_b200 import synth            # noqa: E402
from fmcw_radar_processing_b200.api import FmcwCuda     # noqa: E402
from fmcw_radar_processing_b200.distributed import ShardedRun   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=600, help="frames per rank")
    args = ap.parse_args()
    world, rank, local_rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    n, PN, NTS, n_rx = args.frames, 64, 128, 3
    sx, cfg, scene = build_workload("c2")
    calib = synth.default_calib(n_rx, NTS) / 4095.0
    h = FmcwCuda(cfg, calib, device=local_rank)

    def make_iq(frame0, count):
        tab = scene_tables(scene, cfg, frame0, count)
        iq = torch.empty((count, n_rx, PN, NTS, 2), dtype=torch.int16, device=dev)
        h.synth_frames(tab, scene.seed, frame0, sigma=scene.sigma, dc=scene.dc, rx_step=scene.rx_step, out=iq)
        return iq

    iq = make_iq(rank * n, n)
    out = h.alloc_frame_out(n, device=dev)
    cap = h.max_cols(n) + 20
    run = ShardedRun(h, frame_counts=[n] * world)

    # NCCL collectives
    inten_nccl = torch.zeros((cap, 1024), dtype=torch.float32, device=dev)
    res = run.step_async(iq, out, inten_nccl)
    torch.cuda.synchronize(dev)
    info_nccl = h.info()
    track_nccl = res["track"].clone()

    # peer-memory mailboxes, four passes (reusing the mailboxes; the third captures the exchange as a CUDA graph, the fourth replays it)
    used = run.use_peer_mailbox()
    ok = True
    if used:
        for _ in range(4):
            inten_mb = torch.full((cap, 1024), -1.0, dtype=torch.float32, device=dev)
            res = run.step_async(iq, out, inten_mb)
            torch.cuda.synchronize(dev)
            info_mb = h.info()
            nloc = info_mb["ncol_local"]
            same = (info_mb["L_total"] == info_nccl["L_total"] and nloc == info_nccl["ncol_local"]
                    and info_mb["col_begin"] == info_nccl["col_begin"] and info_mb["pmax_raw"] == info_nccl["pmax_raw"])
            a, b = inten_mb[:nloc], inten_nccl[:nloc]
            same = same and bool(torch.equal(torch.nan_to_num(a, nan=123.0), torch.nan_to_num(b, nan=123.0)))
            same = same and bool(torch.equal(res["track"], track_nccl))
            ok = ok and same

    # single-GPU run of the whole recording on rank 0, compared with this rank's columns
    nloc, cb = info_nccl["ncol_local"], info_nccl["col_begin"]
    mine = inten_nccl[:nloc].cpu()
    gathered = [None] * world
    dist.gather_object((cb, mine.numpy()), gathered if rank == 0 else None, dst=0)
    worst = 0.0
    if rank == 0:
        h1 = FmcwCuda(cfg, calib, device=local_rank)
        o1, full = h1.run(make_iq(0, n * world))
        info1 = h1.info()
        full = full[:info1["ncol_total"]].cpu().numpy()
        pos = 0
        for cb_r, part in gathered:
            assert cb_r == pos, (cb_r, pos)
            ref = full[pos:pos + part.shape[0]]
            fin = np.isfinite(ref)
            ok = ok and np.array_equal(np.isfinite(part), fin)
            if fin.any():
                worst = max(worst, float(np.max(np.abs(part[fin] - ref[fin]))))
            pos += part.shape[0]
        ok = ok and pos == info1["ncol_total"] and info1["L_total"] == info_nccl["L_total"]
        ok = ok and worst <= 2e-4
        tr = track_nccl.cpu().numpy()                       # [world, 3, n]
        ok = ok and np.array_equal(tr[:, 0].reshape(-1).astype(np.int32), o1["range_bin"].cpu().numpy())
        h1.close()
    print(f"rank {rank}/{world}: mailbox={'on' if used else 'unavailable'} ncol_local={nloc} col_begin={cb} "
          f"L_total={info_nccl['L_total']} max|sharded - single|={worst:.2e} dB -> {'OK' if ok else 'MISMATCH'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
