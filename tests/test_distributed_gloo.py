"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: head/halo exchange, offsets, global max,
track gather and column ownership (fmcw_radar_processing_b200/distributed.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fmcw_radar_processing_b200 import distributed as D


def _worker(rank, world, port, lengths, win, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(sum(lengths), dtype=torch.float64) * 0.5 + 1.0        # the global slow-time signal
        off = sum(lengths[:rank])
        x = full[off:off + lengths[rank]]
        lay = D.exchange_heads(x[:win - 1].contiguous(), lengths[rank], win)
        pm = D.allreduce_max(float(rank + 1) * 10.0, torch.device("cpu"))
        nfr = [3, 5, 2, 4][:world]
        rb = torch.arange(nfr[rank], dtype=torch.int32) + 100 * rank
        tr = D.gather_track(rb, rb + 1, rb.to(torch.float32) * 0.25, nfr)
        q.put((rank, lay.lengths, lay.offsets, lay.L_total, lay.halo.tolist(), pm, [t.tolist() for t in tr]))
    finally:
        dist.destroy_process_group()


def _run(world, lengths, win, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lengths, win, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("lengths", [[640, 320], [64, 5], [0, 128], [7, 0]])
def test_heads_halo_offsets_world2(lengths):
    win = 20
    res = _run(2, lengths, win, 29500 + sum(lengths) % 200)
    full = np.arange(sum(lengths)) * 0.5 + 1.0
    for rank, lens, offs, L, halo, pm, tr in res:
        assert lens == lengths and offs == [0, lengths[0]] and L == sum(lengths)
        end = offs[rank] + lengths[rank]
        assert np.allclose(halo, full[end:end + win - 1])           # the next win-1 samples, across shard borders
        assert pm == 20.0
        assert tr[0] == [0, 1, 2, 100, 101, 102, 103, 104] and tr[1][0] == 1 and tr[2][3] == 25.0


def test_halo_spans_several_short_shards_world4():
    lengths, win = [100, 6, 0, 50], 20
    res = _run(4, lengths, win, 29791)
    full = np.arange(sum(lengths)) * 0.5 + 1.0
    for rank, lens, offs, L, halo, pm, tr in res:
        end = offs[rank] + lengths[rank]
        assert np.allclose(halo, full[end:end + win - 1])
        assert pm == 40.0


def test_column_ownership_partitions_all_columns():
    win, ov = 20, 19
    for lengths in ([640, 320, 0, 64], [5, 5, 5, 5, 100]):
        L = sum(lengths)
        offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
        cols = []
        for o, l in zip(offs, lengths):
            b, e, ncol = D.owned_columns(int(o), int(l), L, win, ov)
            cols += list(range(b, e))
        assert cols == list(range(L - ov))
    # hop > 1
    for hop in (3, 10):
        win, ov = 32, 32 - hop
        lengths = [101, 57, 300]
        L = sum(lengths)
        offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
        cols = []
        for o, l in zip(offs, lengths):
            b, e, ncol = D.owned_columns(int(o), int(l), L, win, ov)
            cols += list(range(b, e))
        assert cols == list(range((L - ov) // hop))
