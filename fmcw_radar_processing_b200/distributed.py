"""Frame-sharded multi-GPU run: one process per GPU, contiguous frame ranges,
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) for the plumbing.

The per-frame chain (RP:197-261) needs no communication.  The STFT (RP:270-299) couples the shards
through the concatenated slow-time signal only, so exactly three small collectives remain:

1. one all-gather of ``[local length, first window_length-1 samples]`` per rank: gives every
   shard its global sample offset, the global length L (hence nfft, RP:273) and the halo it needs from the
   shard(s) to its right -- also when a neighbour detected fewer than window_length-1 samples;
2. one all-reduce(max) of the scalar normalisation max(P) (RP:282-283);
3. one all-gather of the per-frame track (range bin, Doppler bin, strength) for the range/speed payload.

On a CUDA node whose GPUs can map each other's memory (NVLink/NVSwitch), collectives 1 and 2 are replaced by
peer-memory mailboxes (``PeerMailbox``): the kernels of every rank store the header / the maximum straight into
all ranks' mailboxes and the consumers wait on flags on the device, so a pass has no collective call and no
stream hand-off on its critical path; NCCL remains for the track gather (3), which overlaps the STFT.

Column ownership: a spectrogram column belongs to the shard that owns its first sample, so columns are
neither duplicated nor lost.  Everything in this file except ``ShardedRun`` is backend agnostic and is
covered by world_size-2 gloo tests on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


@dataclass
class ShardLayout:
    lengths: list          # local slow-time lengths of every rank
    offsets: list          # global index of every rank's first sample
    L_total: int
    halo: torch.Tensor     # the samples following this rank's last sample (<= window_length-1)


def exchange_heads(x_local: torch.Tensor, L_local: int, window_length: int, group=None) -> ShardLayout:
    """Collective 1.  ``x_local`` holds at least ``min(L_local, window_length-1)`` leading samples of this
    rank's slow-time magnitude signal (float64, on the backend's device)."""
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    hw = window_length - 1
    dev = x_local.device
    msg = torch.zeros(1 + hw, dtype=torch.float64, device=dev)
    msg[0] = float(L_local)                  # exact in float64
    n_head = min(L_local, hw)
    if n_head:
        msg[1:1 + n_head] = x_local[:n_head]
    gathered = [torch.empty_like(msg) for _ in range(ws)]
    dist.all_gather(gathered, msg, group=group)
    heads = torch.stack(gathered)
    lengths = [int(v) for v in heads[:, 0].to("cpu").numpy()]
    offsets = [int(v) for v in np.concatenate([[0], np.cumsum(lengths)[:-1]])]
    # assemble the halo from the heads of the following ranks (skipping short / empty shards)
    parts, need = [], hw
    for r in range(rank + 1, ws):
        if need == 0:
            break
        take = min(need, lengths[r], hw)
        if take:
            parts.append(heads[r, 1:1 + take])
            need -= take
    halo = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.float64, device=dev)
    return ShardLayout(lengths, offsets, int(sum(lengths)), halo)


def allreduce_max(value: float, device, group=None) -> float:
    """Collective 2."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_track(range_bin: torch.Tensor, doppler_bin: torch.Tensor, range_mag: torch.Tensor, frame_counts, group=None):
    """Collective 3: per-frame track of the whole recording on every rank (12 B per frame)."""
    ws = dist.get_world_size(group)
    nmax = max(frame_counts)
    dev = range_bin.device
    msg = torch.zeros(3, nmax, dtype=torch.float32, device=dev)
    n = range_bin.shape[0]
    msg[0, :n] = range_bin.to(torch.float32)
    msg[1, :n] = doppler_bin.to(torch.float32)
    msg[2, :n] = range_mag
    out = [torch.empty_like(msg) for _ in range(ws)]
    dist.all_gather(out, msg, group=group)
    rb = torch.cat([out[r][0, :frame_counts[r]] for r in range(ws)]).to(torch.int32)
    db = torch.cat([out[r][1, :frame_counts[r]] for r in range(ws)]).to(torch.int32)
    mg = torch.cat([out[r][2, :frame_counts[r]] for r in range(ws)])
    return rb, db, mg


def owned_columns(offset: int, L_local: int, L_total: int, window_length: int, overlap: int):
    """[begin, end) of the global spectrogram columns whose first sample lies in this shard."""
    hop = window_length - overlap
    ncol = (L_total - overlap) // hop if L_total >= window_length else 0
    b = -(-offset // hop)
    e = min(ncol, -(-(offset + L_local) // hop))
    return min(b, e), e, ncol


class PeerMailbox:
    """One mailbox per rank in symmetric (peer-mapped) device memory; ``ptrs[r]`` is rank r's mailbox as
    addressable from this process.  Built with ``torch.distributed._symmetric_memory`` (CUDA VMM / IPC
    handles exchanged through the group's store); ``PeerMailbox.create`` returns None when that is not available,
    and callers fall back to the NCCL collectives."""

    def __init__(self, buf, ptrs, rank):
        self.buf, self.ptrs, self.rank, self.step = buf, list(ptrs), rank, 0

    @classmethod
    def create(cls, handle, group=None):
        if not torch.cuda.is_available() or dist.get_backend(group) != "nccl":
            return None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            dev = torch.device("cuda", torch.cuda.current_device())
            n = (handle.mailbox_bytes() + 7) // 8
            buf = symm_mem.empty(n, dtype=torch.float64, device=dev)
            buf.zero_()
            g = group if group is not None else dist.group.WORLD
            hdl = symm_mem.rendezvous(buf, g.group_name)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            torch.cuda.synchronize(dev)          # the mailbox is zeroed (use_peer_mailbox adds the barrier)
            return cls(buf, ptrs, dist.get_rank(group))
        except Exception as e:                   # no peer access / symmetric memory not supported here
            import warnings
            warnings.warn(f"peer-memory mailboxes unavailable, using NCCL collectives: {e}")
            return None

    def next_step(self) -> int:
        self.step += 1
        return self.step


class ShardedRun:
    """One rank's share of a frame-sharded recording on its GPU."""

    def __init__(self, handle, group=None, frame_counts=None):
        self.h = handle
        self.group = group
        self.frame_counts = frame_counts      # frames per rank (static); gathered once if not given
        self.mailbox = None                   # PeerMailbox: set by use_peer_mailbox()

    def use_peer_mailbox(self) -> bool:
        """Collective call (every rank): switches ``step_async`` to the peer-memory hand-offs when the GPUs of
        this node can map each other's memory.  Returns whether they are in use."""
        self.mailbox = PeerMailbox.create(self.h, self.group)
        # all ranks must take the same path
        ok = torch.tensor([1 if self.mailbox is not None else 0], device="cuda" if torch.cuda.is_available() else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            self.mailbox = None
        else:
            dist.barrier(group=self.group)    # every mailbox is zeroed before anyone posts
        return self.mailbox is not None

    def _async_buffers(self, dev):
        if getattr(self, "_bufs", None) is None:
            ws = dist.get_world_size(self.group)
            n = 1 + self.h.cfg["window_length"] - 1
            self._bufs = dict(msg=torch.zeros(n, dtype=torch.float64, device=dev),
                              gathered=torch.zeros(ws * n, dtype=torch.float64, device=dev),
                              gmax=torch.zeros(1, dtype=torch.float64, device=dev),
                              lib_stream=torch.cuda.ExternalStream(self.h.stream, device=dev))
        return self._bufs

    def _gather_track_async(self, out, dev):
        """Collective 3 without host work: one stack kernel into a preallocated message, one all-gather.
        Returns [world, 3, n_max] float32 (range bin, Doppler bin, strength); rows past a rank's frame count
        are padding."""
        ws = dist.get_world_size(self.group)
        n = int(out["range_bin"].shape[0])
        tb = getattr(self, "_track_bufs", None)
        if tb is None or tb[0].shape[1] != n:
            tb = (torch.empty(3, n, dtype=torch.float32, device=dev), torch.empty(ws, 3, n, dtype=torch.float32, device=dev))
            self._track_bufs = tb
        msg, gathered = tb
        msg[0].copy_(out["range_bin"])
        msg[1].copy_(out["doppler_bin"])
        msg[2].copy_(out["range_mag"])
        dist.all_gather_into_tensor(gathered, msg, group=self.group)
        return gathered

    def step_async(self, iq, out, intensity, layout=0, gather=True):
        """Same result as ``step`` with every hand-off in device memory: the host only enqueues kernels and
        collectives (stream-ordered with events), so the step has no host round trip.  Sizes are read
        afterwards with ``handle.info()``."""
        h = self.h
        dev = iq.device
        mb = self.mailbox
        if mb is not None:
            # peer-memory path: three library calls on the handle's stream, the exchanges happen inside the kernels
            step = mb.next_step()
            h.process_frames(iq, out)
            if gather:
                torch.cuda.current_stream(dev).wait_stream(self._async_buffers(dev)["lib_stream"])   # track columns are written
            h.mailbox_post_heads(mb.ptrs, mb.rank, step)
            h.mailbox_plan(mb.ptrs, mb.rank, step)
            track = self._gather_track_async(out, dev) if gather else None              # collective 3 (overlaps plan/STFT)
            h.mailbox_stft(mb.ptrs, mb.rank, step, intensity, layout)
            return dict(track=track)
        b = self._async_buffers(dev)
        lib, cur = b["lib_stream"], torch.cuda.current_stream(dev)
        ws, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        h.process_frames(iq, out)                       # library stream
        h.shard_pack(b["msg"])
        cur.wait_stream(lib)
        dist.all_gather_into_tensor(b["gathered"], b["msg"], group=self.group)          # collective 1
        lib.wait_stream(cur)
        h.shard_plan(b["gathered"], ws, rank, b["gmax"])
        track = self._gather_track_async(out, dev) if gather else None                   # collective 3 (overlaps plan/max)
        cur.wait_stream(lib)
        dist.all_reduce(b["gmax"], op=dist.ReduceOp.MAX, group=self.group)              # collective 2
        lib.wait_stream(cur)
        h.shard_stft(b["gmax"], intensity, layout)
        return dict(track=track)

    def step(self, iq, out, intensity, layout=0, gather=True):
        h = self.h
        dev = iq.device
        h.process_frames(iq, out)
        info = h.info()
        L_local = info["L_local"]
        win = h.cfg["window_length"]
        head = torch.zeros(win - 1, dtype=torch.float64, device=dev)
        n_head = min(L_local, win - 1)
        if n_head:
            h.get_slow_time(head, 0, n_head)
        lay = exchange_heads(head, L_local, win, self.group)
        rank = dist.get_rank(self.group)
        halo = lay.halo.contiguous()
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()     # the library works on its own stream
        h.set_halo(halo, int(halo.numel()))
        if lay.L_total < win:
            return dict(layout=lay, ncol_local=0, col_begin=0, pmax_raw=0.0, track=None)
        local_max = h.stft_local_max(lay.L_total, lay.offsets[rank])
        pmax = allreduce_max(local_max, dev, self.group)
        h.stft_sharded(lay.L_total, lay.offsets[rank], pmax, intensity, layout)
        track = None
        if gather:
            if self.frame_counts is None:
                counts = [None] * dist.get_world_size(self.group)
                dist.all_gather_object(counts, int(iq.shape[0]), group=self.group)
                self.frame_counts = counts
            track = gather_track(out["range_bin"], out["doppler_bin"], out["range_mag"], self.frame_counts, self.group)
        b, e, ncol = owned_columns(lay.offsets[rank], L_local, lay.L_total, win, h.cfg["overlap"])
        return dict(layout=lay, ncol_local=e - b, col_begin=b, ncol_total=ncol, pmax_raw=pmax, track=track)
