"""Fleet of independent radars on one GPU (BASELINE config C5: every radar has its own slow-time history, its own
``nfft`` and its own normalisation maximum; whole radars per GPU, no exchange between them).

A small recording (C1 shape: 500 frames) cannot fill a B200: its frame-chain grid is below one resident set and the
plan / max kernels are launch bound.  ``Fleet`` keeps a pool of handles -- each with its own stream, tables and
scratch -- deals the recordings round-robin and leaves every hand-off on the device, so the kernels of different
radars overlap wherever the SMs have room.  The reference has no counterpart (one recording per MATLAB call).
"""
from __future__ import annotations

from .api import FmcwCuda


class Fleet:
    def __init__(self, cfg: dict, calib, n_handles: int = 8, device: int = 0):
        import torch
        self.device = torch.device("cuda", device)
        self.handles = [FmcwCuda(cfg, calib, device=device, torch_stream_sync=False) for _ in range(max(1, n_handles))]
        self._streams = [torch.cuda.ExternalStream(h.stream, device=self.device) for h in self.handles]
        self._bufs = {}                                # radar index -> (n_frames, outputs, intensity), reused by every pass

    def close(self):
        for h in self.handles:
            h.close()
        self.handles = []

    def run(self, recordings, layout: int = 0):
        """``recordings``: int16 CUDA tensors ``[n][rx][PN][NTS][2]``, one per radar (inputs must be complete on the
        current torch stream).  Returns one dict per radar: the per-frame outputs, ``intensity`` (device tensor whose
        first ``ncol`` rows are valid) and the run ``info``.  The result buffers belong to the fleet and are reused by
        the next pass (no allocation, hence no implicit device synchronisation, in the steady state).  Everything is
        enqueued first; a handle is synchronised only when its slot is needed again or at the end."""
        import torch
        cur = torch.cuda.current_stream(self.device)
        h0 = self.handles[0]
        for i, iq in enumerate(recordings):
            n = int(iq.shape[0])
            b = self._bufs.get(i)
            if b is None or b[0] != n or b[3] != layout:
                cols = max(1, h0.max_cols(n))
                shape = (cols, h0.nq) if layout == 0 else (h0.nq, cols)
                self._bufs[i] = (n, h0.alloc_frame_out(n, device=self.device),
                                 torch.empty(shape, dtype=torch.float32, device=self.device), layout)
        for s in self._streams:
            s.wait_stream(cur)                       # inputs are ready
        nh = len(self.handles)
        results = [None] * len(recordings)
        pending = [None] * nh                        # per handle: (radar index, outputs, intensity)

        def collect(slot):
            idx, out, inten = pending[slot]
            info = self.handles[slot].info()         # synchronises this handle's stream
            results[idx] = dict(out, intensity=inten, info=info, ncol=info["ncol_local"])
            pending[slot] = None

        for i, iq in enumerate(recordings):
            slot = i % nh
            if pending[slot] is not None:
                collect(slot)
            _, out, inten, _ = self._bufs[i]
            self.handles[slot].run(iq, out, inten, layout=layout)
            pending[slot] = (i, out, inten)
        for slot in range(nh):
            if pending[slot] is not None:
                collect(slot)
        for s in self._streams:
            cur.wait_stream(s)
        return results
