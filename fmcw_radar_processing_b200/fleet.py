"""Fleet of independent radars on one GPU (BASELINE config C5: every radar has its own slow-time history, its own
``nfft`` and its own normalisation maximum; whole radars per GPU, no exchange between them).

A small recording (C1 shape: 500 frames) cannot fill a B200: its frame-chain grid is below one resident set and the
plan / max kernels are launch bound.  ``Fleet`` keeps a pool of handles -- each with its own stream, tables and
scratch -- deals the recordings round-robin and leaves every hand-off on the device, so the kernels of different
radars overlap wherever the SMs have room.  A pass is bound by the GPU's launch rate, not by its work (about 12 launches
per recording, ~50 us apiece however many streams carry them), so every handle records each recording's whole run as one
CUDA graph (``FMCW_OPT_RUN_GRAPH``) and the per-recording scalars stay on the device (``fmcw_set_info_target``) until one
copy at the end of the pass.  The reference has no counterpart (one recording per MATLAB call).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .api import FmcwCuda


class Fleet:
    def __init__(self, cfg: dict, calib, n_handles: int = 8, device: int = 0):
        import torch
        self.device = torch.device("cuda", device)
        self.handles = [FmcwCuda(cfg, calib, device=device, torch_stream_sync=False) for _ in range(max(1, n_handles))]
        self._streams = [torch.cuda.ExternalStream(h.stream, device=self.device) for h in self.handles]
        self._bufs = {}                                # radar index -> (n_frames, outputs, intensity), reused by every pass
        self._info_size = C.sizeof(_lib.fmcw_device_info)
        self._infos = None                             # device bytes: one fmcw_device_info per recording
        for h in self.handles:
            h.set_option(_lib.OPT_RUN_GRAPH, 1)
            h.set_option(_lib.OPT_STFT_TILES_PER_CTA, 8)

    def close(self):
        for h in self.handles:
            h.close()
        self.handles = []

    def run(self, recordings, layout: int = 0):
        """``recordings``: int16 CUDA tensors ``[n][rx][PN][NTS][2]``, one per radar (inputs must be complete on the
        current torch stream).  Returns one dict per radar: the per-frame outputs, ``intensity`` (device tensor whose
        first ``ncol`` rows are valid) and the run ``info``.  The result buffers belong to the fleet and are reused by
        the next pass (no allocation, hence no implicit device synchronisation, in the steady state).  Everything is
        enqueued without a host round trip; the scalars of all recordings come back in one copy at the end."""
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
        nrec = len(recordings)
        if self._infos is None or self._infos.numel() < nrec * self._info_size:
            self._infos = torch.zeros(nrec * self._info_size, dtype=torch.uint8, device=self.device)
        for s in self._streams:
            s.wait_stream(cur)                       # inputs are ready
        nh = len(self.handles)
        base = self._infos.data_ptr()
        for i, iq in enumerate(recordings):
            h = self.handles[i % nh]
            _, out, inten, _ = self._bufs[i]
            h.set_info_target(base + i * self._info_size)
            h.run(iq, out, inten, layout=layout)
        for s in self._streams:
            cur.wait_stream(s)
        raw = self._infos[:nrec * self._info_size].cpu().numpy()       # waits for every handle (current stream)
        results = []
        for i, (info, status) in enumerate(FmcwCuda.parse_device_infos(raw)):
            if status != 0:
                raise RuntimeError(f"recording {i}: device-side failure {status} (see fmcw_device_info in include/fmcw_cuda.h)")
            _, out, inten, _ = self._bufs[i]
            results.append(dict(out, intensity=inten, info=info, ncol=info["ncol_local"]))
        for h in self.handles:
            h.set_info_target(None)
        return results


class AllRx:
    """``rx_select = all``: every RX antenna of a recording as its own stream.  The reference reads RX 1
    only (RP:202); here one handle per antenna (its own calibration table, slow-time history, ``nfft`` and normalisation
    maximum) works on the SAME frame buffer, each on its own stream, so a 3-RX recording is processed as 3x the frames."""

    def __init__(self, cfg: dict, calib, device: int = 0):
        self.n_rx = int(cfg["num_Rx_antennas"])
        self.fleet = None
        self.handles = []
        for r in range(self.n_rx):
            c = dict(cfg)
            c["rx_select"] = r + 1
            self.handles.append(FmcwCuda(c, calib, device=device, torch_stream_sync=False))

    def close(self):
        for h in self.handles:
            h.close()
        self.handles = []

    def run(self, iq, layout: int = 0):
        """``iq``: int16 ``[n][rx][PN][NTS][2]``, NumPy (host) or a CUDA tensor.  Returns one dict per RX."""
        res = []
        if hasattr(iq, "is_cuda"):
            import torch
            cur = torch.cuda.current_stream(iq.device)
            streams = [torch.cuda.ExternalStream(h.stream, device=iq.device) for h in self.handles]
            for s in streams:
                s.wait_stream(cur)
            outs = [h.run(iq, layout=layout) for h in self.handles]          # enqueued back to back, overlapping on the device
            for h, s, (out, inten) in zip(self.handles, streams, outs):
                info = h.info()
                cur.wait_stream(s)
                res.append(dict(out, intensity=inten, info=info, ncol=info["ncol_local"]))
        else:
            for h in self.handles:
                out, inten = h.run(iq, layout=layout)
                info = h.info()
                res.append(dict(out, intensity=inten, info=info, ncol=info["ncol_local"]))
        return res
