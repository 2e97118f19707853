"""Streaming driver for recordings that do not fit a GPU (BASELINE configs[3]: 4 RX x 256 chirps x 256 samples, 2,000,000
frames = 2.1 TB of samples in and, at the reference's hop 1, 2.1 TB of spectrogram out).

The per-frame chain (RP:197-261) is independent per frame, so frames are processed chunk by chunk and only what couples
them is kept: the slow-time magnitude row of every detected frame (8 B per chirp: 4 GB for the whole recording) and the
per-frame track.  The STFT (RP:270-299) needs two global scalars -- nfft = 2^nextpow2(L) (RP:273) and max(P) (RP:282) --
so it runs in two passes over the kept signal: pass A finds the maximum piece by piece, pass B writes the spectrogram
piece by piece into a buffer the caller drains (to the host, to a file, to a consumer kernel).  A piece is handled exactly
like a shard of the multi-GPU path (its samples + the window_length-1 samples that follow it; a column belongs to the
piece that owns its first sample), so the result equals one run over the whole recording (same bins, same columns; the two
searches of the maximum agree to 1e-6 relative = 9e-6 dB).

With ``torch.distributed`` initialised, every rank streams its contiguous frame range; the ranks meet twice per
recording (heads / lengths, then the maximum), not per chunk.
"""
from __future__ import annotations

import numpy as np

from .api import FmcwCuda


class StreamingRecording:
    def __init__(self, cfg: dict, calib, device: int = 0, group=None, distributed: bool = False):
        import torch
        self.torch = torch
        self.h = FmcwCuda(cfg, calib, device=device)
        self.cfg, self.PN, self.win = cfg, cfg["num_chirps_per_frame"], cfg["window_length"]
        self.hop = cfg["window_length"] - cfg["overlap"]
        self.dev = torch.device("cuda", device)
        self.group, self.distributed = group, distributed
        self.x = None                   # kept slow-time magnitudes of this rank's detected frames (float64, device)
        self.L = 0
        self.track = []                 # per chunk: dict of per-frame outputs (host)

    def close(self):
        self.h.close()

    # ---- pass 1: frames, chunk by chunk ----
    def reserve(self, max_frames: int):
        self.x = self.torch.empty(max_frames * self.PN + self.win, dtype=self.torch.float64, device=self.dev)
        self.L = 0
        self.track = []

    def push_frames(self, iq, out=None, keep_track: bool = True):
        """One chunk of frames (int16 [n][rx][PN][NTS][2], host or device) through RP:197-261; its slow-time rows are
        appended to the kept signal."""
        o = self.h.process_frames(iq, out)
        L = self.h.info()["L_local"]
        if L:
            self.h.get_slow_time(self.x[self.L:self.L + L], 0, L)
        self.L += L
        if keep_track:
            self.track.append({k: (v.cpu().numpy() if hasattr(v, "cpu") else np.array(v)) for k, v in o.items()
                               if k in ("detected", "range_bin", "doppler_bin", "range_mag")})
        return o

    # ---- the two exchanges of a sharded recording ----
    def _layout(self):
        """Global length, this rank's sample offset and the halo that follows this rank's last sample."""
        torch = self.torch
        hw = self.win - 1
        if not self.distributed:
            return self.L, 0, torch.zeros(0, dtype=torch.float64, device=self.dev)
        from .distributed import exchange_heads
        import torch.distributed as dist
        lay = exchange_heads(self.x[:min(self.L, hw)], self.L, self.win, self.group)
        return lay.L_total, lay.offsets[dist.get_rank(self.group)], lay.halo.to(self.dev)

    def _pieces(self, piece_cols: int):
        step = max(self.PN, (piece_cols * self.hop) // self.PN * self.PN)         # samples per piece, whole frames
        a = 0
        while a < self.L:
            b = min(self.L, a + step)
            yield a, b
            a = b

    def _load(self, a, b, halo_tail):
        """Piece [a, b) of the kept signal + the window_length-1 samples that follow it."""
        hw = self.win - 1
        n_halo = min(hw, self.L - b)
        if n_halo < hw and halo_tail.numel():             # the last piece borrows from the next rank(s)
            k = min(hw - n_halo, halo_tail.numel())
            self.x[self.L:self.L + k] = halo_tail[:k]
            n_halo += k
        self.h.load_slow_time(self.x[a:b + n_halo], b - a, n_halo)

    # ---- pass 2 + 3: the STFT over the kept signal ----
    def stft(self, out_buf, piece_cols: int | None = None, consumer=None):
        """Writes the spectrogram piece by piece into ``out_buf`` ([cols][1024] float32, device) and calls
        ``consumer(col_begin, ncol, out_buf)`` after each piece (the buffer is reused).  Returns a summary dict."""
        torch = self.torch
        piece_cols = piece_cols or out_buf.shape[0] - self.win
        L_total, offset, halo_tail = self._layout()
        if L_total < self.win:
            return dict(L_total=L_total, ncol_total=0, pieces=0, pmax_raw=0.0)
        import time
        pm = 0.0
        pieces = list(self._pieces(piece_cols))
        t0 = time.perf_counter()
        for a, b in pieces:                               # pass A: the global maximum (RP:282)
            self._load(a, b, halo_tail)
            pm = max(pm, self.h.stft_local_max(L_total, offset + a))
        self.seconds = {"max_pass": time.perf_counter() - t0}
        if self.distributed:
            from .distributed import allreduce_max
            pm = allreduce_max(pm, self.dev, self.group)
        ncol = 0
        t0 = time.perf_counter()
        t_cons = 0.0
        for a, b in pieces:                               # pass B: the spectrogram (RP:283-299)
            self._load(a, b, halo_tail)
            self.h.stft_sharded(L_total, offset + a, pm, out_buf)
            info = self.h.info()
            ncol += info["ncol_local"]
            if consumer is not None:
                tc = time.perf_counter()
                consumer(info["col_begin"], info["ncol_local"], out_buf)
                t_cons += time.perf_counter() - tc
        self.seconds.update(spectrogram_pass=time.perf_counter() - t0 - t_cons, consumer=t_cons)
        return dict(seconds=self.seconds, L_total=L_total, ncol_total=info["ncol_total"], ncol_local=ncol, pieces=len(pieces), pmax_raw=pm, nfft=info["nfft"])
