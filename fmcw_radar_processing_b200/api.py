"""Python handle over the C ABI (what the MEX gateway does for MATLAB).

Accepts NumPy arrays (host buffers; the library stages them through the device and copies the
results back) or torch CUDA tensors (device buffers; zero copies, asynchronous on the handle's
stream).  torch is used for device memory only."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .config import to_c_config


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "buffers must be C-contiguous"
        return a.ctypes.data
    return a.data_ptr()          # torch tensor


def _is_torch(a):
    return a is not None and not isinstance(a, np.ndarray)


class FmcwCuda:
    """One libfmcw_cuda handle (one GPU, one stream)."""

    def __init__(self, cfg, calib_data=None, device: int = 0, torch_stream_sync: bool = True):
        """``torch_stream_sync``: when torch CUDA tensors are passed, order the library's stream after torch's
        current stream before each call and torch's current stream after the library's afterwards (stream-ordered
        semantics; also what makes recycled caching-allocator blocks safe).  Callers that manage streams and
        buffer lifetimes themselves (bench.py) switch it off."""
        self.lib = _lib.load()
        self.torch_stream_sync = torch_stream_sync
        self._ext_stream = None
        self.cfg = cfg
        self.c_cfg = to_c_config(cfg)
        self.device = device
        self.NTS = cfg["num_ADC_samples_per_chirp"]
        self.PN = cfg["num_chirps_per_frame"]
        self.n_rx = cfg["num_Rx_antennas"]
        self.NR = cfg["range_fft_size"]
        self.ND = cfg["Doppler_fft_size"]
        self.nq = cfg["MAX_FREQ_BINS"]
        self.hop = cfg["window_length"] - cfg["overlap"]
        self._h = C.c_void_p()
        cal = None if calib_data is None else np.ascontiguousarray(calib_data, dtype=np.float64)
        st = self.lib.fmcw_create(C.byref(self.c_cfg), None if cal is None else cal.ctypes.data,
                                  0 if cal is None else cal.size, device, C.byref(self._h))
        if st != _lib.FMCW_OK:
            raise _lib.FmcwError(st, "fmcw_create failed (is a CUDA device visible? there is no CPU fallback)")

    # ---- lifetime ----
    def close(self):
        if self._h:
            self.lib.fmcw_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != _lib.FMCW_OK:
            raise _lib.FmcwError(st, self.lib.fmcw_last_error(self._h).decode())

    @property
    def stream(self) -> int:
        return self.lib.fmcw_get_stream(self._h) or 0

    def synchronize(self):
        self._check(self.lib.fmcw_synchronize(self._h))

    def set_option(self, option: int, value: int):
        self._check(self.lib.fmcw_set_option(self._h, option, value))

    def info(self) -> dict:
        inf = _lib.fmcw_run_info()
        self._check(self.lib.fmcw_get_info(self._h, C.byref(inf)))
        return {n: getattr(inf, n) for n, _ in _lib.fmcw_run_info._fields_}

    def set_info_target(self, device_ptr: int | None):
        """Every later ``run`` ends by writing an ``fmcw_device_info`` (the ``info()`` scalars + a status word) to this
        device address, stream-ordered; ``None`` disables.  ``parse_device_infos`` decodes a host copy."""
        self._check(self.lib.fmcw_set_info_target(self._h, C.c_void_p(device_ptr or 0)))

    @staticmethod
    def parse_device_infos(raw: np.ndarray) -> list:
        """``raw``: uint8 host array holding consecutive ``fmcw_device_info`` structs -> list of (info dict, status)."""
        sz = C.sizeof(_lib.fmcw_device_info)
        out = []
        buf = np.ascontiguousarray(raw, dtype=np.uint8).tobytes()
        for i in range(len(buf) // sz):
            d = _lib.fmcw_device_info.from_buffer_copy(buf, i * sz)
            out.append(({n: getattr(d.info, n) for n, _ in _lib.fmcw_run_info._fields_}, int(d.status)))
        return out

    def timings(self) -> dict:
        """Device ms of the stages of the last run (CUDA events on the handle's stream)."""
        ms = (C.c_float * 4)()
        self._check(self.lib.fmcw_get_timings(self._h, C.byref(ms)))
        return dict(chain_ms=ms[0], compact_ms=ms[1], plan_max_ms=ms[2], stft_main_ms=ms[3])

    # ---- buffers ----
    def alloc_frame_out(self, n_frames: int, device=None) -> dict:
        """Per-frame output buffers: NumPy (host) by default, torch CUDA tensors if ``device`` is given."""
        shapes = dict(range_max_abs=((n_frames, self.NR), np.float32), detected=((n_frames,), np.int32),
                      range_bin=((n_frames,), np.int32), range_mag=((n_frames,), np.float32),
                      doppler_bin=((n_frames,), np.int32), doppler_row=((n_frames, self.ND, 2), np.float32),
                      slow_time_mag=((n_frames, self.PN), np.float32))
        if device is None:
            return {k: np.zeros(s, dtype=d) for k, (s, d) in shapes.items()}
        import torch
        tmap = {np.float32: torch.float32, np.int32: torch.int32}
        return {k: torch.zeros(s, dtype=tmap[d], device=device) for k, (s, d) in shapes.items()}

    def max_cols(self, n_frames: int) -> int:
        L = n_frames * self.PN
        return max(0, (L - self.cfg["overlap"]) // self.hop) if L >= self.cfg["window_length"] else 0

    def _frame_struct(self, out):
        fo = _lib.fmcw_frame_out()
        if out:
            for name, _ in _lib.fmcw_frame_out._fields_:
                setattr(fo, name, _ptr(out.get(name)))
        return fo

    @staticmethod
    def _stft_struct(intensity, layout):
        so = _lib.fmcw_stft_out()
        so.intensity = _ptr(intensity)
        if layout == _lib.LAYOUT_TIME_MAJOR:
            so.capacity_cols = intensity.shape[0]
            so.ld_cols = 0
        else:
            so.capacity_cols = intensity.shape[1]
            so.ld_cols = intensity.shape[1]
        so.layout = layout
        return so

    # ---- stream ordering against torch ----
    def _order_before(self, ref):
        if self.torch_stream_sync and _is_torch(ref) and ref.is_cuda:
            import torch
            if self._ext_stream is None:
                self._ext_stream = torch.cuda.ExternalStream(self.stream, device=ref.device)
            self._ext_stream.wait_stream(torch.cuda.current_stream(ref.device))
            return ref.device
        return None

    def _order_after(self, dev):
        if dev is not None:
            import torch
            torch.cuda.current_stream(dev).wait_stream(self._ext_stream)

    # ---- the chain ----
    def process_frames(self, iq, out: dict | None = None) -> dict:
        """RP:197-261 for every frame of ``iq`` (int16 [n][rx][PN][NTS][2])."""
        n = int(iq.shape[0])
        if out is None:
            out = self.alloc_frame_out(n, device=iq.device if _is_torch(iq) else None)
        fo = self._frame_struct(out)
        dev = self._order_before(iq)
        self._check(self.lib.fmcw_process_frames(self._h, _ptr(iq), n, C.byref(fo)))
        self._order_after(dev)
        return out

    def run(self, iq, out: dict | None = None, intensity=None, layout: int = _lib.LAYOUT_TIME_MAJOR):
        """Fused frames -> STFT.  Returns (frame outputs, intensity buffer); sizes via ``info()``."""
        n = int(iq.shape[0])
        dev = iq.device if _is_torch(iq) else None
        if out is None:
            out = self.alloc_frame_out(n, device=dev)
        if intensity is None:
            cols = max(1, self.max_cols(n))
            shape = (cols, self.nq) if layout == _lib.LAYOUT_TIME_MAJOR else (self.nq, cols)
            if dev is None:
                intensity = np.empty(shape, dtype=np.float32)
            else:
                import torch
                intensity = torch.empty(shape, dtype=torch.float32, device=dev)
        fo = self._frame_struct(out)
        so = self._stft_struct(intensity, layout)
        sdev = self._order_before(iq)
        self._check(self.lib.fmcw_run(self._h, _ptr(iq), n, C.byref(fo), C.byref(so)))
        self._order_after(sdev)
        return out, intensity

    def stft_frames(self, n_frames: int, intensity=None, layout: int = _lib.LAYOUT_TIME_MAJOR):
        """RP:270-299 / RP:538-566 on the slow-time signal of the last ``process_frames`` call (float64 inside)."""
        if intensity is None:
            cols = max(1, self.max_cols(n_frames))
            shape = (cols, self.nq) if layout == _lib.LAYOUT_TIME_MAJOR else (self.nq, cols)
            intensity = np.empty(shape, dtype=np.float32)
        so = self._stft_struct(intensity, layout)
        self._check(self.lib.fmcw_stft_frames(self._h, C.byref(so)))
        return intensity

    def stft(self, x, intensity=None, layout: int = _lib.LAYOUT_TIME_MAJOR):
        """RP:270-299 on an arbitrary non-negative float32 sequence."""
        L = int(x.shape[0])
        cols = max(1, (L - self.cfg["overlap"]) // self.hop)
        if intensity is None:
            shape = (cols, self.nq) if layout == _lib.LAYOUT_TIME_MAJOR else (self.nq, cols)
            if _is_torch(x):
                import torch
                intensity = torch.empty(shape, dtype=torch.float32, device=x.device)
            else:
                intensity = np.empty(shape, dtype=np.float32)
        so = self._stft_struct(intensity, layout)
        sdev = self._order_before(x)
        self._check(self.lib.fmcw_stft(self._h, _ptr(x), L, C.byref(so)))
        self._order_after(sdev)
        return intensity

    def stft_axes(self, L_total: int, col_begin: int = 0, ncol: int | None = None):
        """float64 T (RP:276) and log_freq_bins (RP:293-296) plus nfft (RP:273)."""
        nfft, nct = C.c_uint64(), C.c_uint64()
        self._check_cfg(self.lib.fmcw_stft_axes(C.byref(self.c_cfg), L_total, 0, 0, None, None, C.byref(nfft), C.byref(nct)))
        if ncol is None:
            ncol = nct.value - col_begin
        T = np.empty(ncol, dtype=np.float64)
        F = np.empty(self.nq, dtype=np.float64)
        self._check_cfg(self.lib.fmcw_stft_axes(C.byref(self.c_cfg), L_total, col_begin, ncol, T.ctypes.data, F.ctypes.data,
                                                C.byref(nfft), C.byref(nct)))
        return T, F, nfft.value, nct.value

    @staticmethod
    def _check_cfg(st):
        if st != _lib.FMCW_OK:
            raise _lib.FmcwError(st, "fmcw_stft_axes")

    # ---- sharded path ----
    def get_slow_time(self, dst, first: int, count: int):
        self._check(self.lib.fmcw_get_slow_time(self._h, _ptr(dst), first, count))
        return dst

    def load_slow_time(self, x, L_local: int, n_halo: int):
        """Streaming: hand a piece of the caller-kept slow-time signal back (L_local samples + n_halo following ones)."""
        self._check(self.lib.fmcw_load_slow_time(self._h, _ptr(x) if (L_local + n_halo) else None, L_local, n_halo))

    def set_halo(self, src, count: int):
        self._check(self.lib.fmcw_set_halo(self._h, _ptr(src) if count else None, count))

    def stft_local_max(self, L_total: int, sample_offset: int) -> float:
        v = C.c_double()
        self._check(self.lib.fmcw_stft_local_max(self._h, L_total, sample_offset, C.byref(v)))
        return v.value

    def stft_sharded(self, L_total: int, sample_offset: int, pmax_raw: float, intensity, layout: int = _lib.LAYOUT_TIME_MAJOR):
        so = self._stft_struct(intensity, layout)
        self._check(self.lib.fmcw_stft_sharded(self._h, L_total, sample_offset, pmax_raw, C.byref(so)))
        return intensity

    # ---- asynchronous sharded path (device-side hand-offs) ----
    def shard_pack(self, msg):
        dev = self._order_before(msg)
        self._check(self.lib.fmcw_shard_pack(self._h, _ptr(msg)))
        self._order_after(dev)

    def shard_plan(self, gathered, world: int, rank: int, local_max):
        dev = self._order_before(gathered)
        self._check(self.lib.fmcw_shard_plan(self._h, _ptr(gathered), world, rank, _ptr(local_max)))
        self._order_after(dev)

    def shard_stft(self, global_max, intensity, layout: int = _lib.LAYOUT_TIME_MAJOR):
        so = self._stft_struct(intensity, layout)
        dev = self._order_before(global_max)
        self._check(self.lib.fmcw_shard_stft(self._h, _ptr(global_max), C.byref(so)))
        self._order_after(dev)
        return intensity

    # ---- sharded path over peer-memory mailboxes (NVLink stores + flags, no collective between the steps) ----
    def mailbox_bytes(self) -> int:
        return int(self.lib.fmcw_mailbox_bytes())

    @staticmethod
    def _mailbox_array(ptrs):
        return (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])

    def mailbox_post_heads(self, ptrs, rank: int, step: int):
        """``ptrs[r]``: device address of rank r's mailbox as mapped in this process."""
        self._check(self.lib.fmcw_mailbox_post_heads(self._h, self._mailbox_array(ptrs), len(ptrs), rank, step))

    def mailbox_plan(self, ptrs, rank: int, step: int):
        self._check(self.lib.fmcw_mailbox_plan(self._h, self._mailbox_array(ptrs), len(ptrs), rank, step))

    def mailbox_stft(self, ptrs, rank: int, step: int, intensity, layout: int = _lib.LAYOUT_TIME_MAJOR):
        so = self._stft_struct(intensity, layout)
        dev = self._order_before(intensity)
        self._check(self.lib.fmcw_mailbox_stft(self._h, self._mailbox_array(ptrs), len(ptrs), rank, step, C.byref(so)))
        self._order_after(dev)
        return intensity

    # ---- extras ----
    def stft_finegrid(self, f_lo: float = 0.0, f_hi: float = 150.0, max_rows: int = 2048, out=None):
        """psd band of the last STFT on the fine grid (RP:283, the matrix surf draws at RP:333 between ylim [0 150]).
        Returns (psd [ncol][n_rows] float32, F [n_rows] float64 Hz)."""
        info = self.info()
        ncl = info["ncol_local"]
        if out is None:
            out = np.empty((max(1, ncl), max_rows), dtype=np.float32)
        fb, st, nr, nc = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.fmcw_stft_finegrid(self._h, f_lo, f_hi, max_rows, _ptr(out), out.shape[0], C.byref(fb), C.byref(st),
                                                C.byref(nr), C.byref(nc)))
        F = (fb.value + np.arange(nr.value) * st.value) * (1.0 / self.cfg["PRT"]) / info["nfft"]
        flat = out.reshape(-1)[:nc.value * nr.value] if isinstance(out, np.ndarray) else out.view(-1)[:nc.value * nr.value]
        return flat.reshape(nc.value, nr.value), F

    def range_doppler_map(self, iq, frame: int) -> np.ndarray:
        """Full range-Doppler map of one frame in dB, [range_fft_size][Doppler_fft_size] (RP:216-219 on every range row)."""
        out = np.empty((self.NR, self.ND), dtype=np.float32)
        self._check(self.lib.fmcw_range_doppler_map(self._h, _ptr(iq), int(iq.shape[0]), frame, out.ctypes.data))
        return out

    def range_spectrum(self, iq, frame: int, chirp: int) -> np.ndarray:
        """abs(range_fft(:, chirp)) of one frame (RP:410-411), 0-based indices."""
        out = np.empty(self.NR, dtype=np.float32)
        self._check(self.lib.fmcw_range_spectrum(self._h, _ptr(iq), int(iq.shape[0]), frame, chirp, out.ctypes.data))
        return out

    def synth_frames(self, tables: np.ndarray, seed: int, frame0: int, sigma=2.0, dc=2048.0, rx_step=0.11, out=None):
        n, n_scat, _ = tables.shape
        tables = np.ascontiguousarray(tables, dtype=np.float64)
        if out is None:
            out = np.empty((n, self.n_rx, self.PN, self.NTS, 2), dtype=np.int16)
        self._check(self.lib.fmcw_synth_frames(self._h, tables.ctypes.data, n_scat, seed, frame0, n, sigma, dc, rx_step, _ptr(out)))
        return out
