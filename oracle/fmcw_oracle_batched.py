"""Vectorised, multi-threaded form of the oracle ('no' branch, RP:197-299).

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see fmcw_oracle.py; PARITY UNPINNED for the same reasons).
Same arithmetic as ``fmcw_oracle.radar_processing_no`` but batched over frames with
``scipy.fft(..., workers=N)`` and BLAS threads: this is what ``bench.py --impl reference`` times as "the
reference's CPU implementation with all the host threads it can use".  tests/test_oracle_batched.py checks
it against the serial oracle.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft

from . import fmcw_oracle as O


def search_peak_batched(rmax, cfg):
    """f_search_peak (shim) for every frame at once, max_num_targets = 1.  rmax: [N, NR]."""
    N, NR = rmax.shape
    n1 = np.arange(1, NR + 1)                                 # MATLAB n
    rng = (n1 - 1) * cfg.dist_per_bin
    gate = (n1 >= 3) & (n1 <= NR - 2) & (rng >= cfg.min_distance) & (rng <= cfg.max_distance)
    s = np.pad(rmax, ((0, 0), (2, 2)), constant_values=np.inf)
    c = s[:, 2:-2]
    cand = gate[None, :] & (c >= cfg.range_threshold) & (c >= s[:, 0:-4]) & (c >= s[:, 1:-3]) & (c > s[:, 3:-1]) & (c > s[:, 4:])
    det = cand.any(axis=1)
    if cfg.peak_mode == "strongest":
        masked = np.where(cand, c, -np.inf)
        idx0 = masked.argmax(axis=1)                          # first max = lowest index on ties
    else:
        idx0 = cand.argmax(axis=1)
    return det, idx0


def frame_chain_batched(iq_codes, calib_codes, sxml, workers=1, chunk=512, adc_scale=4095.0, **cfg_kw):
    cfg = O.configure(sxml, **cfg_kw)
    cal = O.calib_rx1(np.asarray(calib_codes, dtype=np.float64) / adc_scale, cfg)
    N = iq_codes.shape[0]
    NR, PN, ND = cfg.range_fft_size, cfg.num_chirps_per_frame, cfg.Doppler_fft_size
    rmax_all = np.empty((N, NR))
    det_all = np.zeros(N, dtype=bool)
    ridx = np.zeros(N, dtype=np.int64)
    rmag = np.zeros(N)
    didx = np.full(N, ND // 2 + 1, dtype=np.int64)
    drows = np.zeros((N, ND), dtype=np.complex128)
    slow = []
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        z = iq_codes[s:e, 0]                                           # RP:202: RX 1 only; [n, PN, NTS, 2]
        x = (z[..., 0].astype(np.float64) + 1j * z[..., 1].astype(np.float64)) / adc_scale
        x = (x - cal[None, None, :]) * cfg.IF_scale                   # RP:203
        x = x - x.mean(axis=2, keepdims=True)                         # RP:204
        X = sfft.fft(x * cfg.range_window_func[None, None, :], NR, axis=2, workers=workers)   # RP:205  [n, PN, NR]
        rmax = np.abs(X).max(axis=1)                                  # RP:210
        det, idx0 = search_peak_batched(rmax, cfg)                    # RP:211
        rmax_all[s:e] = rmax
        det_all[s:e] = det
        rows = X[np.arange(e - s), :, idx0]                           # [n, PN] stored (pre-MTI) rows, RP:207/259
        m = rows - rows.mean(axis=1, keepdims=True)                   # RP:217-218
        rd = sfft.fftshift(sfft.fft(m * cfg.doppler_window_func[None, :], ND, axis=1), axes=1)   # RP:219
        a = np.abs(rd)
        k = a.argmax(axis=1)
        val = a[np.arange(e - s), k]
        dc = ND // 2 + 1
        dd = np.where((val >= cfg.Doppler_threshold) & (k + 1 != dc), k + 1, dc)               # RP:233-238
        ridx[s:e] = np.where(det, idx0 + 1, 0)
        rmag[s:e] = np.where(det, rmax[np.arange(e - s), idx0], 0.0)
        didx[s:e] = np.where(det, dd, dc)
        drows[s:e] = np.where(det[:, None], rd, 0)
        slow.append(rows[det].reshape(-1))
    slow_all = np.concatenate(slow) if slow else np.zeros(0, dtype=np.complex128)
    return dict(cfg=cfg, range_tx1rx1_max_abs=rmax_all.T, detected=det_all, range_idx=ridx, range_mag=rmag,
                doppler_idx=didx, doppler_rows=drows, slow_time_signal_all_frames=slow_all)


def run_no_branch(iq_codes, calib_codes, sxml, L_total=None, workers=1, stft=True, **cfg_kw):
    r = frame_chain_batched(iq_codes, calib_codes, sxml, workers=workers, **cfg_kw)
    cfg = r["cfg"]
    x = np.abs(r["slow_time_signal_all_frames"])                      # RP:270
    if stft and len(x) >= cfg.window_length:
        Lt = max(L_total or 0, len(x))
        ncol = (len(x) - cfg.overlap) // (cfg.window_length - cfg.overlap)
        r["stft"] = O.stft_restated(x, cfg, col_range=(0, ncol), L_total=Lt)
    return r
