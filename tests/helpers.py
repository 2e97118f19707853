"""Shared scenario builders for the parity tests (oracle side + library side)."""
from __future__ import annotations

import numpy as np

from fmcw_radar_processing_b200 import synth
from fmcw_radar_processing_b200.config import fmcw_configurations
from oracle import fmcw_oracle as O


def make_case(n_frames=40, NTS=128, PN=64, n_rx=1, scene=None, seed=1, frame0=0, window_length=20, overlap=None,
              rx_select=1, peak_mode="first", sigma=2.0):
    sx = O.make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=n_rx)
    ocfg = O.configure(sx, window_length=window_length, overlap=overlap, peak_mode=peak_mode)
    cfg = fmcw_configurations(sx, window_length=window_length, overlap=overlap, rx_select=rx_select, peak_mode=peak_mode)
    sc = scene or synth.scene_c1(seed)
    sc.sigma = sigma
    tab = synth.scene_tables(sc, ocfg.dist_per_bin, 256, ocfg.PRT, ocfg.lambda_, frame0, n_frames)
    iq = synth.synth_frames(tab, sc.seed, frame0, n_rx, PN, NTS, sigma=sc.sigma, dc=sc.dc, rx_step=sc.rx_step)
    calib = synth.default_calib(n_rx, NTS)
    return dict(sxml=sx, ocfg=ocfg, cfg=cfg, iq=iq, calib_codes=calib, calib=calib / 4095.0, tables=tab, scene=sc)


def oracle_no(case, stft="restated", rx_select=1):
    frames, n, calib, sx = O.f_parse_data2(case["iq"], case["calib_codes"], case["sxml"])
    if rx_select != 1:   # the reference always reads RX 1 (RP:202); emulate another RX by rotating it to the front
        frames = [f[:, :, rx_select - 1:rx_select] for f in frames]
        ncal = len(calib) // (2 * case["ocfg"].num_Rx_antennas)
        calib = np.concatenate([calib[2 * (rx_select - 1) * ncal:2 * rx_select * ncal],
                                calib[:2 * (rx_select - 1) * ncal], calib[2 * rx_select * ncal:]])
    kw = dict(window_length=case["ocfg"].window_length, overlap=case["ocfg"].overlap, peak_mode=case["ocfg"].peak_mode)
    return O.radar_processing_no(frames, calib, sx, stft=stft, **kw)


def db_errors(gpu_lin, ref_lin):
    """north_star tolerance metrics for linear magnitudes: (max |dB err| above peak-60 dB,
    max relative err elsewhere)."""
    gpu = np.asarray(gpu_lin, dtype=np.float64)
    ref = np.asarray(ref_lin, dtype=np.float64)
    peak = ref.max()
    with np.errstate(divide="ignore", invalid="ignore"):
        level = 20 * np.log10(ref / peak)
        strong = level > -60
        e_db = np.abs(20 * np.log10(gpu[strong] / ref[strong])).max() if strong.any() else 0.0
        weak = ~strong & (ref > 0)
        e_rel = (np.abs(gpu[weak] - ref[weak]) / ref[weak]).max() if weak.any() else 0.0
    return float(e_db), float(e_rel)


def assert_same_finiteness(g, r):
    """A NaN / Inf the GPU produces where the reference is finite (or the reverse) is an error, never masked."""
    bad = np.isfinite(g) != np.isfinite(r)
    assert not bad.any(), f"{int(bad.sum())} bins are finite on one side only (first at {np.argwhere(bad)[0].tolist()})"


# Spectrogram tolerance contract (stated once; include/fmcw_cuda.h and DESIGN.md section 4 quote it):
#   * both modes: 1e-3 dB for bins above -60 dB; 1e-4 relative (of P/max) for bins in (-140, -60] dB;
#   * FMCW_OPT_STFT_PRECISION = 0 (default, TF32 x 2 tensor-core kernel): below -140 dB the operand split (2^-22 of the
#     mean-removed column) is the floor: 1e-3 relative in (-180, -140] dB, unspecified below;
#   * FMCW_OPT_STFT_PRECISION = 1 (float64 kernel): 1e-4 relative at every finite level.
SPEC_TOL_DB, SPEC_TOL_REL, SPEC_FAST_FLOOR_DB, SPEC_FAST_DEEP_REL = 1e-3, 1e-4, -140.0, 1e-3


def assert_spectrogram_contract(gpu_db, ref_db, precise=False):
    e_db, _ = spectrogram_errors(gpu_db, ref_db)
    assert e_db < SPEC_TOL_DB, f"{e_db} dB above -60 dB"
    if precise:
        e = spectrogram_band_rel(gpu_db, ref_db, -np.inf, -60)
        assert e < SPEC_TOL_REL, f"float64 mode: {e} relative below -60 dB"
    else:
        e = spectrogram_band_rel(gpu_db, ref_db, SPEC_FAST_FLOOR_DB, -60)
        assert e < SPEC_TOL_REL, f"{e} relative in ({SPEC_FAST_FLOOR_DB}, -60] dB"
        e = spectrogram_band_rel(gpu_db, ref_db, -180, SPEC_FAST_FLOOR_DB)
        assert e < SPEC_FAST_DEEP_REL, f"{e} relative in (-180, {SPEC_FAST_FLOOR_DB}] dB"
    return e_db


def spectrogram_band_rel(gpu_db, ref_db, lo, hi):
    """Max relative power error of the spectrogram bins whose reference level lies in (lo, hi] dB."""
    g = np.asarray(gpu_db, dtype=np.float64)
    r = np.asarray(ref_db, dtype=np.float64)
    assert_same_finiteness(g, r)
    m = np.isfinite(r) & (r > lo) & (r <= hi)
    return float(np.abs(10 ** ((g[m] - r[m]) / 20) - 1).max()) if m.any() else 0.0


def spectrogram_errors(gpu_db, ref_db):
    """Spectrogram intensities are already 20*log10(P/max(P)) (RP:283).  Returns (max |err| dB where
    ref > -60 dB, max relative power error elsewhere)."""
    g = np.asarray(gpu_db, dtype=np.float64)
    r = np.asarray(ref_db, dtype=np.float64)
    assert_same_finiteness(g, r)
    fin = np.isfinite(r)
    strong = fin & (r > -60)
    weak = fin & ~strong
    e_db = np.abs(g[strong] - r[strong]).max() if strong.any() else 0.0
    e_rel = np.abs(10 ** ((g[weak] - r[weak]) / 20) - 1).max() if weak.any() else 0.0
    return float(e_db), float(e_rel)
