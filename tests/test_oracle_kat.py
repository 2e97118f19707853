"""Analytic known-answer tests that pin the oracle to the reference lines.

The reference ships no tests or golden vectors, so these are derived by hand from RP:203-260.
"""
import numpy as np
import pytest
from scipy.signal import windows

from oracle import fmcw_oracle as O


def _cfg(NTS=128, PN=64, nrx=1):
    sx = O.make_sxml(numSamplesPerChirp=NTS, numChirpsPerFrame=PN, numAntennasRx=nrx)
    return sx, O.configure(sx)


def _tone(NTS, PN, k_r, k_d, A, NR=256, ND=16, dc=0.5 + 0.5j):
    n = np.arange(NTS)[:, None]
    m = np.arange(PN)[None, :]
    return dc + A * np.exp(2j * np.pi * (k_r * n / NR + k_d * m / ND))


def test_derived_parameters():
    sx, cfg = _cfg()
    assert cfg.PRT == pytest.approx(8e-4, rel=1e-12)             # RP:94-97, dashboard screenshot
    assert cfg.Bandwidth == 200e6 and cfg.carrier_frequency == 2.4125e10
    assert cfg.IF_scale == 16 * 3.3 * 256 / 128                  # RP:121
    assert cfg.R_max == 128 * 3e8 / (2 * 200e6)                  # RP:142
    assert cfg.dist_per_bin == cfg.R_max / 256
    assert cfg.fD_per_bin == (1 / (2 * cfg.PRT)) / 16            # RP:152-153 (not 2*fD_max/ND)
    assert cfg.overlap == 19 and cfg.window_length == 20
    assert np.allclose(cfg.range_window_func, 2 * np.blackman(128))


def test_calib_decimation():
    # RP:167-174 with N_cal = 2*NTS, two RX
    sx, cfg = _cfg(NTS=64, PN=16, nrx=2)
    N_cal = 128
    calib = np.arange(2 * 2 * N_cal, dtype=float)
    c = O.calib_rx1(calib, cfg)
    assert c.shape == (64,)
    assert np.array_equal(c.real, calib[0:N_cal:2]) and np.array_equal(c.imag, calib[N_cal:2 * N_cal:2])


@pytest.mark.parametrize("k_r,k_d", [(26, 3), (40, -5), (12, 0)])
def test_single_tone_kat(k_r, k_d):
    """On-bin tone with an integer number of cycles in NTS samples: mean removal leaves it
    untouched; range argmax = k_r+1; Doppler argmax = mod(k_d+ND/2, ND)+1; peak magnitude =
    A*IF_scale*sum(2*blackman) (RP:203-205, 219)."""
    sx, cfg = _cfg(NTS=128, PN=16)
    A = 900 / 4095
    chirp = _tone(128, 16, k_r, k_d, A)[:, :, None]
    cal = np.full(128, 0.5 + 0.5j)
    fo = O.process_frame(chirp, cal, cfg)
    assert list(fo.tgt_range_idx) == [k_r + 1]
    expect = A * cfg.IF_scale * np.sum(2 * windows.blackman(128))
    assert fo.tgt_range_mag[0] == pytest.approx(expect, rel=1e-12)
    if k_d == 0:
        # mean removal over chirps kills a zero-Doppler target: falls back to the DC bin 9
        assert fo.tgt_doppler_idx[0] == 9
    else:
        assert fo.tgt_doppler_idx[0] == (k_d + 8) % 16 + 1
    # slow-time row is the un-mean-removed row (RP:207 stored before RP:218)
    assert np.allclose(np.abs(fo.slow_time_row), expect, rtol=1e-12)


def test_dc_only_no_detection():
    sx, cfg = _cfg(NTS=128, PN=16)
    chirp = np.full((128, 16, 1), 0.5 + 0.5j)
    fo = O.process_frame(chirp, np.full(128, 0.5 + 0.5j), cfg)
    assert fo.tgt_range_idx.size == 0 and fo.slow_time_row is None
    assert np.all(fo.range_max == 0)
    r = O.radar_processing_yes([chirp] * 3, np.full(256, 0.5), sx)
    assert np.isnan(r["range"]).all() and r["batches"] == []          # RP:524-528


def test_doppler_uses_first_16_chirps_only():
    """fft(x, 16, 2) truncates when PN > 16 (RP:219); the mean still uses all PN (RP:217)."""
    sx, cfg = _cfg(NTS=128, PN=64)
    rng = np.random.default_rng(0)
    A = 900 / 4095
    base = _tone(128, 64, 26, 3, A)
    noise = 1e-3 * (rng.standard_normal((128, 64)) + 1j * rng.standard_normal((128, 64)))
    a = base + noise
    b = a.copy()
    # change chirps 17..64 with a zero-mean perturbation over those chirps (keeps RP:217's mean)
    pert = 1e-3 * rng.standard_normal((128, 48))
    pert -= pert.mean(axis=1, keepdims=True)
    b[:, 16:] += pert
    cal = np.full(128, 0.5 + 0.5j)
    fa, fb = O.process_frame(a[:, :, None], cal, cfg), O.process_frame(b[:, :, None], cal, cfg)
    assert list(fa.tgt_range_idx) == list(fb.tgt_range_idx)
    assert np.allclose(fa.doppler_row, fb.doppler_row, rtol=0, atol=1e-9)
    assert not np.allclose(fa.slow_time_row, fb.slow_time_row)


def test_search_peak_shim():
    s = np.zeros(256)
    s[[10, 30, 31, 50, 200]] = [300, 500, 500, 900, 5000]
    d = 0.375
    idx, mag = O.f_search_peak(s, 256, 200, 1, 0.9, 25.0, d, peak_mode="strongest")
    assert list(idx) == [51] and list(mag) == [900]                 # strongest inside the gate (bin 200 is > 25 m)
    idx, mag = O.f_search_peak(s, 256, 200, 1, 0.9, 25.0, d)        # default: the vendor order, nearest peak first
    assert list(idx) == [11] and list(mag) == [300]
    idx, _ = O.f_search_peak(s, 256, 200, 3, 0.9, 25.0, d, peak_mode="strongest")
    assert list(idx) == [51, 32, 11]                                # plateau 30/31: >= left, > right -> index 32 (1-based)
    idx, _ = O.f_search_peak(s, 256, 200, 3, 0.9, 25.0, d)
    assert list(idx) == [11, 32, 51]
    idx, _ = O.f_search_peak(s, 256, 1000, 1, 0.9, 25.0, d)
    assert idx.size == 0
    s2 = np.zeros(256); s2[2] = 900                                 # 0.75 m < min_distance
    assert O.f_search_peak(s2, 256, 200, 1, 0.9, 25.0, d)[0].size == 0


def test_yes_branch_batches_and_break():
    sx, cfg = _cfg(NTS=128, PN=16)
    A = 900 / 4095
    chirp = _tone(128, 16, 26, 3, A)[:, :, None]
    empty = np.full((128, 16, 1), 0.5 + 0.5j)
    frames = [chirp] * 450 + [empty] * 100 + [chirp] * 60
    r = O.radar_processing_yes(frames, np.full(256, 0.5), sx)
    assert [b["batch"] for b in r["batches"]] == [1, 2, 3, 4]         # RP:443, 537
    assert r["batches"][0]["start_frame"] == 1 and r["batches"][0]["end_frame"] == 100
    assert r["batches"][0]["intensity"].shape == (1024, 100 * 16 - 19)
    # batch 5 (frames 401..500, 50 detections) triggers the break at RP:599: frames > 500 stay 0
    assert np.isnan(r["range"][0, 450:500]).all()
    assert np.all(r["range"][0, 500:] == 0)
    assert np.all(r["range"][0, :450] == 26 * cfg.dist_per_bin)
