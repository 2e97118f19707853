"""Pins the oracle's STFT (RP:270-299) against scipy.signal.spectrogram (an independent
implementation with MATLAB-equivalent arguments) and the restated form against the literal."""
import numpy as np
import pytest
import scipy.signal as ss

from oracle import fmcw_oracle as O


def _cfg(**kw):
    return O.configure(O.make_sxml(), **kw)


@pytest.mark.parametrize("L,win,ov", [(500, 20, 19), (1300, 20, 19), (1000, 32, 16), (2048, 64, 57)])
def test_literal_matches_scipy_spectrogram(L, win, ov):
    cfg = _cfg(window_length=win, overlap=ov)
    x = np.abs(np.random.default_rng(L).standard_normal(L)) * 1000 + 50
    lit = O.stft_literal(x, cfg)
    fs = 1 / cfg.PRT
    f, t, P = ss.spectrogram(x, fs, window=ss.windows.kaiser(win, 3.0, sym=True), nperseg=win, noverlap=ov,
                             nfft=lit["nfft"], detrend=False, return_onesided=True, scaling="density", mode="psd")
    assert np.allclose(P, lit["P"], rtol=1e-11, atol=0)
    assert np.allclose(t, lit["T"], rtol=1e-13)
    assert lit["frequency"][0] == pytest.approx(fs / lit["nfft"], rel=1e-14)
    assert lit["frequency"][-1] == pytest.approx(fs / 2, rel=1e-14)


@pytest.mark.parametrize("L,win,ov", [(64, 20, 19), (500, 20, 19), (1500, 20, 19), (4097, 20, 19),
                                     (1000, 32, 16), (3000, 128, 96), (2500, 256, 230)])
def test_restated_equals_literal(L, win, ov):
    cfg = _cfg(window_length=win, overlap=ov)
    rng = np.random.default_rng(L + win)
    x = 2000 + 30 * rng.standard_normal(L) + 200 * np.sin(np.arange(L) * 0.05)
    x = np.abs(x)
    lit, res = O.stft_literal(x, cfg), O.stft_restated(x, cfg)
    assert res["pmax_raw"] / (1 / cfg.PRT * np.sum(ss.windows.kaiser(win, 3.0) ** 2)) == pytest.approx(lit["pmax"], rel=1e-12)
    assert np.abs(res["intensity"] - lit["intensity"]).max() < 1e-8
    assert np.array_equal(res["frequency"], lit["frequency"]) and np.array_equal(res["T"], lit["T"])


def test_global_max_not_at_bin_one():
    """H2: sparse non-negative sequences at small nfft put the PSD max far from bins 0/1; the
    exact bounded search of the restated path must still find it."""
    cfg = _cfg()
    rng = np.random.default_rng(0)
    hits = 0
    for _ in range(60):
        L = int(rng.integers(24, 200))
        x = rng.random(L) * (rng.random(L) < 0.15) * 1000 + 1e-3 * rng.random(L)
        lit, res = O.stft_literal(x, cfg), O.stft_restated(x, cfg)
        hits += int(np.argmax(lit["P"].max(axis=1))) > 1
        scale = 1 / cfg.PRT * np.sum(ss.windows.kaiser(20, 3.0) ** 2)
        assert res["pmax_raw"] / scale == pytest.approx(lit["pmax"], rel=1e-12)
        fin = np.isfinite(lit["intensity"])
        assert np.abs(res["intensity"][fin] - lit["intensity"][fin]).max() < 1e-6
    assert hits >= 3


def test_constant_sequence_kat():
    """STFT of a constant: every column identical; the one-sided doubling puts the max at bin 1,
    so the interpolated value at the first log-frequency (= bin 1) is 0 dB (RP:276-283, 294)."""
    cfg = _cfg()
    lit = O.stft_literal(np.full(400, 7.0), cfg)
    assert np.allclose(lit["intensity"], lit["intensity"][:, :1])
    assert int(np.argmax(lit["P"][:, 0])) == 1
    assert lit["intensity"][0, 0] == pytest.approx(0.0, abs=1e-9)
    w = ss.windows.kaiser(20, 3.0)
    S0 = w.sum() * 7.0
    S1 = abs(np.sum(w * 7.0 * np.exp(-2j * np.pi * np.arange(20) / lit["nfft"])))
    dc_db = 20 * np.log10(S0 ** 2 / (2 * S1 ** 2))
    assert dc_db == pytest.approx(-6.02, abs=0.1)          # -> -6.02 as nfft grows
    assert 20 * np.log10(lit["P"][0, 0] / lit["pmax"]) == pytest.approx(dc_db, abs=1e-9)


def test_column_range_restriction():
    cfg = _cfg()
    x = np.abs(1500 + 40 * np.random.default_rng(5).standard_normal(3000))
    full = O.stft_restated(x, cfg)
    part = O.stft_restated(x, cfg, pmax_raw=full["pmax_raw"], col_range=(100, 900))
    assert np.array_equal(part["intensity"], full["intensity"][:, 100:900])


def test_literal_in_column_blocks_is_bit_identical():
    """The C1-size literal check of the GPU suite walks P in blocks of columns; same bits as the one-shot literal form."""
    sx = O.make_sxml(numSamplesPerChirp=64, numChirpsPerFrame=16)
    cfg = O.configure(sx)
    rng = np.random.default_rng(5)
    x = np.abs(2400 + 30 * rng.standard_normal(700) + 200 * np.sin(np.arange(700) * 0.05))
    a = O.stft_literal(x, cfg)
    b = O.stft_literal_chunked(x, cfg, col_chunk=97)
    assert np.array_equal(a["intensity"], b["intensity"]) and a["pmax"] == b["pmax"]
