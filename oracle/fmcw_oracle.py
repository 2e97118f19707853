"""CPU oracle for the range-Doppler-STFT chain of alepnabil/fmcw_radar_processing.

TEST INFRASTRUCTURE ONLY.  Nothing in ``fmcw_radar_processing_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or the
CPU baseline, never as the product path.

PARITY UNPINNED.  The reference is MATLAB.  Neither MATLAB nor Octave exists in the
build or GPU images, the reference ships no tests / golden vectors / recordings, and
three functions on its hot path are not in the repository at all (``f_parse_data2``,
``f_search_peak``, ``xml2struct``; call sites RP:81-86, RP:211, RP:469).  This file is
therefore a float64 NumPy/SciPy restatement of the reference *lines*, anchored on

* the analytic known-answer tests derived from those lines (tests/test_oracle_kat.py),
* ``scipy.signal.spectrogram`` with MATLAB-equivalent arguments as an independent
  implementation of the ``spectrogram`` call at RP:276 (tests/test_oracle_stft.py),
* equality of the *literal* full-``nfft`` STFT path and the *restated* sampled-DTFT
  path (same test file),

and NOT on outputs of the reference itself.

Citations: ``RP:n`` = /root/reference/radar-etl-pipeline/radar_processing.m line n,
``RPA:n`` = radar_processing_with_azure.m line n.

Third-party arithmetic the reference calls and this file restates with SciPy:
MathWorks Signal Processing Toolbox ``blackman`` / ``chebwin`` / ``kaiser`` /
``spectrogram`` and base ``fft`` / ``fftshift`` / ``interp1`` / ``logspace`` /
``nextpow2`` (proprietary, version unpinned; call sites RP:138, 139, 205, 219, 273,
276, 279-280, 296, 299).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.fft as sfft
from scipy.signal import windows as _win

C0 = 3e8  # RP:67


# --------------------------------------------------------------------------------------
# shims for the functions the reference calls but does not ship (our definitions)
# --------------------------------------------------------------------------------------
def make_sxml(chirpDuration_ns=300000, lowerFrequency_kHz=24025000, upperFrequency_kHz=24225000,
              numAntennasTx=1, numAntennasRx=1, numSamplesPerChirp=128, numChirpsPerFrame=64,
              samplerateHz=426666):
    """xml2struct-shaped nested dict with the ``.Text`` leaves RP:94-115 reads."""
    t = lambda v: {"Text": repr(v) if not isinstance(v, str) else v}
    return {"Device": {
        "BaseEndpoint": {
            "chirpDuration_ns": t(chirpDuration_ns),
            "DeviceInfo": {"numAntennasTx": t(numAntennasTx), "numAntennasRx": t(numAntennasRx)},
            "FrameFormat": {"numSamplesPerChirp": t(numSamplesPerChirp),
                            "numChirpsPerFrame": t(numChirpsPerFrame)},
        },
        "FmcwEndpoint": {"FmcwConfiguration": {"upperFrequency_kHz": t(upperFrequency_kHz),
                                               "lowerFrequency_kHz": t(lowerFrequency_kHz)}},
        "AdcxmcEndpoint": {"AdcxmcConfiguration": {"samplerateHz": t(samplerateHz)}},
    }}


def f_parse_data2(iq_codes, calib_codes, sxml, adc_scale=4095.0):
    """Shim of the unshipped parser called at RP:86.

    ``iq_codes``: int16 [frame][rx][chirp][sample][2] ADC codes 0..4095 (the new
    library's input format).  Returns what RP:86 expects: ``frame`` (list of
    [NTS x PN x RX] complex128 arrays, ``frame(k).Chirp``), ``frame_count``,
    ``calib_data`` (row vector ``[I_rx1 Q_rx1 I_rx2 Q_rx2 ...]``, RP:167-172) and
    ``sXML``.  Normalisation by ``adc_scale`` is our definition (parity unpinned).
    """
    iq = np.asarray(iq_codes)
    n_frames = iq.shape[0]
    z = (iq[..., 0].astype(np.float64) + 1j * iq[..., 1].astype(np.float64)) / adc_scale
    # [frame][rx][chirp][sample] -> per frame [sample, chirp, rx]
    frames = [np.ascontiguousarray(z[f].transpose(2, 1, 0)) for f in range(n_frames)]
    calib = np.asarray(calib_codes, dtype=np.float64).reshape(-1) / adc_scale
    return frames, n_frames, calib, sxml


def f_search_peak(sig, length, threshold, max_num, min_distance, max_distance, dist_per_bin,
                  peak_mode="first"):
    """Shim of the unshipped range peak picker called at RP:211 / RP:469.

    5-point local maximum: ``s[n] >= s[n-1], s[n-2]`` and ``s[n] > s[n+1], s[n+2]`` for
    n = 3..length-2 (1-based), ``s[n] >= threshold`` and ``(n-1)*dist_per_bin`` inside
    ``[min_distance, max_distance]``.  ``peak_mode='first'`` (default) keeps the first ``max_num``
    candidates in increasing range: the vendor SDK demo this function comes from walks n upward and
    stops appending once ``max_num`` peaks are found, so with ``max_num_targets = 1`` (RP:129) the
    NEAREST peak above the threshold is tracked.  ``'strongest'`` keeps the ``max_num`` largest
    candidates (ties: lowest index first), what RP:258's comment "Index of strongest target" reads
    like.  The two differ only when a frame holds several peaks above ``range_threshold`` inside
    the distance gate (tests/test_gpu_parity.py::test_two_targets_peak_modes).  Returns 1-based
    indices and magnitudes (row vectors, empty when none).
    """
    s = np.asarray(sig, dtype=np.float64).reshape(-1)
    cand = []
    for n in range(3, length - 2 + 1):        # MATLAB n = 3 .. length-2
        i = n - 1
        fp = s[i]
        rng = (n - 1) * dist_per_bin
        if rng < min_distance or rng > max_distance:
            continue
        if fp >= threshold and fp >= s[i - 2] and fp >= s[i - 1] and fp > s[i + 1] and fp > s[i + 2]:
            cand.append(n)
    if peak_mode == "strongest":
        cand.sort(key=lambda n: (-s[n - 1], n))
    elif peak_mode != "first":
        raise ValueError(peak_mode)
    idx = np.array(cand[:max_num], dtype=np.int64)
    return idx, s[idx - 1] if idx.size else np.zeros(0)


# --------------------------------------------------------------------------------------
# configuration  (RP:89-179)
# --------------------------------------------------------------------------------------
def _txt(node):
    return float(node["Text"])  # str2double(... .Text)


@dataclass
class Config:
    """The locals of RP:89-179 under the names of the commented-out
    ``fmcw_configurations`` struct (RP:645-672), plus the literals."""
    frame_time: float
    PRT: float
    Bandwidth: float
    num_Tx_antennas: int
    num_Rx_antennas: int
    carrier_frequency: float
    num_ADC_samples_per_chirp: int
    num_chirps_per_frame: int
    sampling_frequency: float
    range_fft_size: int
    Doppler_fft_size: int
    IF_scale: float
    range_threshold: float
    Doppler_threshold: float
    min_distance: float
    max_distance: float
    max_num_targets: int
    lambda_: float
    Hz_to_mps_constant: float
    R_max: float
    dist_per_bin: float
    fD_max: float
    fD_per_bin: float
    window_length: int
    overlap: int
    # literals of the reference
    kaiser_beta: float = 3.0          # RP:276
    MAX_FREQ_BINS: int = 1024         # RP:293
    batch_size: int = 100             # RP:189
    max_plots: int = 4                # RP:443
    peak_mode: str = "first"          # shim choice, see f_search_peak
    range_window_func: np.ndarray = field(default=None, repr=False)
    doppler_window_func: np.ndarray = field(default=None, repr=False)
    array_bin_range: np.ndarray = field(default=None, repr=False)


def configure(sxml, window_length=20, overlap=None, range_fft_size=256, Doppler_fft_size=16,
              peak_mode="first") -> Config:
    """RP:89-154 and RP:178-179, line by line."""
    frame_time = 150 * 1e-3                                                    # RP:91
    up = _txt(sxml["Device"]["BaseEndpoint"]["chirpDuration_ns"]) * 1e-9      # RP:94
    PRT = up + 200e-6 + 300e-6                                                 # RP:95-97
    fm = sxml["Device"]["FmcwEndpoint"]["FmcwConfiguration"]
    BW = (_txt(fm["upperFrequency_kHz"]) - _txt(fm["lowerFrequency_kHz"])) * 1e3        # RP:100
    ntx = int(_txt(sxml["Device"]["BaseEndpoint"]["DeviceInfo"]["numAntennasTx"]))     # RP:102
    nrx = int(_txt(sxml["Device"]["BaseEndpoint"]["DeviceInfo"]["numAntennasRx"]))     # RP:103
    fC = (_txt(fm["upperFrequency_kHz"]) + _txt(fm["lowerFrequency_kHz"])) / 2 * 1e3    # RP:106
    NTS = int(_txt(sxml["Device"]["BaseEndpoint"]["FrameFormat"]["numSamplesPerChirp"]))    # RP:109
    PN = int(_txt(sxml["Device"]["BaseEndpoint"]["FrameFormat"]["numChirpsPerFrame"]))      # RP:112
    fS = _txt(sxml["Device"]["AdcxmcEndpoint"]["AdcxmcConfiguration"]["samplerateHz"])      # RP:115
    IF_scale = 16 * 3.3 * range_fft_size / NTS                                 # RP:121,136
    lam = C0 / fC                                                              # RP:133
    R_max = NTS * C0 / (2 * BW)                                                # RP:142
    dist_per_bin = R_max / range_fft_size                                      # RP:147
    fD_max = 1 / (2 * PRT)                                                     # RP:152
    fD_per_bin = fD_max / Doppler_fft_size                                     # RP:153
    if overlap is None:
        overlap = window_length - 1                                            # RP:179
    return Config(
        frame_time=frame_time, PRT=PRT, Bandwidth=BW, num_Tx_antennas=ntx, num_Rx_antennas=nrx,
        carrier_frequency=fC, num_ADC_samples_per_chirp=NTS, num_chirps_per_frame=PN,
        sampling_frequency=fS, range_fft_size=range_fft_size, Doppler_fft_size=Doppler_fft_size,
        IF_scale=IF_scale, range_threshold=200.0, Doppler_threshold=50.0,      # RP:123-124
        min_distance=0.9, max_distance=25.0, max_num_targets=1,                # RP:126-129
        lambda_=lam, Hz_to_mps_constant=lam / 2,                               # RP:135
        R_max=R_max, dist_per_bin=dist_per_bin, fD_max=fD_max, fD_per_bin=fD_per_bin,
        window_length=window_length, overlap=overlap, peak_mode=peak_mode,
        range_window_func=2 * _win.blackman(NTS, sym=True),                    # RP:138
        doppler_window_func=2 * _win.chebwin(PN, at=100, sym=True) if PN > 1 else np.array([2.0]),  # RP:139
        array_bin_range=np.arange(range_fft_size) * dist_per_bin,              # RP:149
    )


def calib_rx1(calib_data, cfg: Config):
    """RP:167-174: decimate the calibration row vector to NTS complex samples of RX1."""
    calib_data = np.asarray(calib_data, dtype=np.float64).reshape(-1)
    N_cal = len(calib_data) // (2 * cfg.num_Rx_antennas)                       # RP:167
    dec = N_cal // cfg.num_ADC_samples_per_chirp                               # RP:169
    ci = calib_data[0:N_cal:dec]                                               # RP:171
    cq = calib_data[N_cal:2 * N_cal:dec]                                       # RP:172
    return (ci + 1j * cq)                                                      # RP:174 (column)


# --------------------------------------------------------------------------------------
# per-frame chain  (RP:199-260 / RP:458-529)
# --------------------------------------------------------------------------------------
@dataclass
class FrameOut:
    range_fft: np.ndarray            # [NR x PN] complex, as stored at RP:207 (pre-MTI)
    range_max: np.ndarray            # [NR] RP:210
    tgt_range_idx: np.ndarray        # 1-based, possibly empty
    tgt_range_mag: np.ndarray
    doppler_row: np.ndarray          # [ND] complex, fftshifted row at tgt_range_idx(1) (zeros if none)
    tgt_doppler_idx: np.ndarray      # 1-based
    slow_time_row: np.ndarray | None  # [PN] complex, RP:259


def fast_time(chirp_rx1, cal, cfg: Config):
    """RP:203-205."""
    x = (chirp_rx1 - cal[:, None]) * cfg.IF_scale                              # RP:203
    x = x - x.mean(axis=0, keepdims=True)                                      # RP:204
    return sfft.fft(x * cfg.range_window_func[:, None], cfg.range_fft_size, axis=0)   # RP:205


def process_frame(chirp, cal, cfg: Config) -> FrameOut:
    NR, ND = cfg.range_fft_size, cfg.Doppler_fft_size
    m = chirp[:, :, 0]                                                         # RP:202
    rfft = fast_time(m, cal, cfg)
    stored = rfft.copy()                                                       # RP:207
    # RP:210: abs(max(X,[],2)); MATLAB's complex max selects by magnitude
    rmax = np.abs(rfft).max(axis=1)
    idx, mag = f_search_peak(rmax, len(rmax), cfg.range_threshold, cfg.max_num_targets,
                             cfg.min_distance, cfg.max_distance, cfg.dist_per_bin, cfg.peak_mode)   # RP:211
    nt = len(idx)
    drow = np.zeros(ND, dtype=np.complex128)
    didx = np.zeros(nt, dtype=np.int64)
    if nt > 0:
        rows = rfft[idx - 1, :]
        rows = rows - rows.mean(axis=1, keepdims=True)                         # RP:217-218
        rd = sfft.fftshift(sfft.fft(rows * cfg.doppler_window_func[None, :], ND, axis=1), axes=1)  # RP:219
        for j in range(nt):                                                    # RP:232-239
            a = np.abs(rd[j])
            k = int(np.argmax(a))          # first max on ties
            val, dop = a[k], k + 1
            dc = ND // 2 + 1               # the literal 9 at ND = 16
            didx[j] = dop if (val >= cfg.Doppler_threshold and dop != dc) else dc
        drow = rd[0]
    slow = stored[idx[0] - 1, :].copy() if nt > 0 else None                    # RP:257-260
    return FrameOut(stored, rmax, idx, mag, drow, didx, slow)


def range_doppler_map_db(chirp, cal, cfg: Config):
    """RP:216-219 applied to every range row of one frame (the reference fills the detected row only, RP:216-221), in dB:
    20*log10(abs(fftshift(fft((X - mean(X,2)) .* (2*chebwin(PN)).', ND, 2), 2))), [NR x ND]."""
    ND = cfg.Doppler_fft_size
    rfft = fast_time(chirp[:, :, 0], cal, cfg)
    rows = rfft - rfft.mean(axis=1, keepdims=True)
    rd = sfft.fftshift(sfft.fft(rows * cfg.doppler_window_func[None, :], ND, axis=1), axes=1)
    with np.errstate(divide="ignore"):
        return 20 * np.log10(np.abs(rd))


def speed_of(didx, cfg: Config):
    return (didx - cfg.Doppler_fft_size / 2 - 1) * -cfg.fD_per_bin * cfg.Hz_to_mps_constant   # RP:250


# --------------------------------------------------------------------------------------
# STFT + log-frequency resample  (RP:270-299 / RP:538-566)
# --------------------------------------------------------------------------------------
def nextpow2(n):
    """MATLAB nextpow2 for a positive integer: smallest p with 2^p >= n."""
    return 0 if n <= 1 else (int(n) - 1).bit_length()


def matlab_linspace(d1, d2, n):
    n1 = n - 1
    y = d1 + (np.arange(n, dtype=np.float64) * (d2 - d1)) / n1
    y[0], y[-1] = d1, d2
    return y


def matlab_logspace(a, b, n):
    return 10.0 ** matlab_linspace(a, b, n)


def stft_axes(L, cfg: Config):
    """nfft (RP:273), fs, column count, T (RP:276) and log_freq_bins (RP:293-296)."""
    win, nov = cfg.window_length, cfg.overlap
    nfft = 2 ** nextpow2(L)
    fs = 1.0 / cfg.PRT
    hop = win - nov
    ncol = (L - nov) // hop
    T = (win / 2 + np.arange(ncol) * hop) / fs
    df = fs / nfft
    F_min, F_max = df, (nfft // 2) * df          # min(F(F>0)), max(F)
    fq = matlab_logspace(math.log10(F_min), math.log10(F_max), cfg.MAX_FREQ_BINS)
    return nfft, fs, hop, ncol, T, fq


def stft_literal(iq_abs, cfg: Config):
    """RP:270-299 exactly as written (full nfft one-sided PSD).  Memory O(nfft * ncol):
    only for small L."""
    x = np.asarray(iq_abs, dtype=np.float64).reshape(-1)
    L = len(x)
    win = cfg.window_length
    w = _win.kaiser(win, cfg.kaiser_beta, sym=True)
    nfft, fs, hop, ncol, T, fq = stft_axes(L, cfg)
    seg = np.lib.stride_tricks.sliding_window_view(x, win)[::hop][:ncol] * w[None, :]
    S = sfft.fft(seg, nfft, axis=1)[:, :nfft // 2 + 1].T         # one-sided (real input)
    P = (np.abs(S) ** 2) / (fs * np.sum(w ** 2))
    P[1:nfft // 2, :] *= 2                                       # not DC, not Nyquist (nfft even)
    F = np.arange(nfft // 2 + 1) * fs / nfft
    Fs_, Ps_ = sfft.fftshift(F), sfft.fftshift(P, axes=0)         # RP:279-280
    G = Ps_.max()                                                 # RP:282 (max(max(P)))
    with np.errstate(divide="ignore"):
        psd = 20 * np.log10(np.abs(Ps_) / G)                      # RP:283
    # RP:299 interp1(F, psd, fq, 'linear', 'extrap'): interp1 sorts its sample points
    order = np.argsort(Fs_, kind="stable")
    Fo, Po = Fs_[order], psd[order, :]
    j = np.clip(np.searchsorted(Fo, fq, side="right") - 1, 0, len(Fo) - 2)
    a = (fq - Fo[j]) / (Fo[j + 1] - Fo[j])
    out = Po[j, :] + a[:, None] * (Po[j + 1, :] - Po[j, :])
    return dict(T=T, frequency=fq, intensity=out, nfft=nfft, pmax=G, P=P)


def stft_literal_chunked(iq_abs, cfg: Config, col_chunk=2048):
    """The same operations as ``stft_literal`` (full-nfft one-sided PSD, fftshift, max(max(P)), 20*log10, interp1 over
    the sorted fine grid), applied to blocks of columns so that C1's 16,385 x 31,981 matrix P (4.2 GB, RP:276) never has
    to be resident: pass 1 finds max(max(P)) (RP:282), pass 2 normalises and resamples each block (RP:283-299).  Columns
    are independent in every step but the maximum, so the results are bit-identical to ``stft_literal``
    (tests/test_oracle_stft.py)."""
    x = np.asarray(iq_abs, dtype=np.float64).reshape(-1)
    win = cfg.window_length
    w = _win.kaiser(win, cfg.kaiser_beta, sym=True)
    nfft, fs, hop, ncol, T, fq = stft_axes(len(x), cfg)
    segs = np.lib.stride_tricks.sliding_window_view(x, win)[::hop][:ncol]
    F = np.arange(nfft // 2 + 1) * fs / nfft
    Fs_ = sfft.fftshift(F)
    order = np.argsort(Fs_, kind="stable")
    Fo = Fs_[order]
    j = np.clip(np.searchsorted(Fo, fq, side="right") - 1, 0, len(Fo) - 2)
    a = (fq - Fo[j]) / (Fo[j + 1] - Fo[j])

    def block(c0):
        seg = segs[c0:c0 + col_chunk] * w[None, :]
        S = sfft.fft(seg, nfft, axis=1)[:, :nfft // 2 + 1].T
        P = (np.abs(S) ** 2) / (fs * np.sum(w ** 2))
        P[1:nfft // 2, :] *= 2
        return sfft.fftshift(P, axes=0)

    G = max(float(block(c0).max()) for c0 in range(0, ncol, col_chunk))
    out = np.empty((len(fq), ncol))
    for c0 in range(0, ncol, col_chunk):
        with np.errstate(divide="ignore"):
            psd = 20 * np.log10(np.abs(block(c0)) / G)
        Po = psd[order, :]
        out[:, c0:c0 + col_chunk] = Po[j, :] + a[:, None] * (Po[j + 1, :] - Po[j, :])
    return dict(T=T, frequency=fq, intensity=out, nfft=nfft, pmax=G)


def _dtft_bins(fq, nfft, fs):
    df = fs / nfft
    j = np.clip(np.floor(fq / df).astype(np.int64), 0, nfft // 2 - 1)
    a = (fq - j * df) / df
    bins = np.unique(np.concatenate([j, j + 1]))
    pos = np.searchsorted(bins, j)
    return j, a, bins, pos


def stft_global_max(x, w, nfft, hop, ncol):
    """Exact max over the full one-sided fine grid of |S|^2 with the one-sided doubling,
    without materialising P.  Columns are visited in descending order of the
    bound 2*S(0)^2 >= max_j c_j |S(w_j)|^2 (x >= 0, w > 0) and each visited column gets the
    full nfft FFT; the loop stops once the bound of the next column cannot beat the max."""
    win = len(w)
    seg = np.lib.stride_tricks.sliding_window_view(x, win)[::hop][:ncol] * w[None, :]
    s0 = np.abs(seg).sum(axis=1)
    ub = 2 * s0 ** 2
    order = np.argsort(-ub, kind="stable")
    best = -1.0
    seen = set()
    for t in order:
        if ub[t] <= best:
            break
        key = seg[t].tobytes()
        if key in seen:
            continue
        seen.add(key)
        S = sfft.rfft(seg[t], nfft)
        p = np.abs(S) ** 2
        if nfft > 1:
            p[1:nfft // 2] *= 2
        best = max(best, float(p.max()))
    return best


def stft_restated(iq_abs, cfg: Config, pmax_raw=None, col_chunk=8192, col_range=None, L_total=None):
    """Memory-safe algebraic restatement of RP:270-299: the windowed DTFT is evaluated only
    at the fine-grid bins that bracket the log-spaced query frequencies; dB; linear interp.
    Equal to ``stft_literal`` to ~1e-11 dB (tests/test_oracle_stft.py).

    ``pmax_raw`` (max of c_j*|S|^2, unnormalised) may be supplied; else it is computed exactly.
    ``col_range=(c0,c1)`` restricts the output columns (for bounded CPU-baseline samples).
    """
    x = np.asarray(iq_abs, dtype=np.float64).reshape(-1)
    L = len(x) if L_total is None else L_total
    win = cfg.window_length
    w = _win.kaiser(win, cfg.kaiser_beta, sym=True)
    nfft, fs, hop, ncol, T, fq = stft_axes(L, cfg)
    j, a, bins, pos = _dtft_bins(fq, nfft, fs)
    if pmax_raw is None:
        pmax_raw = stft_global_max(x, w, nfft, hop, ncol)
    n = np.arange(win)
    E = np.exp(-2j * np.pi * np.outer(n, bins) / nfft)                       # [win x nb]
    cj = np.where((bins == 0) | (bins == nfft // 2), 1.0, 2.0)
    c0, c1 = (0, ncol) if col_range is None else col_range
    out = np.empty((len(fq), c1 - c0), dtype=np.float64)
    view = np.lib.stride_tricks.sliding_window_view(x, win)[::hop]
    for s in range(c0, c1, col_chunk):
        e = min(s + col_chunk, c1)
        seg = view[s:e] * w[None, :]
        S = seg @ E
        praw = cj[None, :] * (S.real ** 2 + S.imag ** 2)
        with np.errstate(divide="ignore"):
            db = 20 * np.log10(praw / pmax_raw)
        lo, hi = db[:, pos], db[:, pos + 1]
        out[:, s - c0:e - c0] = (lo + a[None, :] * (hi - lo)).T
    return dict(T=T[c0:c1], frequency=fq, intensity=out, nfft=nfft, pmax_raw=pmax_raw, bins=bins)


# --------------------------------------------------------------------------------------
# the two branches of radar_processing()  (RP:195-299, RP:440-607)
# --------------------------------------------------------------------------------------
def radar_processing_no(frames, calib_data, sxml, stft="restated", **cfg_kw):
    """'no' branch, RP:197-299.  Returns every intermediate the payloads need."""
    cfg = configure(sxml, **cfg_kw)
    cal = calib_rx1(calib_data, cfg)
    N = len(frames)
    NR, PN, ND = cfg.range_fft_size, cfg.num_chirps_per_frame, cfg.Doppler_fft_size
    range_max_abs = np.zeros((NR, N))
    det = np.zeros(N, dtype=bool)
    ridx = np.zeros(N, dtype=np.int64)
    rmag = np.zeros(N)
    didx = np.full(N, ND // 2 + 1, dtype=np.int64)
    drows = np.zeros((N, ND), dtype=np.complex128)
    slow = []
    for f in range(N):
        fo = process_frame(frames[f], cal, cfg)
        range_max_abs[:, f] = fo.range_max                                     # RP:265
        if len(fo.tgt_range_idx):
            det[f] = True
            ridx[f], rmag[f], didx[f] = fo.tgt_range_idx[0], fo.tgt_range_mag[0], fo.tgt_doppler_idx[0]
            drows[f] = fo.doppler_row
            slow.append(fo.slow_time_row)
    slow_all = np.concatenate(slow) if slow else np.zeros(0, dtype=np.complex128)
    out = dict(cfg=cfg, range_tx1rx1_max_abs=range_max_abs, detected=det, range_idx=ridx,
               range_mag=rmag, doppler_idx=didx, doppler_rows=drows,
               slow_time_signal_all_frames=slow_all,
               range=np.where(det, (ridx - 1) * cfg.dist_per_bin, 0.0),        # RP:248
               speed=np.where(det, speed_of(didx, cfg), 0.0),                  # RP:250
               strength=np.where(det, rmag, 0.0))                              # RP:245
    if stft and len(slow_all) >= cfg.window_length:
        iq = np.abs(slow_all)                                                  # RP:270
        out["stft"] = stft_literal(iq, cfg) if stft == "literal" else stft_restated(iq, cfg)
    return out


def radar_processing_yes(frames, calib_data, sxml, stft="restated", **cfg_kw):
    """'yes' branch, RP:444-607: per 100-frame batch, at most 4 spectrograms, NaN fill."""
    cfg = configure(sxml, **cfg_kw)
    cal = calib_rx1(calib_data, cfg)
    N = len(frames)
    nb = -(-N // cfg.batch_size)                                               # RP:190
    strength = np.zeros((cfg.max_num_targets, N))
    rng = np.zeros((cfg.max_num_targets, N))
    spd = np.zeros((cfg.max_num_targets, N))
    batches = []
    plot_counter = 0
    for b in range(1, nb + 1):
        s0 = (b - 1) * cfg.batch_size + 1                                      # RP:446
        e0 = min(b * cfg.batch_size, N)                                        # RP:447
        slow = []
        for fr in range(s0, e0 + 1):
            fo = process_frame(frames[fr - 1], cal, cfg)
            if len(fo.tgt_range_idx):                                          # RP:478
                strength[0, fr - 1] = fo.tgt_range_mag[0]
                rng[0, fr - 1] = (fo.tgt_range_idx[0] - 1) * cfg.dist_per_bin
                spd[0, fr - 1] = speed_of(fo.tgt_doppler_idx[0], cfg)
                slow.append(fo.slow_time_row)
            else:
                strength[:, fr - 1] = rng[:, fr - 1] = spd[:, fr - 1] = np.nan  # RP:524-528
        sig = np.concatenate(slow) if slow else np.zeros(0, dtype=np.complex128)
        if len(sig) >= cfg.window_length:                                      # RP:534
            plot_counter += 1
            if plot_counter <= cfg.max_plots:                                  # RP:537
                iq = np.abs(sig)
                r = stft_literal(iq, cfg) if stft == "literal" else stft_restated(iq, cfg)
                r.update(batch=b, start_frame=s0, end_frame=e0)
                batches.append(r)
            else:
                break                                                          # RP:599
    return dict(cfg=cfg, batches=batches, strength=strength, range=rng, speed=spd)
