"""Radar / algorithm configuration: the locals of radar_processing.m RP:89-179 gathered under the
field names of the reference's (commented-out) ``fmcw_configurations`` struct, RP:645-672."""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

from . import _lib

C0 = 3e8  # RP:67


def _txt(node) -> float:
    """str2double(node.Text) for an xml2struct-shaped leaf."""
    return float(node["Text"])


def fmcw_configurations(sXML, window_length: int = 20, overlap: int | None = None, range_fft_size: int = 256,
                        Doppler_fft_size: int = 16, rx_select: int = 1, peak_mode: str = "first",
                        kaiser_beta: float = 3.0, MAX_FREQ_BINS: int = 1024, adc_scale: float = 4095.0,
                        batch_size: int = 100) -> "OrderedDict[str, float]":
    """RP:89-154, 178-179.  ``sXML`` is the xml2struct-shaped dict the parser returns.  ``rx_select`` is
    1-based like the reference's ``matrix_raw_data(:,:,1)`` (RP:202)."""
    dev = sXML["Device"]
    frame_time = 150 * 1e-3                                                   # RP:91
    up = _txt(dev["BaseEndpoint"]["chirpDuration_ns"]) * 1e-9                # RP:94
    PRT = up + 200e-6 + 300e-6                                                # RP:95-97
    fm = dev["FmcwEndpoint"]["FmcwConfiguration"]
    hi, lo = _txt(fm["upperFrequency_kHz"]), _txt(fm["lowerFrequency_kHz"])
    BW = (hi - lo) * 1e3                                                      # RP:100
    fC = (hi + lo) / 2 * 1e3                                                  # RP:106
    NTS = int(_txt(dev["BaseEndpoint"]["FrameFormat"]["numSamplesPerChirp"]))         # RP:109
    PN = int(_txt(dev["BaseEndpoint"]["FrameFormat"]["numChirpsPerFrame"]))           # RP:112
    lam = C0 / fC                                                             # RP:133
    R_max = NTS * C0 / (2 * BW)                                               # RP:142
    fD_max = 1 / (2 * PRT)                                                    # RP:152
    cfg = OrderedDict(
        frame_time=frame_time, PRT=PRT, Bandwidth=BW,
        num_Tx_antennas=int(_txt(dev["BaseEndpoint"]["DeviceInfo"]["numAntennasTx"])),     # RP:102
        num_Rx_antennas=int(_txt(dev["BaseEndpoint"]["DeviceInfo"]["numAntennasRx"])),     # RP:103
        carrier_frequency=fC, num_ADC_samples_per_chirp=NTS, num_chirps_per_frame=PN,
        sampling_frequency=_txt(dev["AdcxmcEndpoint"]["AdcxmcConfiguration"]["samplerateHz"]),   # RP:115
        range_fft_size=range_fft_size, Doppler_fft_size=Doppler_fft_size,      # RP:118-119
        IF_scale=16 * 3.3 * range_fft_size / NTS,                              # RP:121
        range_threshold=200.0, Doppler_threshold=50.0,                         # RP:123-124
        min_distance=0.9, max_distance=25.0, max_num_targets=1,                # RP:126-129
        **{"lambda": lam}, Hz_to_mps_constant=lam / 2,                         # RP:133-135
        R_max=R_max, dist_per_bin=R_max / range_fft_size,                      # RP:142-147
        fD_max=fD_max, fD_per_bin=fD_max / Doppler_fft_size,                   # RP:152-153
        window_length=window_length,                                           # RP:178
        max_slider_index=None,                                                 # RP:287, known after the STFT
        overlap=window_length - 1 if overlap is None else overlap,             # RP:179
    )
    # literals of the reference that the library takes as parameters
    cfg.update(kaiser_beta=kaiser_beta, MAX_FREQ_BINS=MAX_FREQ_BINS, adc_scale=adc_scale, rx_select=rx_select,
               peak_mode=peak_mode, batch_size=batch_size, max_plots=4)
    return cfg


def to_c_config(cfg) -> _lib.fmcw_config:
    c = _lib.fmcw_config()
    c.struct_size = C.sizeof(_lib.fmcw_config)
    for name, _ in _lib.fmcw_config._fields_:
        if name in ("struct_size", "reserved0"):
            continue
        key = "lambda" if name == "lambda_" else name
        v = cfg[key]
        if name == "rx_select":
            v = int(v) - 1                       # MATLAB 1-based -> C 0-based
        elif name == "peak_mode":
            v = {"strongest": _lib.PEAK_STRONGEST, "first": _lib.PEAK_FIRST}[v] if isinstance(v, str) else int(v)
        setattr(c, name, v)
    return c


def array_bin_range(cfg):
    """RP:149."""
    import numpy as np
    return np.arange(cfg["range_fft_size"]) * cfg["dist_per_bin"]
