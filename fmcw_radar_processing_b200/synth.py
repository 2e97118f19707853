"""Synthetic FMCW scene generator.

A counter-based generator: every ADC sample is a pure function of
``(seed, global frame, rx, chirp, sample)``, so any frame range can be produced
independently on the CPU (this file, NumPy float64) or on any GPU shard
(``fmcw_synth_frames`` in ``csrc/synth.cu``) with the same bits.

Sample model for frame f, rx r, chirp m, sample n (ADC codes, 12 bit)::

    s = dc*(1+1j) + sum_i A_i * exp(j*2*pi*(kr_i(f)*n + dop_i(f)*m + ph_i(f) + r*rx_step))
        + sigma*(g1 + 1j*g2)
    code = clip(floor(s + 0.5), 0, 4095)          (separately for I and Q)

``kr_i`` (cycles per sample = range bin / NR), ``dop_i`` (cycles per chirp =
2*v*PRT/lambda) and ``ph_i`` are tabulated per frame on the host in float64
(``scene_tables``); the device only evaluates the formula above with separate
multiplies and adds (no FMA contraction), so CPU and GPU differ at most in the last
ulp of sin/cos, which changes a code only if ``s`` sits within ~1e-13 of a rounding
boundary.  ``g1, g2`` are Irwin-Hall(4) variates built from one splitmix64 hash each
with integer arithmetic only (unit variance).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
IH4_OFFSET = 131070.0                       # 4 * 65535 / 2
IH4_INV_STD = 1.0 / 37837.226637012174      # sqrt((65536^2 - 1) / 3)


@dataclass
class Scatterer:
    A: float            # amplitude in ADC LSB
    R0: float           # range at t = 0 [m]
    v: float            # radial velocity [m/s] (R(t) = R0 + v t)
    phi: float = 0.0    # phase offset [cycles]
    limb_a: float = 0.0     # micro-motion amplitude [m]
    limb_f: float = 0.0     # micro-motion frequency [Hz]
    limb_phi: float = 0.0   # micro-motion phase [rad]


@dataclass
class Scene:
    seed: int
    scatterers: List[Scatterer] = field(default_factory=list)
    sigma: float = 2.0      # noise floor, LSB
    dc: float = 2048.0
    rx_step: float = 0.11   # per-RX phase offset [cycles]
    frame_time: float = 0.15
    r_min: float = 2.0      # targets walk back and forth between r_min and r_max (inside the 0.9-25 m gate, RP:126-127)
    r_max: float = 22.0


def scene_c1(seed=1):
    """C1/C3: single moving point target, A = 900 LSB, R0 = 10 m, v = -1.0 m/s."""
    return Scene(seed=seed, scatterers=[Scatterer(A=900.0, R0=10.0, v=-1.0)])


def scene_c2(seed=2):
    """C2: walking animal: torso + 4 limb scatterers with 2 Hz micro-motion."""
    sc = [Scatterer(A=900.0, R0=12.0, v=0.8)]
    for i, (a, ph) in enumerate([(0.05, 0.0), (0.08, 1.3), (0.12, 2.6), (0.15, 4.1)]):
        sc.append(Scatterer(A=300.0, R0=12.0, v=0.8, phi=0.17 * (i + 1), limb_a=a, limb_f=2.0, limb_phi=ph))
    return Scene(seed=seed, scatterers=sc)


def scene_tables(scene: Scene, dist_per_bin: float, range_fft_size: int, PRT: float, lam: float,
                 frame0: int, n_frames: int) -> np.ndarray:
    """float64 [n_frames][n_scat][4] = (A, kr cycles/sample, dop cycles/chirp, ph cycles)."""
    f = frame0 + np.arange(n_frames, dtype=np.float64)
    t = f * scene.frame_time
    tab = np.zeros((n_frames, max(1, len(scene.scatterers)), 4), dtype=np.float64)
    for i, s in enumerate(scene.scatterers):
        arg = 2 * np.pi * s.limb_f * t + s.limb_phi
        span = scene.r_max - scene.r_min
        u = np.mod(s.R0 + s.v * t - scene.r_min, 2 * span)          # triangle path: reflect at r_min / r_max
        fwd = u <= span
        R = scene.r_min + np.where(fwd, u, 2 * span - u) + s.limb_a * np.sin(arg)
        vr = np.where(fwd, s.v, -s.v) + s.limb_a * 2 * np.pi * s.limb_f * np.cos(arg)
        tab[:, i, 0] = s.A
        tab[:, i, 1] = (R / dist_per_bin) / range_fft_size
        tab[:, i, 2] = 2.0 * vr * PRT / lam
        tab[:, i, 3] = s.phi
    return tab


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _ih4(h: np.ndarray) -> np.ndarray:
    m = np.uint64(0xFFFF)
    u = (h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))
    return (u.astype(np.float64) - IH4_OFFSET) * IH4_INV_STD


def synth_frames(tables: np.ndarray, seed: int, frame0: int, n_rx: int, PN: int, NTS: int,
                 sigma: float = 2.0, dc: float = 2048.0, rx_step: float = 0.11) -> np.ndarray:
    """int16 [n_frames][n_rx][PN][NTS][2] ADC codes for global frames frame0 .. frame0+n-1."""
    n_frames, n_scat, _ = tables.shape
    out = np.empty((n_frames, n_rx, PN, NTS, 2), dtype=np.int16)
    n = np.arange(NTS, dtype=np.float64)[None, None, :]
    m = np.arange(PN, dtype=np.float64)[None, :, None]
    r = np.arange(n_rx, dtype=np.float64)[:, None, None]
    per_frame = n_rx * PN * NTS
    lin = np.arange(per_frame, dtype=np.uint64).reshape(n_rx, PN, NTS)
    with np.errstate(over="ignore"):
        seedmix = np.uint64(seed) * _GOLD
    for k in range(n_frames):
        sI = np.full((n_rx, PN, NTS), dc, dtype=np.float64)
        sQ = np.full((n_rx, PN, NTS), dc, dtype=np.float64)
        for i in range(n_scat):
            A, kr, dop, ph = tables[k, i]
            if A == 0.0:
                continue
            p = ((kr * n + dop * m) + ph) + r * rx_step
            p = p - np.floor(p)
            ang = 2.0 * np.pi * p
            sI = sI + A * np.cos(ang)
            sQ = sQ + A * np.sin(ang)
        with np.errstate(over="ignore"):
            idx = np.uint64(frame0 + k) * np.uint64(per_frame) + lin
            kI = seedmix + idx * np.uint64(2)
            kQ = kI + np.uint64(1)
        sI = sI + sigma * _ih4(splitmix64(kI))
        sQ = sQ + sigma * _ih4(splitmix64(kQ))
        out[k, ..., 0] = np.clip(np.floor(sI + 0.5), 0, 4095).astype(np.int16)
        out[k, ..., 1] = np.clip(np.floor(sQ + 0.5), 0, 4095).astype(np.int16)
    return out


def default_calib(n_rx: int, NTS: int, dc: float = 2047.5) -> np.ndarray:
    """Calibration codes, row vector [I_rx1 Q_rx1 I_rx2 Q_rx2 ...] (RP:167-172), constant so
    that the ADC mid-scale cancels (0.5 + 0.5j in normalised units)."""
    return np.full(2 * n_rx * NTS, dc, dtype=np.float64)
