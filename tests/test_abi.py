"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol declared in
include/fmcw_cuda.h (no compute calls without a GPU), and the ctypes mirror matches the header."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fmcw_radar_processing_b200 import build as B
    B.build()
    from fmcw_radar_processing_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "fmcw_cuda.h")).read()
    declared = set(re.findall(r"FMCW_API\s+[\w\s\*]+?\b(fmcw_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    l = lib.load()
    for name in declared:
        assert hasattr(l, name), name
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)


def test_every_entry_point_is_documented_in_integration_md():
    """INTEGRATION.md names, for every C entry point, the reference lines it replaces (or says the reference has none)."""
    hdr = open(os.path.join(ROOT, "include", "fmcw_cuda.h")).read()
    declared = set(re.findall(r"FMCW_API\s+[\w\s\*]+?\b(fmcw_\w+)\s*\(", hdr))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = sorted(n for n in declared if n not in doc)
    assert not missing, missing


def test_struct_layout_matches_header(lib):
    assert C.sizeof(lib.fmcw_config) == 14 * 4 + 18 * 8
    assert C.sizeof(lib.fmcw_frame_out) == 7 * 8
    assert C.sizeof(lib.fmcw_stft_out) == 32
    assert C.sizeof(lib.fmcw_run_info) == 9 * 8 + 8 + 8
    assert C.sizeof(lib.fmcw_device_info) == C.sizeof(lib.fmcw_run_info) + 8
    assert lib.load().fmcw_version().startswith(b"libfmcw_cuda")


def test_create_without_gpu_fails_loudly(lib):
    """No CPU fallback: without a CUDA device fmcw_create returns FMCW_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fmcw_radar_processing_b200.api import FmcwCuda
    from fmcw_radar_processing_b200.config import fmcw_configurations
    from fmcw_radar_processing_b200.parse import make_sxml
    with pytest.raises(lib.FmcwError) as ei:
        FmcwCuda(fmcw_configurations(make_sxml()), None)
    assert ei.value.status == 3


def test_axes_are_host_only_and_match_oracle(lib):
    """fmcw_stft_axes needs no GPU: T (RP:276) and log_freq_bins (RP:293-296)."""
    import numpy as np
    from fmcw_radar_processing_b200.config import fmcw_configurations, to_c_config
    from fmcw_radar_processing_b200.parse import make_sxml
    from oracle import fmcw_oracle as O
    sx = make_sxml()
    c = to_c_config(fmcw_configurations(sx))
    l = lib.load()
    for L in (20, 1000, 32000, 320000, 12_800_000):
        nfft, nct = C.c_uint64(), C.c_uint64()
        F = np.empty(1024)
        T = np.empty(5)
        assert l.fmcw_stft_axes(C.byref(c), L, 3, 5 if L > 30 else 1, T.ctypes.data, F.ctypes.data, C.byref(nfft), C.byref(nct)) == 0
        onfft, fs, hop, ncol, oT, ofq = O.stft_axes(L, O.configure(sx))
        assert nfft.value == onfft and nct.value == ncol
        assert np.allclose(F, ofq, rtol=1e-14, atol=0)
        if L > 30:
            assert np.allclose(T, oT[3:8], rtol=1e-15, atol=0)
