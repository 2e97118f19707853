"""The MEX gateway and the Node-API addon cannot be built here (no MATLAB, no Node headers); they are
syntax-checked against the minimal stub headers and must only bind entry points the C ABI declares."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("src,stub", [("mex/fmcw_cuda_mex.cpp", "mex/stub"), ("node/fmcw_napi.cc", "node/stub")])
def test_gateway_compiles_against_stub_headers(src, stub):
    if not shutil.which("g++"):
        pytest.skip("no g++")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", f"-I{ROOT}/{stub}", f"-I{ROOT}/include", f"{ROOT}/{src}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("src", ["mex/fmcw_cuda_mex.cpp", "node/fmcw_napi.cc"])
def test_gateway_binds_declared_entry_points_only(src):
    hdr = open(os.path.join(ROOT, "include", "fmcw_cuda.h")).read()
    declared = set(re.findall(r"FMCW_API\s+[\w\s\*]+?\b(fmcw_\w+)\s*\(", hdr))
    used = set(re.findall(r"\b(fmcw_[a-z_]+)\s*\(", open(os.path.join(ROOT, src)).read()))
    used -= {"fmcw_cuda_mex", "fmcw_napi", "fmcw_gpu_chain"}
    assert used and used <= declared, used - declared
