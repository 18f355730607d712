"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/x266_b200.h declares
(no compute calls -- there is no GPU here), and fails loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "x266_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    funcs = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(funcs) | {"g_t32"})


def test_header_declares_the_reference_entry_points():
    syms = header_symbols()
    for s in ("dct32_genNew", "dct32_getDiff", "dct32_getDct", "satd8x8_genNew", "satd8x8_getDiff", "satd8x8_getSatd",
              "partialButterfly32", "satd8x8", "g_t32", "xDct32Batch", "xSatd8x8Batch", "xSatd8x8Search", "xIntra32Pred"):
        assert s in syms


def test_library_exports_every_declared_symbol(x266):
    L = ctypes.CDLL(x266.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_g_t32_symbol_equals_oracle(x266, orc):
    assert (x266.g_t32() == orc.g32()).all()


def test_no_oracle_in_product():
    """The product library must not link or reference anything under oracle/."""
    import subprocess
    import x266_b200
    out = subprocess.run(["nm", "-D", x266_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out and "ref_" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "x266_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("no CPU", ""), f"{f} mentions oracle"


def test_fails_loudly_without_gpu(x266):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(x266.X266Error):
        x266.xDct32Batch(np.zeros(1024, np.int16))
    with pytest.raises(x266.X266Error):
        x266.xSatd8x8Batch(np.zeros(64, np.int16))


def test_host_copy_pool(tmp_path):
    """host logic of the pageable path (no GPU): the staging copy pool is exact for every thread count / store kind, with
    misaligned ends and concurrent callers"""
    import subprocess
    csrc = os.path.join(ROOT, "x266_b200", "csrc")
    exe = str(tmp_path / "hostcopy_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", csrc, os.path.join(ROOT, "tests", "c", "hostcopy_test.cpp"),
                           os.path.join(csrc, "hostcopy.cpp"), "-o", exe, "-lpthread"])
    assert subprocess.check_output([exe], text=True, stderr=subprocess.DEVNULL).strip() == "ok"
