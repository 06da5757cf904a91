"""The C-ABI library loads and exports every symbol include/physher_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

import physher_b200 as phb
from physher_b200 import build as phb_build
from physher_b200.treelikelihood import LIB_PATH, SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    phb_build.build()
    return phb.load_library()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "physher_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert len(names) >= 30
    raw = ctypes.CDLL(LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/physher_b200.h but not exported"


def test_python_mirror_binds_every_declared_symbol():
    assert sorted(n for n, _, _ in SYMBOLS) == declared_functions()


def test_no_cpu_fallback_without_gpu(lib):
    """Without a CUDA device construction must fail loudly (PHB_ECUDA), never compute on the CPU."""
    if phb.device_count() > 0:
        pytest.skip("GPU present")
    import numpy as np

    with pytest.raises(phb.PhysherB200Error, match="no CUDA device"):
        phb.SingleTreeLikelihood(np.array([-1, -1, 0], np.int32), np.array([-1, -1, 1], np.int32), 2, 4, 1, 8)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (task ③)."""
    pkg = os.path.join(ROOT, "physher_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"
