"""pytest configuration: the `gpu` marker, oracle / emulation builds, shared fixtures."""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The C-ABI library is the product: build it in-tree if it is missing or older than its sources
    # (nvcc cross-compiles sm_100a without a GPU; a minute on the first run, a no-op afterwards).
    from honeybadger_b200 import build as hb_build
    hb_build.build()


def _has_gpu() -> bool:
    try:
        from honeybadger_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def emul():
    """Host build of the CUDA chain bodies (tests/emul/hb_emul.cpp) — a test aid, not a product path."""
    import ctypes as C
    so = os.path.join(EMUL_DIR, "libhb_emul.so")
    src = os.path.join(EMUL_DIR, "hb_emul.cpp")
    csrc = os.path.join(ROOT, "honeybadger_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared",
                               "-ffp-contract=off", "-mfma", f"-I{csrc}", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    for f in ("emul_x87_mean", "emul_x87_sd", "emul_div_by_const", "emul_binom_logpmf_host",
              "emul_binom_logpmf_dev"):
        getattr(L, f).restype = C.c_double
    L.emul_x87_mean.argtypes = [C.c_void_p, C.c_int64]
    L.emul_x87_sd.argtypes = [C.c_void_p, C.c_int64]
    L.emul_x87_addcheck.restype = C.c_int64
    L.emul_x87_addcheck.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_int]
    L.emul_x87_sd_blocked.restype = C.c_double
    L.emul_x87_sd_blocked.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L.emul_x87_blockcheck.restype = C.c_int64
    L.emul_x87_blockcheck.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_void_p]
    L.emul_x87_paircheck.restype = C.c_int64
    L.emul_x87_paircheck.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_int, C.c_void_p]
    L.emul_x87_pair_add80.restype = C.c_int64
    L.emul_x87_pair_add80.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_int]
    L.emul_x87_pair_cover.argtypes = [C.c_void_p, C.c_int]
    L.emul_div_by_const.argtypes = [C.c_double, C.c_double]
    L.emul_binom_logpmf_host.argtypes = [C.c_double] * 3
    L.emul_binom_logpmf_dev.argtypes = [C.c_double] * 3
    L.emul_exp_pair.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
    L.emul_norm3_fast.restype = C.c_int
    L.emul_set_tiled.argtypes = [C.c_int]
    L.emul_set_packed.argtypes = [C.c_int]
    L.emul_binom2_fast.restype = C.c_int
    return L


@pytest.fixture(scope="session")
def golden_allele():
    return np.load(os.path.join(GOLDEN, "allele_mgh31.npz"))


@pytest.fixture(scope="session")
def golden_gexp():
    return np.load(os.path.join(GOLDEN, "gexp_mgh31.npz"))


@pytest.fixture(scope="session")
def golden_inputs():
    return np.load(os.path.join(GOLDEN, "inputs_raw_mgh31.npz"))
