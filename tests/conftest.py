"""pytest configuration: the `gpu` marker, import paths, and the CPU-side test binaries."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_DIR = os.path.join(ROOT, "finalproject-losslessimagecompression_b200")
BUILD_DIR = os.path.join(ROOT, "tests", "_build")
HOST_HARNESS_SO = os.path.join(BUILD_DIR, "libflic_host.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def build_host_harness() -> str:
    """g++ build of tests/host_harness.cpp: the device arithmetic header compiled for the host."""
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    hdr = os.path.join(PKG_DIR, "csrc", "flic_core.cuh")
    os.makedirs(BUILD_DIR, exist_ok=True)
    if (not os.path.exists(HOST_HARNESS_SO)
            or os.path.getmtime(HOST_HARNESS_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17",
                               "-I", os.path.join(PKG_DIR, "csrc"), "-x", "c++", src, "-o", HOST_HARNESS_SO])
    return HOST_HARNESS_SO


@pytest.fixture(scope="session")
def host_harness():
    import ctypes as C
    H = C.CDLL(build_host_harness())
    f32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    H.hh_expf.restype = C.c_float
    H.hh_expf.argtypes = [C.c_float]
    H.hh_lower.argtypes = [C.c_float]
    H.hh_part1.argtypes = [C.c_float]
    H.hh_cdf.argtypes = [C.c_int, C.c_float, C.c_float]
    H.hh_tables.argtypes = [f32p, f32p, f32p, C.c_int64, u32p, u32p]
    H.hh_encode.argtypes = [f32p, f32p, f32p, C.c_int64, u32p, C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
    H.hh_decode.argtypes = [u32p, C.c_int64, C.c_uint64, f32p, f32p, C.c_int64, f32p, C.POINTER(C.c_uint64),
                            C.POINTER(C.c_int64)]
    H.hh_decode_fast.argtypes = [u32p, C.c_int64, C.c_uint64, f32p, f32p, C.c_int64, f32p, C.POINTER(C.c_uint64)]
    H.hh_div_check.restype = C.c_int64
    H.hh_div_check.argtypes = [C.c_int64, C.c_uint64, C.c_int]
    H.hh_push_check.restype = C.c_int64
    H.hh_push_check.argtypes = [C.c_int64, C.c_uint64]
    return H


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle
