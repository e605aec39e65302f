import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import locus_b200
        return locus_b200.device_count() > 0
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def harness():
    """TEST-ONLY CPU build of the product's __host__ __device__ headers (tests/hd_harness.cpp)."""
    import ctypes as C
    bdir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libhd_harness.so")
    src = os.path.join(ROOT, "tests", "hd_harness.cpp")
    hdrs = [os.path.join(ROOT, "locus_b200", "csrc", h) for h in ("hd.h", "grid.h", "bfgs.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    H = C.CDLL(so)
    H.hh_grid_build.restype = C.c_void_p
    H.hh_grid_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
    H.hh_grid_free.argtypes = [C.c_void_p]
    H.hh_grid_dims.argtypes = [C.c_void_p, C.c_void_p]
    H.hh_nn1_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    H.hh_knn_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    H.hh_cov_knn.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    H.hh_mahalanobis.argtypes = [C.c_void_p] * 4
    H.hh_apply_state.argtypes = [C.c_void_p, C.c_void_p]
    H.hh_align.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_double,
                           C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p]
    return H


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ndt_harness():
    """TEST-ONLY CPU build of the product's NDT header (locus_b200/csrc/ndt.h) with a serial backend (tests/ndt_harness.cpp)."""
    import ctypes as C
    bdir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libndt_harness.so")
    src = os.path.join(ROOT, "tests", "ndt_harness.cpp")
    hdrs = [os.path.join(ROOT, "locus_b200", "csrc", h) for h in ("hd.h", "ndt.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-o", so, src])
    H = C.CDLL(so)
    vp = C.c_void_p
    H.hn_target_build.restype = vp
    H.hn_target_build.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double, C.c_int, C.c_double]
    H.hn_target_free.argtypes = [vp]
    H.hn_target_info.argtypes = [vp] * 4
    H.hn_target_leaves.argtypes = [vp] * 6
    H.hn_svd6_rounds.argtypes = [vp, vp, vp]
    H.hn_svd6_serial.argtypes = [vp, vp, vp]
    H.hn_neighbours.argtypes = [vp, vp, C.c_int, vp, vp]
    H.hn_eval.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]
    H.hn_eval_grouped.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]
    H.hn_align.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_double, C.c_double, C.c_int] + [vp] * 6
    return H
