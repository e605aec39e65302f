"""Host logic of lb_odometry_* (locus_b200/csrc/odometry.cu) without a GPU: the pipeline source is compiled against
fake stages (tests/odometry_stub.h) that tag every cloud with its scan id.  Checks submission-order results, the
source/target pairing of every registration, that no ring slot is overwritten while still in use (random stage
latencies), back-pressure, the worker limit, and per-scan error propagation.  The GPU twin is test_odometry_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import locus_b200
from locus_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEP = 16


@pytest.fixture(scope="module")
def oh():
    bdir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libodometry_harness.so")
    srcs = [os.path.join(ROOT, "tests", "odometry_harness.cpp"), os.path.join(ROOT, "tests", "odometry_stub.h"),
            os.path.join(ROOT, "locus_b200", "csrc", "odometry.cu"), os.path.join(ROOT, "include", "locus_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-x", "c++", srcs[0],
                               "-I", os.path.join(ROOT, "tests"), "-o", so])
    H = C.CDLL(so)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    H.lb_odometry_create.argtypes = [i32, i32, sz, C.c_uint32, C.POINTER(vp)]
    H.lb_odometry_destroy.argtypes = [vp]
    H.lb_odometry_set_gicp_params.argtypes = [vp, C.POINTER(api.GicpParams)]
    H.lb_odometry_submit.argtypes = [vp, vp, sz, C.c_uint32, C.POINTER(api.Field), i32, i32, vp, vp, i32, C.POINTER(C.c_uint64)]
    H.lb_odometry_next.argtypes = [vp, C.POINTER(api.OdometryResult), i32]
    H.lb_odometry_pending.argtypes = [vp, C.POINTER(sz)]
    H.lb_odometry_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    return H


def _scan(i, flags=0, n=8):
    b = np.zeros(n * STEP, dtype=np.uint8)
    b[:8] = np.frombuffer(np.uint64(i).tobytes(), dtype=np.uint8)
    b[8:12] = np.frombuffer(np.uint32(flags).tobytes(), dtype=np.uint8)
    return b


def _fields():
    return api.VoxelGridB200._fields([("x", 0, api.LB_FLOAT32, 1), ("y", 4, api.LB_FLOAT32, 1), ("z", 8, api.LB_FLOAT32, 1)])


@pytest.mark.parametrize("share", [False, True])
@pytest.mark.parametrize("depth", [1, 2, 4, 8])
def test_pipeline_order_pairing_and_ring_safety(oh, depth, share):
    oh.oh_reset(400)
    h = C.c_void_p()
    assert oh.lb_odometry_create(0, depth, 8, STEP, C.byref(h)) == 0
    oh.lb_odometry_set_cloud_sharing.argtypes = [C.c_void_p, C.c_int]
    assert oh.lb_odometry_set_cloud_sharing(h, int(share)) == 0
    p = api.GicpParams(); p.max_iterations = 37
    assert oh.lb_odometry_set_gicp_params(h, C.byref(p)) == 0
    fa = _fields()
    n_scans = 150
    ids = [1000 + 3 * i for i in range(n_scans)]
    scans = [_scan(i) for i in ids]
    fout = [np.zeros(8 * STEP, dtype=np.uint8) for _ in scans]
    got = []
    r = api.OdometryResult()
    t = C.c_uint64(0)
    guess = np.zeros(16, dtype=np.float32)
    for k, s in enumerate(scans):
        guess[3] = k
        assert oh.lb_odometry_submit(h, s.ctypes.data_as(C.c_void_p), 8, STEP, fa, 3, 0, guess.ctypes.data_as(C.c_void_p),
                                     fout[k].ctypes.data_as(C.c_void_p), 0, C.byref(t)) == 0
        assert t.value == k
        while True:                      # poll: results only ever come back in submission order
            st = oh.lb_odometry_next(h, C.byref(r), 0)
            if st != 0:
                assert st in (1, -10)
                break
            got.append((r.ticket, r.status, r.has_pose, r.n_filtered, list(r.gicp.final_transformation[:5]), r.gicp.iterations))
    n = C.c_size_t(0)
    while oh.lb_odometry_pending(h, C.byref(n)) == 0 and n.value:
        assert oh.lb_odometry_next(h, C.byref(r), 1) == 0
        got.append((r.ticket, r.status, r.has_pose, r.n_filtered, list(r.gicp.final_transformation[:5]), r.gicp.iterations))
    assert [g[0] for g in got] == list(range(n_scans))
    for k, (ticket, status, has_pose, n_f, T, iters) in enumerate(got):
        assert status == 0
        assert n_f == 3 + ids[k] % 5
        assert np.frombuffer(fout[k][:8].tobytes(), dtype=np.uint64)[0] == ids[k]        # filtered cloud handed back
        if k == 0:
            assert has_pose == 0
        else:
            assert has_pose == 1 and iters == 37
            assert T[0] == ids[k] and T[1] == ids[k - 1], (k, T)                       # source = scan k, target = scan k-1
            assert T[2] == 3 + ids[k] % 5 and T[3] == 3 + ids[k - 1] % 5
            assert T[4] == k                                                               # the caller's prior reached align()
    assert 1 <= oh.oh_aligns_peak() <= depth
    cnt = C.c_uint64(0)
    assert oh.lb_odometry_launch_count(h, C.byref(cnt)) == 0
    assert cnt.value == 17 * n_scans + 40 * (n_scans - 1) + (12 * n_scans if share else 0)   # one preparation per scan when sharing
    # cloud sharing cannot be switched once scans have gone through
    assert oh.lb_odometry_set_cloud_sharing(h, int(not share)) == -9
    assert oh.lb_odometry_next(h, C.byref(r), 1) == -10
    assert oh.lb_odometry_destroy(h) == 0
    assert oh.oh_clouds_alive() == 0           # every shared-cloud reference was released
    if share:                                  # a published reference lives until the next worker has adopted it:
        assert 1 <= oh.oh_clouds_peak() <= depth + 1     # never more than the scans being worked on, plus one


@pytest.mark.parametrize("share", [False, True])
def test_pipeline_error_propagation(oh, share):
    """a scan that fails in the VoxelGrid stage: no pose for it nor for the next one, in both modes"""
    oh.oh_reset(50)
    h = C.c_void_p()
    assert oh.lb_odometry_create(0, 3, 8, STEP, C.byref(h)) == 0
    oh.lb_odometry_set_cloud_sharing.argtypes = [C.c_void_p, C.c_int]
    assert oh.lb_odometry_set_cloud_sharing(h, int(share)) == 0
    fa = _fields()
    t = C.c_uint64(0)
    flags = [0, 0, 1, 0, 0, 0]
    scans = [_scan(10 + i, f) for i, f in enumerate(flags)]
    for s_ in scans:
        assert oh.lb_odometry_submit(h, s_.ctypes.data_as(C.c_void_p), 8, STEP, fa, 3, 0, None, None, 0, C.byref(t)) == 0
    r = api.OdometryResult()
    poses = []
    for _ in scans:
        assert oh.lb_odometry_next(h, C.byref(r), 1) == 0
        poses.append((r.has_pose, r.status))
    assert [p[0] for p in poses] == [0, 1, 0, 0, 1, 1]
    assert poses[2][1] != 0 and poses[3][1] == 0
    assert oh.lb_odometry_destroy(h) == 0
    assert oh.oh_clouds_alive() == 0


def test_pipeline_error_propagation_and_limits(oh):
    oh.oh_reset(50)
    h = C.c_void_p()
    assert oh.lb_odometry_create(0, 0, 8, STEP, C.byref(h)) == -1
    assert oh.lb_odometry_create(0, 17, 8, STEP, C.byref(h)) == -1
    assert oh.lb_odometry_create(0, 2, 8, STEP, C.byref(h)) == 0
    fa = _fields()
    t = C.c_uint64(0)
    big = _scan(1, n=9)
    assert oh.lb_odometry_submit(h, big.ctypes.data_as(C.c_void_p), 9, STEP, fa, 3, 0, None, None, 0, C.byref(t)) == -8
    nofield = api.VoxelGridB200._fields([("a", 0, api.LB_FLOAT32, 1)])
    assert oh.lb_odometry_submit(h, big.ctypes.data_as(C.c_void_p), 8, STEP, nofield, 1, 0, None, None, 0, C.byref(t)) == -1
    flags = [0, 0, 1, 0, 0]               # scan 2 fails in the voxel stage
    scans = [_scan(10 + i, f) for i, f in enumerate(flags)]
    for s in scans:
        assert oh.lb_odometry_submit(h, s.ctypes.data_as(C.c_void_p), 8, STEP, fa, 3, 0, None, None, 0, C.byref(t)) == 0
    r = api.OdometryResult()
    out = []
    for _ in scans:
        assert oh.lb_odometry_next(h, C.byref(r), 1) == 0
        out.append((r.status, r.has_pose, r.error.decode()))
    assert [o[1] for o in out] == [0, 1, 0, 0, 1]       # scan 2 has no pose, scan 3 has nothing to register against
    assert out[2][0] != 0 and "stub voxel failure" in out[2][2]
    assert out[3][0] == 0
    # the pipeline must be idle to change parameters
    p = api.GicpParams()
    assert oh.lb_odometry_submit(h, scans[0].ctypes.data_as(C.c_void_p), 8, STEP, fa, 3, 0, None, None, 0, C.byref(t)) == 0
    st = oh.lb_odometry_set_gicp_params(h, C.byref(p))
    assert st in (0, -1)                                  # not idle until its result has been returned
    assert oh.lb_odometry_next(h, C.byref(r), 1) == 0
    assert oh.lb_odometry_set_gicp_params(h, C.byref(p)) == 0
    assert oh.lb_odometry_destroy(h) == 0


def test_destroy_with_scans_in_flight(oh):
    oh.oh_reset(2000)
    h = C.c_void_p()
    assert oh.lb_odometry_create(0, 3, 8, STEP, C.byref(h)) == 0
    fa = _fields()
    scans = [_scan(i) for i in range(6)]
    t = C.c_uint64(0)
    for s in scans:
        assert oh.lb_odometry_submit(h, s.ctypes.data_as(C.c_void_p), 8, STEP, fa, 3, 0, None, None, 0, C.byref(t)) == 0
    assert oh.lb_odometry_destroy(h) == 0      # joins the stage threads; nothing may touch `scans` afterwards
