"""SURVEY 8f row f4 on the GPU: lb_ndt_* (through the C ABI) against oracle/ndt_oracle.c -- the target's voxel
Gaussians bit for bit, the score / gradient / Hessian sums of one evaluation, align() poses, error semantics."""
import numpy as np
import pytest

import fixtures as F
from tools import gen_lidar as G

pytestmark = pytest.mark.gpu

TOL_T, TOL_R = 1e-4, 1e-4          # north_star's pose bar (m, rad)


def _finite(a):
    return np.ascontiguousarray(a[np.isfinite(a).all(1)])


def _scans(seed=3, beams=32, az=1024, n=2):
    scene = G.make_scene(seed)
    poses = G.trajectory(max(n, 3), seed)
    return [_finite(G.scan(scene, poses[i], 5 + i, beams=beams, az=az).view(np.float32).reshape(-1, 8)[:, :3].copy()) for i in range(n)]


def _gpu(tgt, method=0, eps=0.1, resolution=1.0, maxit=35):
    import locus_b200
    n = locus_b200.NdtB200(0)
    n.setResolution(resolution); n.setNeighborhoodSearchMethod(method); n.setTransformationEpsilon(eps); n.setMaximumIterations(maxit)
    n.setInputTarget(tgt)
    return n


def _assert_voxels_equal(n, T):
    a, b = n.targetVoxels(), T.leaves()
    assert len(a["leaf_idx"]) == T.n_valid
    for k in ("leaf_idx", "nr_points", "centroid", "mean", "icov"):
        assert np.array_equal(a[k], b[k], equal_nan=True), (k, np.flatnonzero((a[k] != b[k]).reshape(len(a[k]), -1).any(1))[:5])


def test_target_voxels_bit_exact(oracle):
    """setInputTarget: voxel membership, point counts, float centroids, double means and inverse covariances of every
    searchable voxel equal the oracle's bit for bit (one thread per voxel adds its points in input order)."""
    s0, s1 = _scans()
    for res in (1.0, 0.5, 2.0):
        T = oracle.NdtTarget(s0, oracle.ndt_params(resolution=res))
        _assert_voxels_equal(_gpu(s0, resolution=res), T)
    bad = s0.copy(); bad[::97, 1] = np.nan; bad[5, 0] = np.inf            # non-finite target points are skipped
    _assert_voxels_equal(_gpu(bad), oracle.NdtTarget(bad, oracle.ndt_params()))
    rng = np.random.default_rng(1)                                        # planar target: eigenvalue floor in every voxel
    plane = np.zeros((4000, 3), np.float32); plane[:, :2] = rng.uniform(-10, 10, (4000, 2)); plane[:, 2] = 0.25
    _assert_voxels_equal(_gpu(plane), oracle.NdtTarget(plane, oracle.ndt_params()))
    big = _finite(np.concatenate(_scans(seed=5, beams=64, az=2048, n=3)))  # > 262 k points: several sort tiles
    _assert_voxels_equal(_gpu(big), oracle.NdtTarget(big, oracle.ndt_params()))


@pytest.mark.parametrize("method", [0, 1, 2, 3])
def test_derivatives_match_oracle(oracle, method):
    """One computeDerivatives / computeHessian pass at a fixed pose.  The per-pair terms are the oracle's (float32, same
    order); the sums differ by the association of a parallel reduction (and by the rare argument where expf and a rounded
    double exp disagree): 1e-9 relative for the float pass, 1e-12 for the double Hessian pass."""
    s0, s1 = _scans()
    T = oracle.NdtTarget(s0, oracle.ndt_params(num_threads=8, search_method=method))
    n = _gpu(s0, method=method)
    n.setInputSource(s1)
    for pose in (np.zeros(6), np.array([0.05, -0.02, 0.01, 0.01, -0.02, 0.015]), np.array([-0.3, 0.2, 0.05, -0.02, 0.03, -0.04])):
        M = oracle.ndt_pose_to_matrix(pose)
        for ch in (1, 0):
            sc, g, H = n.derivatives(M, pose, ch)
            osc, og, oH = T.derivatives(s1, M, pose, bool(ch))
            assert abs(sc - osc) <= 1e-9 * abs(osc), (sc, osc)
            assert np.abs(g - og).max() <= 1e-9 * np.abs(og).max()
            assert np.abs(H - oH).max() <= 1e-9 * max(np.abs(oH).max(), 1e-300)
        _, _, Hd = n.derivatives(M, pose, 2)
        oHd = T.hessian(s1, M, pose)
        assert np.abs(Hd - oHd).max() <= 1e-12 * np.abs(oHd).max()


def _check_align(oracle, n, T, src, guess=None, tag=""):
    n.setInputSource(src)
    r = n.align(guess)
    o = T.align(src, guess=guess)
    dt, dr = F.pose_delta(o["T"], n.getFinalTransformation())
    assert dt <= TOL_T and dr <= TOL_R, (tag, dt, dr, r.nr_iterations, o["iterations"])
    assert r.nr_iterations == o["iterations"] and r.n_evaluations == o["evaluations"] and bool(r.converged) == o["converged"], \
        (tag, r.nr_iterations, o["iterations"], r.n_evaluations, o["evaluations"])
    assert abs(r.trans_probability - o["trans_probability"]) <= 1e-6 * abs(o["trans_probability"])
    assert np.abs(np.array(r.pose) - o["pose"]).max() < 1e-6
    return dt, dr


@pytest.mark.parametrize("method", [0, 1, 2, 3])
@pytest.mark.parametrize("eps", [0.1, 0.01, 1e-3])
def test_align_matches_oracle(oracle, method, eps):
    """align() on a scan pair (the reference's default epsilon 0.1, LOCUS-like 1e-2 / 1e-3), with and without a guess:
    pose within 1e-4 m / 1e-4 rad of the oracle, the same number of Newton steps and evaluations."""
    s0, s1 = _scans()
    T = oracle.NdtTarget(s0, oracle.ndt_params(num_threads=8, transformation_epsilon=eps, search_method=method))
    n = _gpu(s0, method=method, eps=eps)
    _check_align(oracle, n, T, s1, tag="pair")
    _check_align(oracle, n, T, s1, guess=oracle.ndt_pose_to_matrix(np.array([-0.1, -0.05, 0.0, 0.0, 0.01, 0.01])), tag="guess")
    _check_align(oracle, n, T, s1, guess=oracle.ndt_pose_to_matrix(np.array([0.1, 0.1, 0.0, -0.02, 0.0, -0.01])), tag="guess, negative roll")


def test_align_known_offset_and_determinism(oracle):
    """A moved copy of the target comes back to the known transform (tight epsilon); two aligns give identical bits;
    changing the resolution rebuilds the voxel structure of the stored target (setResolution, ndt_omp.h:124-131)."""
    s0, _ = _scans()
    pose = np.array([0.15, -0.1, 0.02, 0.01, -0.015, 0.02])
    M = oracle.ndt_pose_to_matrix(pose).astype(np.float64); Mi = np.linalg.inv(M)
    s1 = (s0.astype(np.float64) @ Mi[:3, :3].T + Mi[:3, 3]).astype(np.float32)
    n = _gpu(s0, eps=1e-3)
    n.setInputSource(s1)
    r1 = n.align(); T1 = n.getFinalTransformation().copy()
    assert r1.converged and np.abs(np.array(r1.pose)[:3] - pose[:3]).max() < 2e-3 and np.abs(np.array(r1.pose)[3:] - pose[3:]).max() < 5e-4
    n.align()
    assert np.array_equal(T1, n.getFinalTransformation())
    n.setResolution(2.0)
    T2 = oracle.NdtTarget(s0, oracle.ndt_params(resolution=2.0, transformation_epsilon=1e-3, num_threads=8))
    _assert_voxels_equal(n, T2)
    _check_align(oracle, n, T2, s1, tag="resolution 2")


def test_full_size_scan_to_scan(oracle):
    """BASELINE config-2 shapes: a ~30 k-point filtered scan against the previous 131 072-ray scan, LOCUS's epsilon 1e-3."""
    import locus_b200
    scene = G.make_scene(11)
    poses = G.trajectory(3, 11)
    raw = [G.scan(scene, poses[i], 70 + i, beams=64, az=2048) for i in range(2)]
    tgt = _finite(raw[0].view(np.float32).reshape(-1, 8)[:, :3].copy())
    vg = locus_b200.VoxelGridB200(0)
    vg.setLeafSize(0.13)
    f = vg.filter(raw[1], 32, locus_b200.xyzi_fields())
    src = _finite(np.ascontiguousarray(f).view(np.float32).reshape(-1, 8)[:, :3].copy())
    assert 15000 < len(src) < 60000 and len(tgt) > 100000
    T = oracle.NdtTarget(tgt, oracle.ndt_params(num_threads=16, transformation_epsilon=1e-3))
    n = _gpu(tgt, eps=1e-3)
    _assert_voxels_equal(n, T)
    _check_align(oracle, n, T, src, tag="c2")
    dt, dr = F.pose_delta(np.linalg.inv(poses[0]) @ poses[1], n.getFinalTransformation())
    assert dt < 0.2 and dr < 0.02                   # NDT with 1 m voxels: coarse, but it is the scan's motion


def test_error_semantics():
    import locus_b200
    s0, s1 = _scans()
    n = locus_b200.NdtB200(0)
    n.setInputSource(s1)
    with pytest.raises(locus_b200.LocusB200Error) as e:
        n.align()
    assert e.value.status == -5                     # LB_ERR_NO_TARGET
    n.setInputTarget(s0)
    r0 = n.align(); T0 = n.getFinalTransformation().copy()
    with pytest.raises(locus_b200.LocusB200Error) as e:
        n.setInputSource(np.zeros((0, 3), np.float32))
    assert e.value.status == -4                     # LB_ERR_EMPTY_SOURCE, previous source kept
    bad = s1.copy(); bad[7, 2] = np.nan
    with pytest.raises(locus_b200.LocusB200Error) as e:
        n.setInputSource(bad)
    assert e.value.status == -1
    far = np.array([[0, 0, 0], [3e4, 3e4, 3e4]] * 4, np.float32)
    n.setResolution(1.0)
    with pytest.raises(locus_b200.LocusB200Error) as e:
        m = locus_b200.NdtB200(0); m.setResolution(0.05); m.setInputTarget(far)
    assert e.value.status == -7                     # LB_ERR_VOXEL_OVERFLOW
    with pytest.raises(locus_b200.LocusB200Error):
        n.setInputTarget(np.full((10, 3), np.nan, np.float32))
    n.align()                                       # refused clouds left the previous source / target in place
    assert np.array_equal(T0, n.getFinalTransformation()) and r0.nr_iterations == n.getFinalNumIteration()
    with pytest.raises(locus_b200.LocusB200Error):
        n.setNeighborhoodSearchMethod(4)            # not a pclomp::NeighborSearchMethod
    with pytest.raises(locus_b200.LocusB200Error):
        n.setMinPointPerVoxel(2)
    with pytest.raises(locus_b200.LocusB200Error):
        n.align(np.full((4, 4), np.nan, np.float32))
    assert n.launchCount() > 0


def test_pcl_ndt_shim_runs(oracle):
    """shim/b200_ndt_pcl.hpp -- the pcl::Registration subclass a `registration_method: ndt` case would instantiate --
    compiled against the PCL mock and driven through the base-class pointer like LOCUS drives icp_ (tests/shim_harness.cpp),
    on the reference's hollow-cube fixture with 0.5 m voxels: pose and Newton steps equal the oracle's"""
    import subprocess
    from test_cabi_cpu import _build_shim_harness
    r = subprocess.run([_build_shim_harness()], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    kv = dict(l.split("=", 1) for l in r.stdout.splitlines() if "=" in l)
    assert kv["ndt_converged"] == "1"
    T = np.array([float(x) for x in kv["ndt_T"].split(",")], dtype=np.float32).reshape(4, 4)
    box = F.hollow_cube(); moved = box.copy(); moved[:, 0] += np.float32(0.05); moved[:, 1] += np.float32(0.05)
    o = oracle.NdtTarget(box, oracle.ndt_params(resolution=0.5, transformation_epsilon=1e-3, max_iterations=20)).align(moved)
    dt, dr = F.pose_delta(o["T"], T)
    assert dt <= TOL_T and dr <= TOL_R and int(kv["ndt_iterations"]) == o["iterations"], (dt, dr, kv["ndt_iterations"], o["iterations"])
    assert abs(T[0, 3] + 0.05) < 1e-2 and abs(T[1, 3] + 0.05) < 1e-2
    assert float(kv["ndt_output_err"]) < 1e-6 and kv["ndt_tree_built_by_align"] == "0" and kv["ndt_source_kept"] == "1"
