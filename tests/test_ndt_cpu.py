"""SURVEY 8f row f4 (NDT) without a GPU: sanity of the oracle restatement (the reference holds no NDT test, so it is
PARITY UNPINNED -- oracle/ndt_oracle.c header) and the product's host+device NDT logic (locus_b200/csrc/ndt.h, compiled
by g++ into tests/ndt_harness.cpp with a serial backend) against that oracle."""
import ctypes as C

import numpy as np
import pytest

import fixtures as F
from tools import gen_lidar as G


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _scans(seed=3, beams=32, az=1024):
    scene = G.make_scene(seed)
    poses = G.trajectory(3, seed)
    s = [G.scan(scene, poses[i], 5 + i, beams=beams, az=az).view(np.float32).reshape(-1, 8)[:, :3].copy() for i in range(2)]
    s = [np.ascontiguousarray(a[np.isfinite(a).all(1)]) for a in s]          # rays without a return are NaN rows
    return s[0], s[1], poses


def test_oracle_recovers_known_offset(oracle):
    """target = a scan, source = the same scan moved by a known rigid transform: with a tight epsilon the Newton /
    More-Thuente loop must come back to that transform (the reference's default epsilon 0.1 stops after two steps)."""
    s0, _, _ = _scans()
    pose = np.array([0.15, -0.1, 0.02, 0.01, -0.015, 0.02])
    M = oracle.ndt_pose_to_matrix(pose).astype(np.float64)
    Mi = np.linalg.inv(M)
    s1 = (s0.astype(np.float64) @ Mi[:3, :3].T + Mi[:3, 3]).astype(np.float32)
    for method in (0, 1, 2, 3):
        T = oracle.NdtTarget(s0, oracle.ndt_params(num_threads=4, transformation_epsilon=1e-3, search_method=method))
        r = T.align(s1)
        assert r["status"] == 0 and r["converged"]
        # DIRECT26 visits the 26 cells AROUND the point's own cell and not that cell itself (pcl::getAllNeighborCellIndices):
        # it converges to a slightly biased pose
        bar = 6e-3 if method == 1 else 2e-3
        assert np.abs(r["pose"][:3] - pose[:3]).max() < bar and np.abs(r["pose"][3:] - pose[3:]).max() < 5e-4, (method, r["pose"])
        dt, dr = F.pose_delta(r["T"], M.astype(np.float32))
        assert dt < bar and dr < 1e-3


def test_oracle_pose_matrix_round_trip(oracle):
    rng = np.random.default_rng(0)
    for _ in range(50):
        p = np.concatenate([rng.uniform(-5, 5, 3), rng.uniform(-0.6, 0.6, 3)])
        M = oracle.ndt_pose_to_matrix(p)
        R = M[:3, :3].astype(np.float64)
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(R) - 1) < 1e-6
        # Eigen 3.3's eulerAngles(0, 1, 2) returns the first angle in [0, pi]: for a negative roll it is the other
        # representation of the same rotation (roll + pi, pi - pitch, yaw + pi), which is what the reference then iterates on
        e = oracle.ndt_euler_xyz(M)
        assert 0 <= e[0] <= np.pi + 1e-6
        if p[3] >= 0:
            assert np.abs(e - p[3:]).max() < 1e-6, (p, e)
        M2 = oracle.ndt_pose_to_matrix(np.concatenate([p[:3], e.astype(np.float64)]))
        assert np.abs(M2 - M).max() < 2e-6
        assert np.array_equal(M[:3, 3], p[:3].astype(np.float32))


def test_oracle_voxels_vs_numpy(oracle):
    """The voxel Gaussians against a plain numpy restatement: membership by floor(x / leaf), mean, sample covariance
    as the reference scales it, eigenvalues floored at 0.01 of the largest, inverse."""
    s0, _, _ = _scans()
    T = oracle.NdtTarget(s0, oracle.ndt_params())
    L = T.leaves()
    ijk = np.floor(s0 * np.float32(1.0)).astype(np.int64) - T.min_b
    idx = ijk[:, 0] + ijk[:, 1] * T.div_b[0] + ijk[:, 2] * T.div_b[0] * T.div_b[1]
    uniq, counts = np.unique(idx, return_counts=True)
    assert np.array_equal(uniq[counts >= 6], L["leaf_idx"]) and T.n_all == len(uniq)
    assert np.array_equal(counts[counts >= 6], np.where(L["nr_points"] > 0, L["nr_points"], counts[counts >= 6]))
    checked = 0
    for v in range(0, T.n_valid, 7):
        pts = s0[idx == L["leaf_idx"][v]].astype(np.float64)
        assert np.allclose(L["mean"][v], pts.mean(0), rtol=0, atol=1e-9)
        assert np.allclose(L["centroid"][v], pts.mean(0), atol=1e-4)
        if L["nr_points"][v] < 0:
            continue
        # voxel_grid_covariance_omp_impl.hpp:319-320: the 1/n covariance, then multiplied by (n - 1) / n (sic)
        cov = np.cov(pts.T, ddof=0) * (len(pts) - 1.0) / len(pts)
        w, E = np.linalg.eigh(cov)
        w = np.maximum(w, 0.01 * w[2])
        ref = np.linalg.inv(E @ np.diag(w) @ E.T)
        got = L["icov"][v].reshape(3, 3)
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max()), (v, got, ref)
        checked += 1
    assert checked > 20


def test_oracle_derivatives_vs_finite_differences(oracle):
    """Gradient and Hessian of the score against central differences of the score / gradient.  The score is only
    piecewise smooth (the neighbourhood of a point changes), so the bar is a few percent of the largest entry."""
    s0, s1, _ = _scans()
    T = oracle.NdtTarget(s0, oracle.ndt_params(num_threads=4))
    p0 = np.array([0.02, -0.01, 0.005, 0.004, -0.006, 0.008])
    sc, g, H = T.derivatives(s1, oracle.ndt_pose_to_matrix(p0), p0)
    eps = 2e-4
    gn = np.zeros(6); Hn = np.zeros((6, 6))
    for k in range(6):
        pp = p0.copy(); pp[k] += eps
        pm = p0.copy(); pm[k] -= eps
        sp, gp, _ = T.derivatives(s1, oracle.ndt_pose_to_matrix(pp), pp, False)
        sm, gm, _ = T.derivatives(s1, oracle.ndt_pose_to_matrix(pm), pm, False)
        gn[k] = (sp - sm) / (2 * eps); Hn[k] = (gp - gm) / (2 * eps)
    assert np.abs(gn - g).max() < 0.03 * np.abs(g).max(), (g, gn)
    assert np.abs(Hn - H).max() < 0.05 * np.abs(H).max()
    Hd = T.hessian(s1, oracle.ndt_pose_to_matrix(p0), p0)          # the double-precision pass the line search ends with
    assert np.abs(Hd - H).max() < 1e-5 * np.abs(H).max()


class _HT:
    """ndt.h through the serial harness."""

    def __init__(self, H, tgt, resolution=1.0, min_pts=6, eig_mult=0.01, method=0, outlier_ratio=0.55):
        self.H = H
        self.tgt = np.ascontiguousarray(tgt, dtype=np.float32)
        self.h = H.hn_target_build(_p(self.tgt), len(self.tgt), self.tgt.shape[1], resolution, min_pts, eig_mult, method, outlier_ratio)
        nv = C.c_int(); self.min_b = np.zeros(3, np.int32); self.div_b = np.zeros(3, np.int32)
        self.status = H.hn_target_info(self.h, C.byref(nv), _p(self.min_b), _p(self.div_b))
        self.n_valid = nv.value

    def __del__(self):
        self.H.hn_target_free(self.h)

    def leaves(self):
        n = self.n_valid
        o = {"leaf_idx": np.zeros(n, np.int32), "nr_points": np.zeros(n, np.int32), "mean": np.zeros((n, 3)),
             "icov": np.zeros((n, 9)), "centroid": np.zeros((n, 3), np.float32)}
        self.H.hn_target_leaves(self.h, _p(o["leaf_idx"]), _p(o["nr_points"]), _p(o["mean"]), _p(o["icov"]), _p(o["centroid"]))
        return o

    def eval(self, src, T, pose, want):
        src = np.ascontiguousarray(src, dtype=np.float32)
        T = np.ascontiguousarray(T, dtype=np.float32).reshape(16); pose = np.ascontiguousarray(pose, dtype=np.float64)
        sums = np.zeros(43)
        self.H.hn_eval(self.h, _p(src), len(src), src.shape[1], _p(T), _p(pose), want, _p(sums))
        return sums[0], sums[1:7].copy(), sums[7:].reshape(6, 6).copy()

    def align(self, src, guess=None, step=0.1, eps=0.1, maxit=35):
        src = np.ascontiguousarray(src, dtype=np.float32)
        g = None if guess is None else np.ascontiguousarray(guess, dtype=np.float32).reshape(16)
        fin = np.zeros(16, np.float32); cv = C.c_int(); it = C.c_int(); ev = C.c_int(); ps = np.zeros(6); tp = C.c_double()
        self.H.hn_align(self.h, _p(src), len(src), src.shape[1], _p(g), step, eps, maxit, _p(fin), C.byref(cv), C.byref(it), C.byref(ev),
                        _p(ps), C.byref(tp))
        return {"T": fin.reshape(4, 4), "converged": bool(cv.value), "iterations": it.value, "evaluations": ev.value, "pose": ps,
                "trans_probability": tp.value}


@pytest.mark.parametrize("method", [0, 1, 2, 3])
def test_product_headers_match_oracle(ndt_harness, oracle, method):
    """ndt.h (what the CUDA kernels run per voxel / per point / in the controller thread) with a serial backend:
    voxel Gaussians and the double-precision Hessian pass bit-identical to the oracle; the float pass to ~1e-10 (the
    oracle calls expf, ndt.h rounds a double exp -- the same float in all but rare arguments); align(): the same float
    transform bit for bit, the same number of Newton steps and evaluations."""
    s0, s1, _ = _scans()
    prm = oracle.ndt_params(num_threads=4, transformation_epsilon=0.01, search_method=method)
    T = oracle.NdtTarget(s0, prm)
    h = _HT(ndt_harness, s0, method=method)
    assert h.status == 0 and h.n_valid == T.n_valid and np.array_equal(h.min_b, T.min_b) and np.array_equal(h.div_b, T.div_b)
    a, b = h.leaves(), T.leaves()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    pose = np.array([0.05, -0.02, 0.01, 0.01, -0.02, 0.015]); M = oracle.ndt_pose_to_matrix(pose)
    for want in (1, 2):
        sc, g, Hm = h.eval(s1, M, pose, want)
        osc, og, oH = T.derivatives(s1, M, pose, want == 1)
        assert abs(sc - osc) <= 1e-9 * abs(osc) and np.abs(g - og).max() <= 1e-9 * np.abs(og).max()
        assert np.abs(Hm - oH).max() <= 1e-9 * max(np.abs(oH).max(), 1e-300)
    _, _, Hd = h.eval(s1, M, pose, 3)
    assert np.array_equal(Hd, T.hessian(s1, M, pose))
    for guess in (None, oracle.ndt_pose_to_matrix(np.array([-0.1, -0.05, 0.0, 0.0, 0.01, 0.01]))):
        r = h.align(s1, guess=guess, eps=0.01)
        o = T.align(s1, guess=guess)
        assert r["iterations"] == o["iterations"] and r["evaluations"] == o["evaluations"] and r["converged"] == o["converged"]
        dt, dr = F.pose_delta(o["T"], r["T"])
        assert dt < 1e-6 and dr < 1e-6, (dt, dr)
        assert np.abs(r["pose"] - o["pose"]).max() < 1e-8 and abs(r["trans_probability"] - o["trans_probability"]) < 1e-9


def test_product_headers_edge_cases(ndt_harness, oracle):
    """Non-finite target points are skipped; a planar target (singular covariances -> eigenvalue floor); a target whose
    voxel index would overflow is refused by both; default epsilon 0.1 (the reference's loose stop)."""
    s0, s1, _ = _scans()
    bad = s0.copy(); bad[::97, 1] = np.nan; bad[5, 0] = np.inf
    T = oracle.NdtTarget(bad, oracle.ndt_params()); h = _HT(ndt_harness, bad)
    a, b = h.leaves(), T.leaves()
    assert h.n_valid == T.n_valid and all(np.array_equal(a[k], b[k]) for k in a)
    rng = np.random.default_rng(1)
    plane = np.zeros((4000, 3), np.float32); plane[:, :2] = rng.uniform(-10, 10, (4000, 2)); plane[:, 2] = 0.25
    T = oracle.NdtTarget(plane, oracle.ndt_params()); h = _HT(ndt_harness, plane)
    a, b = h.leaves(), T.leaves()
    assert h.n_valid == T.n_valid > 50 and all(np.array_equal(a[k], b[k]) for k in a)
    assert np.isfinite(b["icov"]).all() or (b["nr_points"] < 0).any()
    far = np.array([[0, 0, 0], [3e4, 3e4, 3e4]] * 4, np.float32)      # 6e5 voxels per axis: the int32 index would overflow
    T = oracle.NdtTarget(far, oracle.ndt_params(resolution=0.05)); h = _HT(ndt_harness, far, resolution=0.05)
    assert T.status == -2 and h.status == -2
    T = oracle.NdtTarget(s0, oracle.ndt_params(num_threads=4)); h = _HT(ndt_harness, s0)
    r, o = h.align(s1), T.align(s1)
    dt, dr = F.pose_delta(o["T"], r["T"])
    assert r["iterations"] == o["iterations"] and r["evaluations"] == o["evaluations"] and dt < 1e-6 and dr < 1e-6


@pytest.mark.parametrize("resolution", [1.0, 0.4, 2.5])
def test_neighbourhood_search_exact_vs_kdtree(ndt_harness, oracle, resolution):
    """The reference answers `radiusSearch(x, resolution)` from a kd-tree over the voxel centroids; ndt.h answers it from the
    lattice cells around x (ndt_kd_range: conservative range, ndt_kd_probe: hash lookup + the same float d2 < r2 test).  Same
    voxels in the same (d2, index) order for points inside, on the rim of and far outside the target -- incl. queries placed
    exactly one voxel side away from a centroid and on voxel faces."""
    s0, s1, _ = _scans()
    h = _HT(ndt_harness, s0, resolution=resolution)
    L = h.leaves()
    cen = np.ascontiguousarray(L["centroid"])
    kt = oracle.KdTree(cen)
    rng = np.random.default_rng(4)
    r = np.float32(resolution)
    edge = cen[rng.integers(0, len(cen), 400)].copy(); edge[:, 0] += r                     # d = r up to rounding: the strict test decides
    edge2 = cen[rng.integers(0, len(cen), 400)].copy(); edge2[:, 1] -= r * np.float32(0.99999994)
    faces = (np.floor(s1[:400] / r) * r).astype(np.float32)                                 # on voxel faces / corners
    far = (s1[:200] * np.float32(3.0) + np.float32(500.0)).astype(np.float32)               # outside the lattice
    q = np.ascontiguousarray(np.concatenate([s1[:3000], s0[:1000] + rng.normal(0, 0.3, (1000, 3)).astype(np.float32), edge, edge2, faces, far]), dtype=np.float32)
    counts = np.zeros(len(q), np.int32); slots = np.zeros((len(q), 32), np.int32)
    ndt_harness.hn_neighbours(h.h, _p(q), len(q), _p(counts), _p(slots))
    r2 = np.float32(np.float64(r) * np.float64(r))
    n_multi = 0
    for i in range(len(q)):
        idx, _ = kt.radius(q[i], r2)
        assert counts[i] == len(idx) and np.array_equal(slots[i, :counts[i]], idx), (i, q[i], counts[i], idx)
        n_multi += len(idx) > 1
    assert n_multi > 1000 and counts.max() <= 27


def test_oracle_pose_conventions_vs_scipy(oracle):
    """Independent cross-check of the pose <-> matrix conventions: Translation * Rx * Ry * Rz = scipy's intrinsic 'XYZ'
    Euler angles, and eulerAngles(0, 1, 2) returns them (first angle >= 0 here, where Eigen's range folding is the identity)."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(5)
    for _ in range(40):
        ang = np.array([rng.uniform(0.0, 1.2), rng.uniform(-1.2, 1.2), rng.uniform(-3.0, 3.0)])
        p = np.concatenate([rng.uniform(-3, 3, 3), ang])
        M = oracle.ndt_pose_to_matrix(p)
        assert np.abs(M[:3, :3] - R.from_euler("XYZ", ang).as_matrix()).max() < 2e-6
        assert np.abs(oracle.ndt_euler_xyz(M) - R.from_matrix(M[:3, :3].astype(np.float64)).as_euler("XYZ")).max() < 2e-6


def test_oracle_derivatives_vs_independent_score(oracle):
    """The oracle's gradient and Hessian (the reference's closed-form angle tables, float32 per-pair terms) against finite
    differences of an INDEPENDENT float64 restatement of the NDT score (Magnusson 2009 eq. 6.9-6.10) with the neighbourhoods
    frozen at the evaluation pose -- then the score is smooth and the differences are accurate.  Pins the 8 + 15 rows of the
    angle tables, the placement of the second-derivative blocks and the sign / scale of every term (measured: score 6e-8, gradient
    4e-7, Hessian 6e-5 relative -- the last one is the finite-difference error)."""
    from scipy.spatial import cKDTree
    from scipy.spatial.transform import Rotation as R
    s0, s1, _ = _scans()
    src = s1[::3].copy()
    T = oracle.NdtTarget(s0, oracle.ndt_params(num_threads=4))
    L = T.leaves()
    ok = L["nr_points"] > 0
    cen = L["centroid"].astype(np.float64); mean = L["mean"]; icov = L["icov"].reshape(-1, 3, 3)
    c1 = 10 * (1 - 0.55); c2 = 0.55 / 1.0 ** 3
    d3 = -np.log(c2); d1 = -np.log(c1 + c2) - d3; d2 = -2 * np.log((-np.log(c1 * np.exp(-0.5) + c2) - d3) / d1)
    p0 = np.array([0.03, -0.02, 0.01, 0.006, -0.009, 0.012])
    x = src.astype(np.float64)

    def transform(p):
        return x @ R.from_euler("XYZ", p[3:]).as_matrix().T + p[:3]

    q0 = (x @ oracle.ndt_pose_to_matrix(p0)[:3, :3].astype(np.float64).T + oracle.ndt_pose_to_matrix(p0)[:3, 3].astype(np.float64))
    nb = cKDTree(cen).query_ball_point(q0, 1.0 - 1e-6)
    pi = np.concatenate([np.full(len(n), i) for i, n in enumerate(nb)]).astype(np.int64)
    vi = np.concatenate([np.array(n, dtype=np.int64) for n in nb])
    keep = ok[vi]; pi, vi = pi[keep], vi[keep]
    assert len(pi) > 3 * len(src)

    def score(p):
        d = transform(p)[pi] - mean[vi]
        m = np.einsum("ni,nij,nj->n", d, icov[vi], d)
        return float(np.sum(-d1 * np.exp(-d2 * m / 2)))

    sc, g, H = T.derivatives(src, oracle.ndt_pose_to_matrix(p0), p0)
    assert abs(sc - score(p0)) < 1e-6 * abs(sc)
    e = 1e-5
    gn = np.array([(score(p0 + e * np.eye(6)[k]) - score(p0 - e * np.eye(6)[k])) / (2 * e) for k in range(6)])
    assert np.abs(gn - g).max() < 1e-5 * np.abs(g).max(), (g, gn)
    e = 2e-4
    Hn = np.zeros((6, 6))
    f0 = score(p0)
    for a in range(6):
        ea = e * np.eye(6)[a]
        Hn[a, a] = (score(p0 + ea) - 2 * f0 + score(p0 - ea)) / e ** 2
        for b in range(a + 1, 6):
            eb = e * np.eye(6)[b]
            Hn[a, b] = Hn[b, a] = (score(p0 + ea + eb) - score(p0 + ea - eb) - score(p0 - ea + eb) + score(p0 - ea - eb)) / (4 * e ** 2)
    print("ndt independent check: score rel %.2e, gradient rel %.2e, hessian rel %.2e" % (abs(sc - score(p0)) / abs(sc), np.abs(gn - g).max() / np.abs(g).max(), np.abs(Hn - H).max() / np.abs(H).max()))
    assert np.abs(Hn - H).max() < 5e-4 * np.abs(H).max(), np.abs(Hn - H).max() / np.abs(H).max()
    Hd = T.hessian(src, oracle.ndt_pose_to_matrix(p0), p0)
    assert np.abs(Hn - Hd).max() < 5e-4 * np.abs(Hd).max()


def test_newton_solve_rounds_equal_serial(ndt_harness):
    """The 6x6 Newton solve: (i) rotating the three disjoint column pairs of a round side by side (what the controller warp
    does on the device) gives the same bits as the serial sweep; (ii) the result is the (pseudo-)inverse solution -- numpy's
    lstsq -- for well-conditioned, nearly symmetric Hessian-like matrices, for ill-conditioned ones (1e10) and for a
    rank-deficient one (Eigen's rank threshold drops the null direction)."""
    rng = np.random.default_rng(9)
    xs = np.zeros(6); xr = np.zeros(6)
    for trial in range(200):
        B = rng.normal(0, 1, (6, 6))
        A = B @ B.T * 10.0 ** rng.uniform(-2, 6) + rng.normal(0, 1e-7, (6, 6))        # nearly symmetric, like the float Hessian
        if trial % 4 == 1:
            Uq, _ = np.linalg.qr(B); A = Uq @ np.diag(10.0 ** np.linspace(0, -10, 6)) @ Uq.T
        if trial % 4 == 2:
            A[:, 5] = A[:, 4]; A[5, :] = A[4, :]                                       # rank 5
        A = np.ascontiguousarray(A); b = np.ascontiguousarray(rng.normal(0, 1, 6))
        ndt_harness.hn_svd6_serial(_p(A), _p(b), _p(xs)); ndt_harness.hn_svd6_rounds(_p(A), _p(b), _p(xr))
        assert np.array_equal(xs, xr), trial
        ref = np.linalg.lstsq(A, b, rcond=6 * np.finfo(float).eps)[0]
        cond = np.linalg.cond(A) if trial % 4 != 2 else 1e6
        assert np.abs(xs - ref).max() <= 1e-9 * max(cond, 1e3) * max(np.abs(ref).max(), 1e-300) / 1e3, (trial, xs, ref)


@pytest.mark.parametrize("method", [0, 1, 2, 3])
def test_grouped_evaluation_equals_per_point(ndt_harness, method):
    """The organisation of ndt_eval_group_kernel (eight lanes per point: cells probed in lane order, hits ranked by (key, slot),
    pairs in rounds of eight, every sum accumulated over the pairs in rank order) emulated serially with the helpers the kernel
    calls: bit-identical sums to the one-thread-per-point form, for both float passes -- the discovery order of the hits does
    not matter, and the float-reciprocal cell decode is exact."""
    s0, s1, _ = _scans()
    h = _HT(ndt_harness, s0, method=method)
    src = np.ascontiguousarray(s1[:6000], dtype=np.float32)
    pose = np.array([0.05, -0.02, 0.01, 0.01, -0.02, 0.015])
    from oracle import oracle as O
    M = np.ascontiguousarray(O.ndt_pose_to_matrix(pose).reshape(16))
    for want in (1, 2):
        a = np.zeros(43); b = np.zeros(43)
        ndt_harness.hn_eval(h.h, _p(src), len(src), 3, _p(M), _p(pose), want, _p(a))
        ndt_harness.hn_eval_grouped(h.h, _p(src), len(src), 3, _p(M), _p(pose), want, _p(b))
        assert np.isfinite(b).all() and np.array_equal(a, b), (method, want, np.abs(a - b).max())
        assert a[0] != 0
