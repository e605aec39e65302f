"""the GPU-resident rolling submap (lb_submap_*) vs its CPU restatement (oracle/submap_oracle.py):
point sets after insert / crop bit-exact, neighbours exact, and the GICP pose against the resident submap (cached
per-point covariances) vs the oracle run with the same cached covariances."""
import numpy as np
import pytest

import fixtures as F

pytestmark = pytest.mark.gpu


def _clouds(seed, n, shift):
    rng = np.random.default_rng(seed)
    return (F.random_scene(n, seed) + np.float32(shift)).astype(np.float32), rng


def test_insert_crop_neighbors_match_oracle(oracle):
    import locus_b200
    from oracle.submap_oracle import SubmapOracle
    res = 0.25
    m = locus_b200.SubmapB200(0, res)
    o = SubmapOracle(res)
    a, rng = _clouds(1, 20000, 0.0)
    a[17] = np.nan                                       # non-finite points never enter the map
    b = np.concatenate([a[:5000] + np.float32(0.01), F.random_scene(15000, 2, extent=(30.0, 20.0, 3.0))]).astype(np.float32)
    n1, inc1 = m.InsertPoints(a, want_incremental=True)
    e1 = o.insert(a)
    assert n1 == len(e1) and np.array_equal(inc1, e1) and m.size() == len(o.pts)
    assert np.array_equal(m.points(), o.pts)            # insertion order: the first point of every voxel, input order
    g1 = m.generation()
    n2, inc2 = m.InsertPoints(b, want_incremental=True)
    e2 = o.insert(b)
    assert n2 == len(e2) and np.array_equal(inc2, e2) and np.array_equal(m.points(), o.pts)
    assert 0 < n2 < len(b) and m.generation() != g1
    assert m.InsertPoints(a) == 0 and m.generation() == g1 + 1       # nothing new: the map (and its index) stay
    # ApproxNearestNeighbors: exact nearest map point, ties -> lowest map index
    q = (a[rng.integers(0, len(a), 4000)] + rng.normal(0, 0.3, (4000, 3))).astype(np.float32)
    q = q[np.isfinite(q).all(axis=1)]
    nb, idx, d2 = m.ApproxNearestNeighbors(q)
    onb, oidx, od2 = o.neighbors(q)
    assert np.array_equal(idx, oidx) and np.array_equal(d2, od2) and np.array_equal(nb, onb)
    # Refresh: sliding window box crop, order kept; the occupancy follows (points can re-enter emptied voxels)
    removed = m.Refresh([2.0, -1.0, 0.0], 12.0)
    assert removed == o.crop([2.0, -1.0, 0.0], 12.0) and removed > 0
    assert np.array_equal(m.points(), o.pts)
    nb, idx, d2 = m.ApproxNearestNeighbors(q[:1000])
    onb, oidx, od2 = o.neighbors(q[:1000])
    assert np.array_equal(idx, oidx) and np.array_equal(d2, od2)
    n3 = m.InsertPoints(a)
    assert n3 == len(o.insert(a)) and n3 > 0 and np.array_equal(m.points(), o.pts)
    assert m.Refresh([0.0, 0.0, 0.0], 1000.0) == 0      # everything inside: nothing changes


def test_gicp_against_resident_submap(oracle):
    """the submap as GICP target: covariances of a map point are computed once, from the map as it was when the first
    registration after its insertion ran, and cached -- pose vs the oracle given exactly those covariances"""
    import locus_b200
    from oracle.submap_oracle import SubmapOracle
    from test_gicp_gpu import _mk, TOL_T, TOL_R
    res = 0.2
    m = locus_b200.SubmapB200(0, res)
    o = SubmapOracle(res)
    world = F.random_scene(60000, 7)
    part1, part2 = world[:35000], world[30000:]
    m.InsertPoints(part1); o.insert(part1)
    prm = oracle.default_params(transformation_epsilon=1e-4, corr_dist_threshold=0.5, max_iterations=30, num_threads=8)
    Tg = F.se3([0.06, -0.04, 0.02], [0.004, -0.003, 0.006])
    scan = world[np.random.default_rng(3).choice(len(world), 8000, replace=False)]
    src = ((scan.astype(np.float64) - Tg[:3, 3]) @ Tg[:3, :3]).astype(np.float32)
    g = _mk(prm, 0)

    def check(tag):
        g.setTargetSubmap(m)
        g.setInputSource(src)
        res_ = g.align()
        cov = o.covariances(20, 1e-3)
        assert np.allclose(g.covariances(1), cov, rtol=0, atol=1e-12), tag       # the cache, in map order
        ref = oracle.PreparedTarget(o.pts, prm, cov=cov).align(src, prm)
        dt, dr = F.pose_delta(ref["T"], g.getFinalTransformation())
        assert dt <= TOL_T and dr <= TOL_R, (tag, dt, dr)
        assert res_.iterations == ref["iterations"] and res_.n_correspondences == ref["n_corr"], tag
        return g.getFinalTransformation().copy()

    T1 = check("first")
    launches = m.launchCount()
    g.setTargetSubmap(m); g.setInputSource(src); g.align()
    assert m.launchCount() == launches                   # unchanged map: no index rebuild, no covariance kernel
    assert np.array_equal(g.getFinalTransformation(), T1)
    # keyframe: more points enter; only THEIR covariances are computed (from the grown map), the old ones are kept
    m.InsertPoints(part2); o.insert(part2)
    check("after insert")
    # sliding window: the crop removes points, the survivors keep their cached covariances
    m.Refresh([0.0, 0.0, 0.0], 18.0); o.crop([0.0, 0.0, 0.0], 18.0)
    check("after crop")
    # a covariance parameter changed: the cache is recomputed
    g.setCorrespondenceRandomness(15)
    g.setTargetSubmap(m); g.setInputSource(src); g.align()
    o.cov_key = None
    assert np.allclose(g.covariances(1), o.covariances(15, 1e-3), rtol=0, atol=1e-12)
