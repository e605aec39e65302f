"""lb_gicp_compute_normals (point_cloud_filter::NormalComputation, k-NN mode) vs the oracle's
restatement of pcl::NormalEstimationOMP.  The neighbour sets and the float32 accumulators are bit-identical on both
sides (tests/test_hd_cpu.py pins the shared arithmetic on the CPU); the only platform difference is the last bit of
atan2f / cosf / sinf inside PCL's closed-form eigen-solver, which moves a well-conditioned normal by ~1e-7 rad and
can pick another vector only where two eigenvalues coincide.  Bars: >= 99.5 % of the normals within 1e-5 of the
oracle's (as 1 - |cos|), every normal unit length and flipped towards the viewpoint, curvature within 1e-5."""
import numpy as np
import pytest

import fixtures as F
from tools import gen_lidar as G

pytestmark = pytest.mark.gpu


def _compare(pts, k, vp, which=0):
    import locus_b200
    from oracle import oracle as O
    g = locus_b200.GicpB200()
    if which == 0:
        g.setInputSource(pts)
    else:
        g.setInputTarget(pts)
    out = g.computeNormals(which=which, k=k, viewpoint=vp)
    ref = O.normals_knn(pts, k, viewpoint=(0, 0, 0) if vp is None else vp)
    assert out.shape == ref.shape == (len(pts), 4)
    n_gpu = out[:, :3].astype(np.float64); n_ref = ref[:, :3].astype(np.float64)
    assert np.isfinite(n_gpu).all()
    assert np.abs(np.linalg.norm(n_gpu, axis=1) - 1).max() < 1e-5
    v = (np.zeros(3) if vp is None else np.asarray(vp, np.float64)) - pts[:, :3].astype(np.float64)
    assert ((v * n_gpu).sum(1) >= -1e-5 * np.linalg.norm(v, axis=1)).all()          # flipped towards the viewpoint
    err = 1.0 - np.abs((n_gpu * n_ref).sum(1))
    assert (err < 1e-5).mean() >= 0.995, (err < 1e-5).mean()
    assert (np.abs(out[:, 3] - ref[:, 3]) < 1e-5).mean() >= 0.995
    same_sign = ((n_gpu * n_ref).sum(1) > 0) | (np.abs((v * n_ref).sum(1)) < 1e-4 * np.linalg.norm(v, axis=1))
    assert same_sign[err < 1e-5].all()
    return out, ref


@pytest.mark.parametrize("k", [5, 20])
def test_normals_scene(k):
    pts = F.random_scene(20000, 3)
    _compare(pts, k, None)
    _compare(pts, k, np.array([1.0, -2.0, 0.5], np.float32), which=1)


def test_normals_filtered_lidar_scan():
    """the nodelet's real input: a voxel-filtered 64-beam scan (sparse far range -> the tail kernel path)"""
    import locus_b200
    scene, poses, blobs = G.stream(2, 1)
    vg = locus_b200.VoxelGridB200()
    vg.setLeafSize(0.1088); vg.setFilterFieldName("z"); vg.setFilterLimits(-100.0, 100.0)
    f = vg.filter(blobs[0], 32, locus_b200.xyzi_fields())
    pts = np.ascontiguousarray(f).view(np.float32).reshape(-1, 8)[:, :3].copy()
    _compare(pts, 20, None)


def test_normals_plane_known_answer_and_device_output():
    import torch
    import locus_b200
    rng = np.random.default_rng(0)
    xy = rng.uniform(-5, 5, (6000, 2))
    z = 0.3 * xy[:, 0] + 0.1 * xy[:, 1] + 5.0 + rng.normal(0, 0.002, 6000)
    pts = np.c_[xy, z].astype(np.float32)
    out, _ = _compare(pts, 20, None)
    nref = np.array([-0.3, -0.1, 1.0]); nref /= np.linalg.norm(nref)
    assert (np.abs(out[:, :3] @ nref) > 1 - 1e-3).all() and (out[:, 2] < 0).all()    # sensor at the origin, plane above it
    assert out[:, 3].max() < 1e-2
    # device-resident output buffer
    g = locus_b200.GicpB200()
    g.setInputSource(pts)
    d = torch.zeros(len(pts), 4, dtype=torch.float32, device="cuda")
    L = locus_b200.lib()
    import ctypes as C
    assert L.lb_gicp_compute_normals(g._h, 0, 20, None, C.c_void_p(d.data_ptr()), locus_b200.LB_MEM_DEVICE) == 0
    assert np.array_equal(d.cpu().numpy(), out)


def test_normals_errors():
    import locus_b200
    from locus_b200 import api
    g = locus_b200.GicpB200()
    with pytest.raises(api.LocusB200Error):
        g.computeNormals(which=1, k=20)                    # no target cloud
    pts = F.random_scene(500, 1)
    g.setInputSource(pts[:10])
    for k, status in ((2, -9), (21, -9), (12, -6)):
        with pytest.raises(api.LocusB200Error) as e:
            g.computeNormals(which=0, k=k)
        assert e.value.status == status


@pytest.mark.parametrize("radius", [0.3, 0.6])
def test_normals_radius_mode_and_nan_removal(radius):
    """the nodelet's radius mode (normal_computation.cc:73-77) + removeNaNNormalsFromPointCloud (:53-57): the same NaN
    rows and the same kept indices as the oracle (neighbour sets are exact: d2 < float(radius^2), ascending distance
    order), normals within 1e-5 where defined"""
    import locus_b200
    from oracle import oracle as O
    pts = F.random_scene(12000, 5)
    g = locus_b200.GicpB200()
    g.setInputSource(pts)
    vp = np.array([0.5, -1.0, 2.0], np.float32)
    out, keep = g.computeNormalsRadius(which=0, radius=radius, viewpoint=vp)
    ref, rkeep = O.normals_radius(pts, radius, viewpoint=vp)
    assert np.array_equal(keep, rkeep)                                   # the nodelet's output cloud: same points, same order
    assert np.array_equal(np.isnan(out[:, 0]), np.isnan(ref[:, 0])) and 0 < len(keep) < len(pts)
    n_gpu = out[keep, :3].astype(np.float64); n_ref = ref[keep, :3].astype(np.float64)
    err = 1.0 - np.abs((n_gpu * n_ref).sum(1))
    assert (err < 1e-5).mean() >= 0.995 and np.abs(np.linalg.norm(n_gpu, axis=1) - 1).max() < 1e-5
    assert (np.abs(out[keep, 3] - ref[keep, 3]) < 1e-5).mean() >= 0.995
    # the filtered lidar scan the nodelet really sees, as target cloud
    import locus_b200 as lb
    scene, poses, blobs = G.stream(2, 1)
    vg = lb.VoxelGridB200()
    vg.setLeafSize(0.25); vg.setFilterFieldName("z"); vg.setFilterLimits(-100.0, 100.0)
    f = np.ascontiguousarray(vg.filter(blobs[0], 32, lb.xyzi_fields())).view(np.float32).reshape(-1, 8)[:, :3].copy()
    g.setInputTarget(f)
    out, keep = g.computeNormalsRadius(which=1, radius=radius)
    ref, rkeep = O.normals_radius(f, radius)
    assert np.array_equal(keep, rkeep)
    err = 1.0 - np.abs((out[keep, :3].astype(np.float64) * ref[keep, :3].astype(np.float64)).sum(1))
    assert (err < 1e-5).mean() >= 0.99
