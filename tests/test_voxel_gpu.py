"""VoxelGrid (K1) on the GPU vs the oracle, through the C ABI.  Bar (north_star): voxel counts and
indices bit-exact, centroids within 1 ULP (we assert bit-exact: same float32 summation order)."""
import numpy as np
import pytest

from tools import gen_lidar as G

pytestmark = pytest.mark.gpu

FIELDS = None


def _fields():
    import locus_b200
    return locus_b200.xyzi_fields()


def _ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a); b = np.where(b < 0, -(b & 0x7fffffff), b)
    return np.abs(a - b)


def _run(blob, leaf, ff="z", lo=-100.0, hi=100.0, neg=False, min_pts=0, all_data=True, body=None):
    import locus_b200
    from oracle import oracle as O
    vg = locus_b200.VoxelGridB200()
    if np.isscalar(leaf):
        leaf = (leaf, leaf, leaf)
    vg.setLeafSize(*leaf)
    if ff:
        vg.setFilterFieldName(ff); vg.setFilterLimits(lo, hi); vg.setFilterLimitsNegative(neg)
    vg.setMinimumPointsNumberPerVoxel(min_pts)
    vg.setDownsampleAllData(all_data)
    if body is not None:
        vg.setBodyFilter(*body)
    out, vidx = vg.filter(blob, 32, _fields(), want_voxel_idx=True)
    ffo = {"x": 0, "y": 4, "z": 8, "intensity": 16}.get(ff, -1) if ff else -1
    r = O.voxel_filter(blob, 32, leaf, float_fields=G.FLOAT_FIELDS, filter_field_offset=ffo, limit_min=lo,
                       limit_max=hi, negative=neg, min_points_per_voxel=min_pts, downsample_all_data=all_data, body=body)
    return out, vidx, r, vg


def _check(out, vidx, r):
    assert r["rc"] == 0
    assert out.shape == r["out"].shape, (out.shape, r["out"].shape)
    assert np.array_equal(vidx, r["voxel_idx"])
    a = out.view(np.float32).reshape(-1, 8); b = r["out"].view(np.float32).reshape(-1, 8)
    for col in (0, 1, 2, 4):
        assert _ulp_diff(a[:, col], b[:, col]).max(initial=0) <= 1
    assert np.array_equal(out, r["out"])   # in fact bit-exact, padding bytes included


@pytest.mark.parametrize("leaf", [0.25, 0.6, (0.3, 0.5, 0.2)])
def test_voxel_small_scan(leaf):
    scene = G.make_scene(3)
    blob = G.scan(scene, np.eye(4), 5, beams=16, az=512)
    out, vidx, r, _ = _run(blob, leaf)
    _check(out, vidx, r)


def test_voxel_limits_and_negative():
    scene = G.make_scene(4)
    blob = G.scan(scene, np.eye(4), 6, beams=16, az=512)
    for (lo, hi, neg) in [(-0.5, 0.4, False), (-0.5, 0.4, True), (-100, 100, False)]:
        out, vidx, r, _ = _run(blob, 0.3, lo=lo, hi=hi, neg=neg)
        _check(out, vidx, r)
    out, vidx, r, _ = _run(blob, 0.3, ff="intensity", lo=5.0, hi=50.0)
    _check(out, vidx, r)
    out, vidx, r, _ = _run(blob, 0.3, ff=None)
    _check(out, vidx, r)


def test_voxel_min_points_and_xyz_only():
    scene = G.make_scene(5)
    blob = G.scan(scene, np.eye(4), 7, beams=16, az=512)
    out, vidx, r, _ = _run(blob, 0.4, min_pts=3)
    _check(out, vidx, r)
    out, vidx, r, _ = _run(blob, 0.4, all_data=False)
    _check(out, vidx, r)


def test_voxel_edge_cases():
    import locus_b200
    vg = locus_b200.VoxelGridB200(); vg.setLeafSize(0.25)
    # empty input
    out = vg.filter(np.zeros(0, np.uint8), 32, _fields())
    assert out.shape[0] == 0
    # all NaN
    blob = np.full((100, 8), np.nan, np.float32).view(np.uint8).reshape(-1)
    assert vg.filter(blob, 32, _fields()).shape[0] == 0
    # single point, and all points in one voxel
    one = np.zeros((1, 8), np.float32); one[0, :3] = (1.5, -2.5, 0.25); one[0, 4] = 7
    o = vg.filter(one.view(np.uint8).reshape(-1), 32, _fields())
    assert o.shape[0] == 1 and np.array_equal(o.view(np.float32)[0, :3], one[0, :3])
    rng = np.random.default_rng(0)
    many = np.zeros((5000, 8), np.float32); many[:, :3] = rng.uniform(0.01, 0.24, (5000, 3)); many[:, 4] = rng.uniform(0, 9, 5000)
    out, vidx, r, _ = _run(many.view(np.uint8).reshape(-1), 0.25, ff=None)
    _check(out, vidx, r)
    assert out.shape[0] == 1
    # leaf too small -> overflow status, like PCL's warning path
    far = np.zeros((2, 8), np.float32); far[1, :3] = 4000.0
    vg.setLeafSize(0.001)
    with pytest.raises(locus_b200.LocusB200Error) as e:
        vg.filter(far.view(np.uint8).reshape(-1), 32, _fields())
    assert e.value.status == -7


def test_voxel_full_size_c2():
    """BASELINE config 2 shape: 64 x 2048 = 131072 rays -> ~30k voxels; bit-exact vs oracle plus
    size-independent properties (idempotence of the voxel set, point conservation)."""
    from oracle import oracle as O
    scene, poses, blobs = G.stream(2, 1)
    out, vidx, r, vg = _run(blobs[0], 0.1088)
    _check(out, vidx, r)
    assert 25000 < out.shape[0] < 35000
    assert int(r["count"].sum()) == int(np.isfinite(G.blob_xyz(blobs[0])).all(1).sum())
    assert np.all(np.diff(vidx) > 0)    # ascending, unique
    # filtering the output again with the same leaf keeps every voxel (idempotent voxel set)
    out2 = vg.filter(out.reshape(-1), 32, _fields())
    assert out2.shape[0] == out.shape[0]


@pytest.mark.parametrize("n_scans,leaf", [(3, 0.5), (3, 0.06), (1, 2.5)])
def test_voxel_large_inputs_and_long_segments(n_scans, leaf):
    """393 k points (above the 262 144-point limit of the scan-free radix-sort path -> the scan path) and a coarse
    leaf whose voxels hold thousands of points (the centroid kernel's multi-tile path): bit-exact vs the oracle."""
    scene = G.make_scene(4)
    poses = G.trajectory(4, n_scans, t_step=0.05, r_step_deg=0.5)
    blob = np.concatenate([G.scan(scene, poses[i], 90 + i, beams=64, az=2048) for i in range(n_scans)])
    out, vidx, r, _ = _run(blob, leaf)
    _check(out, vidx, r)
    assert int(r["count"].max()) > (2000 if leaf > 1 else 1)


def test_voxel_with_body_filter():
    """row f4: the BodyFilter nodelet (CropBox, negative, rotated about z; cfg/BodyFilter.cfg defaults scaled up so
    that the synthetic scan has points inside) folded into the voxel filter: bit-exact vs the oracle, and switching
    it off again restores the plain result"""
    scene = G.make_scene(3)
    blob = G.scan(scene, np.eye(4), 5, beams=32, az=1024)
    body = (np.array([-6.0, -3.0, -1.5], np.float32), np.array([2.5, 5.0, 0.4], np.float32), -0.785398)
    out, vidx, r, vg = _run(blob, 0.3, body=body)
    _check(out, vidx, r)
    plain, pvidx, rp, _ = _run(blob, 0.3)
    assert int(r["count"].sum()) < int(rp["count"].sum())          # something was cropped
    vg.setBodyFilter(enabled=False)
    again = vg.filter(blob, 32, _fields())
    assert np.array_equal(again, plain)


def _front_end_reference(oracle, blobs, transforms, pass_limits, leaf, body=None):
    """CPU restatement of the chain the fused call replaces (locus.launch:90-186): per lidar pcl/PassThrough (NaN
    removal, raw z limits, Eigen Matrix4f * Vector4f into base_link), point_cloud_merger's a + (b + c), then BodyFilter
    and VoxelGrid (the oracle).  numpy float32 arithmetic is IEEE-exact per operation, like the kernels' (no FMA)."""
    parts = []
    for b, T in zip(blobs, transforms):
        a = np.ascontiguousarray(b).view(np.float32).reshape(-1, 8).copy()
        x, y, z = a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()
        keep = np.isfinite(x) & np.isfinite(y) & np.isfinite(z)
        if pass_limits is not None:
            keep &= ~((z.astype(np.float64) > pass_limits[1]) | (z.astype(np.float64) < pass_limits[0]))
        if T is not None:
            T = np.asarray(T, dtype=np.float32)
            with np.errstate(invalid="ignore"):
                for r in range(3):
                    a[:, r] = ((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3]
        parts.append(a[keep])
    merged = np.ascontiguousarray(np.concatenate(parts))
    return oracle.voxel_filter(merged.view(np.uint8).reshape(-1), 32, leaf, float_fields=G.FLOAT_FIELDS, filter_field_offset=8,
                               limit_min=-100.0, limit_max=100.0, body=body)


def test_merged_inputs_passthrough_and_transform(oracle):
    """three lidars -> PassThrough (each, sensor frame) -> transform to base_link -> concatenation ->
    BodyFilter -> VoxelGrid as ONE call, bit-exact vs the chain restated on the CPU"""
    import locus_b200
    from tools import gen_lidar as G2
    scene = G2.make_scene(4)
    poses = [np.eye(4), G2.pose_matrix([0.4, 0.1, -0.2], np.deg2rad([3.0, -20.0, 45.0])), G2.pose_matrix([-0.5, 0.0, 0.1], np.deg2rad([0.0, 15.0, 180.0]))]
    blobs = [G2.scan(scene, poses[i], 70 + i, beams=16, az=1024) for i in range(3)]
    Ts = [None, poses[1].astype(np.float32), poses[2].astype(np.float32)]
    body = (np.array([-0.8, -0.5, -0.4], np.float32), np.array([0.6, 0.5, 0.3], np.float32), np.float32(0.3))
    vg = locus_b200.VoxelGridB200()
    vg.setLeafSize(0.2); vg.setFilterFieldName("z"); vg.setFilterLimits(-100.0, 100.0)
    vg.setInputPassThrough("z", -1.2, 100.0)
    vg.setBodyFilter(body[0], body[1], body[2])
    out, vidx = vg.filterMerged(blobs, 32, locus_b200.xyzi_fields(), transforms=Ts, want_voxel_idx=True)
    ref = _front_end_reference(oracle, blobs, Ts, (-1.2, 100.0), 0.2, body=body)
    assert ref["rc"] == 0 and len(ref["out"]) > 1000
    assert np.array_equal(vidx, ref["voxel_idx"])
    # x, y, z, intensity (the averaged fields) bit-exact; the padding bytes come from each voxel's first raw point
    a = np.ascontiguousarray(out).view(np.float32).reshape(-1, 8); b = np.ascontiguousarray(ref["out"]).view(np.float32).reshape(-1, 8)
    assert np.array_equal(a[:, [0, 1, 2, 4]].view(np.uint32), b[:, [0, 1, 2, 4]].view(np.uint32))
    # one input, no transform, no PassThrough == the plain call
    vg2 = locus_b200.VoxelGridB200()
    vg2.setLeafSize(0.2); vg2.setFilterFieldName("z"); vg2.setFilterLimits(-100.0, 100.0)
    assert np.array_equal(vg2.filterMerged([blobs[0]], 32, locus_b200.xyzi_fields()), vg2.filter(blobs[0], 32, locus_b200.xyzi_fields()))
    # two inputs == the concatenated blob
    cat = np.concatenate([blobs[0], blobs[1]])
    assert np.array_equal(vg2.filterMerged([blobs[0], blobs[1]], 32, locus_b200.xyzi_fields()), vg2.filter(cat, 32, locus_b200.xyzi_fields()))
