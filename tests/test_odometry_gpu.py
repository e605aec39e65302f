"""lb_odometry_* (the pipelined Locus.cc:451-453 chain): results must be those of the sequential C-ABI calls
(bit-exact: same kernels, same inputs), in submission order, for every pipeline depth; errors surface per scan."""
import ctypes as C

import numpy as np
import pytest

from tools import gen_lidar as G

pytestmark = pytest.mark.gpu

STEP = 32
CFG = dict(max_iterations=30, max_optimizer_iterations=20, max_correspondence_distance=1.0,
           transformation_epsilon=1e-3, k_correspondences=20, ransac_iterations=0)


def _stream(n, beams=32, az=1024, seed=4):
    scene, poses, blobs = G.stream(seed, n, beams, az)
    return blobs


def _sequential(blobs, leaf):
    import locus_b200
    vg = locus_b200.VoxelGridB200()
    vg.setLeafSize(leaf); vg.setFilterFieldName("z"); vg.setFilterLimits(-100.0, 100.0)
    g = locus_b200.GicpB200()
    for k, v in CFG.items():
        setattr(g._p, k, v)
    g._apply()
    out, prev = [], None
    for b in blobs:
        f = vg.filter(b, STEP, locus_b200.xyzi_fields())
        rec = {"n": f.shape[0], "filtered": f.copy(), "T": None}
        if prev is not None:
            g.setInputSource(np.ascontiguousarray(f).view(np.float32).reshape(-1, 8)[:, :3].copy())
            g.setInputTarget(np.ascontiguousarray(prev).view(np.float32).reshape(-1, 8)[:, :3].copy())
            g.align()
            rec["T"] = g.getFinalTransformation().copy()
            rec["iters"] = g._res.iterations
            rec["evals"] = g._res.n_objective_evals
        out.append(rec)
        prev = f
    return out


@pytest.mark.parametrize("depth,share", [(1, False), (3, False), (1, True), (3, True)])
def test_pipeline_equals_sequential_calls(depth, share):
    import locus_b200
    blobs = _stream(7)
    leaf = 0.35
    ref = _sequential(blobs, leaf)
    n = blobs[0].size // STEP
    odo = locus_b200.OdometryB200(0, depth=depth, max_points=n, max_point_step=STEP)
    odo.setCloudSharing(share)            # share: every scan's index + covariances computed once, adopted as the next target
    odo.voxel.setLeafSize(leaf); odo.voxel.setFilterFieldName("z"); odo.voxel.setFilterLimits(-100.0, 100.0)
    odo.setGicpParams(**CFG)
    filt = [np.zeros(n * STEP, dtype=np.uint8) for _ in blobs]
    for b, fo in zip(blobs, filt):
        odo.submit(b, n, STEP, locus_b200.xyzi_fields(), filtered_out=fo)
    assert odo.pending() == len(blobs)
    for i, rec in enumerate(ref):
        r = odo.next()
        assert r.ticket == i and r.status == 0, (r.status, r.error)
        assert r.n_filtered == rec["n"]
        assert np.array_equal(filt[i][: rec["n"] * STEP].reshape(-1, STEP), rec["filtered"])
        if i == 0:
            assert r.has_pose == 0
        else:
            assert r.has_pose == 1
            T = np.array(r.gicp.final_transformation, dtype=np.float32).reshape(4, 4)
            assert np.array_equal(T, rec["T"]), (i, T - rec["T"])
            assert r.gicp.iterations == rec["iters"] and r.gicp.n_objective_evals == rec["evals"]
    assert odo.pending() == 0
    assert odo.launchCount() > 0
    odo.close()


def test_pipeline_device_buffers_and_reuse():
    """device-resident scans, more scans than ring slots and than the in-flight limit, results polled non-blocking"""
    import torch
    import locus_b200
    blobs = _stream(5, beams=16, az=1024)
    leaf = 0.4
    ref = _sequential(blobs, leaf)
    n = blobs[0].size // STEP
    d = [torch.from_numpy(b).cuda() for b in blobs]
    odo = locus_b200.OdometryB200(0, depth=2, max_points=n, max_point_step=STEP)
    odo.voxel.setLeafSize(leaf); odo.voxel.setFilterFieldName("z"); odo.voxel.setFilterLimits(-100.0, 100.0)
    odo.setGicpParams(**CFG)
    order = [0, 1, 2, 3, 4, 3, 2, 1, 0, 1, 2, 3, 4, 3, 2, 1, 0, 1, 2, 3]
    got = []
    for s in order:
        odo.submit(d[s].data_ptr(), n, STEP, locus_b200.xyzi_fields(), mem=locus_b200.LB_MEM_DEVICE)
        r = odo.next(block=False)
        if r is not None:
            got.append(r)
    while odo.pending():
        got.append(odo.next())
    assert [r.ticket for r in got] == list(range(len(order)))
    fwd = {i: ref[i]["T"] for i in range(1, 5)}
    for i, r in enumerate(got):
        assert r.status == 0
        assert r.n_filtered == ref[order[i]]["n"]
        if i and order[i] == order[i - 1] + 1:       # forward pairs are the ones the sequential run registered
            T = np.array(r.gicp.final_transformation, dtype=np.float32).reshape(4, 4)
            assert np.array_equal(T, fwd[order[i]])
    odo.close()


def test_pipeline_errors():
    import locus_b200
    from locus_b200 import api
    blobs = _stream(2, beams=16, az=512)
    n = blobs[0].size // STEP
    odo = locus_b200.OdometryB200(0, depth=2, max_points=n, max_point_step=STEP)
    odo.setGicpParams(**CFG)
    with pytest.raises(api.LocusB200Error) as e:                      # no leaf set -> stage V fails -> per-scan status
        odo.submit(blobs[0], n + 1, STEP, locus_b200.xyzi_fields())
    assert e.value.status == -8
    with pytest.raises(api.LocusB200Error):
        odo.submit(blobs[0], n, STEP, [("a", 0, api.LB_FLOAT32, 1)])
    odo.submit(blobs[0], n, STEP, locus_b200.xyzi_fields())           # leaf size never set: lb_voxel_filter refuses
    r = odo.next()
    assert r.status != 0 and r.has_pose == 0 and len(r.error) > 0
    odo.voxel.setLeafSize(0.4)
    odo.submit(blobs[0], n, STEP, locus_b200.xyzi_fields())
    odo.submit(blobs[1], n, STEP, locus_b200.xyzi_fields())
    r1, r2 = odo.next(), odo.next()
    assert r1.status == 0 and r1.has_pose == 0        # previous scan failed: nothing to register against
    assert r2.status == 0 and r2.has_pose == 1
    with pytest.raises(api.LocusB200Error) as e:
        odo.next()
    assert e.value.status == -10
    odo.close()
