#!/usr/bin/env python
"""Experiment: T independent scan streams driven by T host threads on ONE GPU (each its own CUDA stream and handles).
Measures how much of the B200 a single latency-bound align() leaves idle.  Prints one JSON line per T."""
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench as B
import locus_b200
from locus_b200 import api


def worker(tid, leaf, blobs, steps, warmup, start, out):
    torch.cuda.set_device(0)
    L = locus_b200.lib()
    stream = torch.cuda.Stream()
    fields = locus_b200.xyzi_fields()
    vg = locus_b200.VoxelGridB200(0, stream=stream.cuda_stream)
    gicp = locus_b200.GicpB200(0, stream=stream.cuda_stream)
    gicp.setMaximumIterations(B.GICP_CFG["max_iterations"]); gicp.setMaximumOptimizerIterations(B.GICP_CFG["max_inner"])
    gicp.setMaxCorrespondenceDistance(B.GICP_CFG["corr_dist"]); gicp.setTransformationEpsilon(B.GICP_CFG["tf_eps"])
    gicp.setCorrespondenceRandomness(B.GICP_CFG["k"]); gicp.setRANSACIterations(0)
    vg.setFilterFieldName("z"); vg.setFilterLimits(-100.0, 100.0); vg.setLeafSize(leaf)
    nraw = blobs[0].size // B.POINT_STEP
    with torch.cuda.stream(stream):
        d_scans = [torch.from_numpy(b).cuda() for b in blobs]
        d_filt = [torch.empty(nraw * B.POINT_STEP, dtype=torch.uint8, device="cuda") for _ in range(2)]
    fa = api.VoxelGridB200._fields(fields)
    n_out = C.c_size_t(0)
    res = api.GicpResult()
    n_prev = 0
    torch.cuda.synchronize()

    def step(i):
        nonlocal n_prev
        cur, prv = d_filt[i & 1], d_filt[(i + 1) & 1]
        s = L.lb_voxel_filter(vg._h, C.c_void_p(d_scans[B.seq(i)].data_ptr()), nraw, B.POINT_STEP, fa, len(fields),
                              None, 0, C.c_void_p(cur.data_ptr()), nraw, C.byref(n_out), None, 1, 1)
        assert s == 0
        n_cur = n_out.value
        if n_prev:
            assert L.lb_gicp_set_source(gicp._h, C.c_void_p(cur.data_ptr()), n_cur, B.POINT_STEP, 0, -1, 1) == 0
            assert L.lb_gicp_set_target(gicp._h, C.c_void_p(prv.data_ptr()), n_prev, B.POINT_STEP, 0, -1, 1, None) == 0
            assert L.lb_gicp_align(gicp._h, None, C.byref(res)) == 0
        n_prev = n_cur

    for i in range(1 + warmup):
        step(i)
    start.wait()
    t0 = time.perf_counter()
    for k in range(steps):
        step(1 + warmup + k)
    torch.cuda.synchronize()
    out[tid] = (time.perf_counter() - t0, res.iterations)


def main():
    steps = int(os.environ.get("STEPS", "40"))
    leaf = float(os.environ.get("LEAF", "0.10808803886175156"))
    poses, blobs = B.make_stream(0)
    for T in [1, 2, 3, 4]:
        start = threading.Barrier(T)
        out = {}
        th = [threading.Thread(target=worker, args=(t, leaf, blobs, steps, 5, start, out)) for t in range(T)]
        [t.start() for t in th]
        [t.join() for t in th]
        wall = max(v[0] for v in out.values())
        print(json.dumps({"threads": T, "scans_per_s": T * steps / wall, "wall_s": wall, "iters": [v[1] for v in out.values()]}), flush=True)


if __name__ == "__main__":
    main()
