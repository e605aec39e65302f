#!/usr/bin/env python
"""SURVEY 8f row f2 measurement: NormalComputation (k-NN 20) on a voxel-filtered 64-beam scan (~30 k points), the
nodelet's real input.  One step = lb_gicp_set_source (upload from a device buffer, index) + lb_gicp_compute_normals,
CUDA events around it, L2 flushed between steps; CPU arm = oracle/normals_oracle.c (PCL NormalEstimationOMP
restated) on the host cores.  One JSON line.   python tools/bench_normals.py [--steps 50]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tools import gen_lidar as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--k", type=int, default=20)
    a = ap.parse_args()
    import torch
    import locus_b200
    from oracle import oracle as O
    L = locus_b200.lib()
    scene, poses, blobs = G.stream(2, 2)
    vg = locus_b200.VoxelGridB200(0)
    vg.setLeafSize(0.10808803886175156); vg.setFilterFieldName("z"); vg.setFilterLimits(-100.0, 100.0)
    clouds = [np.ascontiguousarray(vg.filter(b, 32, locus_b200.xyzi_fields())) for b in blobs]
    d_clouds = [torch.from_numpy(c.reshape(-1)).cuda() for c in clouds]
    n = [c.shape[0] for c in clouds]
    d_out = torch.zeros(max(n), 4, dtype=torch.float32, device="cuda")
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.int32, device="cuda")
    g = locus_b200.GicpB200(0)

    def step(i):
        j = i % len(clouds)
        assert L.lb_gicp_set_source(g._h, C.c_void_p(d_clouds[j].data_ptr()), n[j], 32, 0, -1, 1) == 0
        assert L.lb_gicp_compute_normals(g._h, 0, a.k, None, C.c_void_p(d_out.data_ptr()), 1) == 0

    for w in range(a.warmup):
        step(w)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    torch.cuda.synchronize()
    for s in range(a.steps):
        flush.fill_(s)
        ev[s][0].record(); step(s); ev[s][1].record()
    torch.cuda.synchronize()
    ms = sum(x.elapsed_time(y) for x, y in ev) / a.steps
    gpu_n = d_out[: n[(a.steps - 1) % len(clouds)]].cpu().numpy()
    # CPU arm + parity on the same cloud
    j = (a.steps - 1) % len(clouds)
    xyz = clouds[j].view(np.float32).reshape(-1, 8)[:, :3].copy()
    threads = min(os.cpu_count() or 1, 64)
    O.normals_knn(xyz[:2000], a.k, num_threads=threads)
    t0 = time.perf_counter(); ref = O.normals_knn(xyz, a.k, num_threads=threads); t_cpu = time.perf_counter() - t0
    err = 1.0 - np.abs((gpu_n[:, :3].astype(np.float64) * ref[:, :3]).sum(1))
    print(json.dumps({"metric": "normal_clouds_per_sec", "value": 1e3 / ms, "unit": "clouds/s", "ms_per_cloud": ms,
                      "points": int(n[j]), "k": a.k, "steps": a.steps, "data": "synthetic",
                      "config": {"workload": "NormalComputation k-NN(%d) on a voxel-filtered 64-beam scan" % a.k,
                                 "l2": "flushed between steps"},
                      "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "clouds/s", "cores": threads, "kind": "port",
                                       "sample": "1 cloud (oracle/normals_oracle.c: PCL NormalEstimationOMP restated)"},
                      "parity": {"frac_within_1e-5": float((err < 1e-5).mean()), "max_err": float(err.max()),
                                 "curvature_max_abs_diff": float(np.abs(gpu_n[:, 3] - ref[:, 3]).max())}}))


if __name__ == "__main__":
    main()
