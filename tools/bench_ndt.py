#!/usr/bin/env python
"""SURVEY 8f row f4 measurement: NDT registration (LOCUS's `registration_method: ndt`) on BASELINE config-2 shapes --
a ~30 k-point voxel-filtered 64-beam scan against the previous 131 072-ray scan, LOCUS's epsilon 1e-3, 1 m voxels.
One step = the blocking calls of the seam: lb_ndt_set_target (voxel Gaussians of the previous scan), lb_ndt_set_source,
lb_ndt_align, with host buffers; timed with the host clock around the blocking calls (each ends with a stream
synchronisation), L2 flushed between steps.  CPU arm = oracle/ndt_oracle.c (the OpenMP NDT fork restated) on the host
cores.  One JSON line.   python tools/bench_ndt.py [--steps 20] [--profile]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tools import gen_lidar as G  # noqa: E402


def finite(a):
    return np.ascontiguousarray(a[np.isfinite(a).all(1)])


def submap(a, locus_b200, F):
    """scan-to-submap localization: the submap's voxel Gaussians are built once (timed separately) and stay resident; per scan
    lb_ndt_set_source + lb_ndt_align(prior) with host buffers."""
    vg = locus_b200.VoxelGridB200(0)

    def voxel_fn(blob, leaf):
        vg.setLeafSize(float(leaf))
        return np.ascontiguousarray(vg.filter(blob, 32, locus_b200.xyzi_fields()))

    w = G.c3_workload(10, a.scans, voxel_fn)
    sub = np.ascontiguousarray(w["submap"], dtype=np.float32)
    srcs = [finite(voxel_fn(b, 0.13).view(np.float32).reshape(-1, 8)[:, :3].copy()) for b in w["blobs"]]
    nd = locus_b200.NdtB200(0)
    nd.setTransformationEpsilon(a.eps)
    t0 = time.perf_counter(); nd.setInputTarget(sub); nd.setInputSource(srcs[0]); nd.align(w["guesses"][0]); t_first = time.perf_counter() - t0
    import torch
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.int32, device="cuda")
    tt = np.zeros(2); evals = iters = 0; poses_gpu = {}
    for k in range(a.warmup + a.steps):
        i = k % len(srcs)
        flush.fill_(k); torch.cuda.synchronize()
        c0 = time.perf_counter(); nd.setInputSource(srcs[i])
        c1 = time.perf_counter(); r = nd.align(w["guesses"][i])
        c2 = time.perf_counter()
        if k >= a.warmup:
            tt += (c1 - c0, c2 - c1); evals += r.n_evaluations; iters += r.nr_iterations
        poses_gpu[i] = nd.getFinalTransformation().copy()
    tt /= a.steps
    out = {"metric": "NDT scans/sec (scan-to-submap: ~30k-point filtered scan vs a resident %d-point submap, 1 m voxels, eps %g, prior as guess)" % (len(sub), a.eps),
           "value": 1.0 / tt.sum(), "unit": "scans/s", "ms_per_scan": 1e3 * tt.sum(), "ms_set_source": 1e3 * tt[0], "ms_align": 1e3 * tt[1],
           "ms_first_call_incl_target_build": 1e3 * t_first, "evaluations_per_align": evals / a.steps, "newton_steps_per_align": iters / a.steps,
           "source_points": int(np.mean([len(s) for s in srcs])), "target_points": int(len(sub)), "target_voxels": int(r.n_target_voxels),
           "timing": "host clock around blocking C-ABI calls, host buffers, L2 flushed between steps", "gpu_launches": int(nd.launchCount())}
    if not a.profile:
        from oracle import oracle as O
        threads = min(os.cpu_count() or 1, 64)
        c0 = time.perf_counter(); T = O.NdtTarget(sub, O.ndt_params(num_threads=threads, transformation_epsilon=a.eps)); c1 = time.perf_counter()
        ct = 0.0; worst = (0.0, 0.0); truth = (0.0, 0.0)
        for i in range(len(srcs)):
            c2 = time.perf_counter(); o = T.align(srcs[i], guess=w["guesses"][i]); ct += time.perf_counter() - c2
            d = F.pose_delta(o["T"], poses_gpu[i]); worst = (max(worst[0], float(d[0])), max(worst[1], float(d[1])))
            e = F.pose_delta(w["poses"][i], poses_gpu[i]); truth = (max(truth[0], float(e[0])), max(truth[1], float(e[1])))
        ct /= len(srcs)
        out["cpu_baseline"] = {"value": 1.0 / ct, "unit": "scans/s", "cores": threads, "kind": "port",
                               "sample": "%d scans against the kept submap (its voxel Gaussians built once, %.0f ms, untimed): align %.1f ms" % (len(srcs), 1e3 * (c1 - c0), 1e3 * ct)}
        out["speedup_vs_cpu"] = out["value"] / out["cpu_baseline"]["value"]
        out["pose_delta_vs_cpu"] = {"scans": len(srcs), "max_m": worst[0], "max_rad": worst[1]}
        out["pose_error_vs_truth"] = {"max_m": truth[0], "max_rad": truth[1]}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scans", type=int, default=6)
    ap.add_argument("--eps", type=float, default=1e-3)
    ap.add_argument("--profile", action="store_true", help="GPU arm only, no CPU arm (for ncu)")
    ap.add_argument("--submap", action="store_true", help="BASELINE config-3 shape: filtered scans vs a resident 500 k-point submap, with the prior as guess")
    a = ap.parse_args()
    import locus_b200
    import fixtures as F
    if a.submap:
        return submap(a, locus_b200, F)
    scene = G.make_scene(11)
    poses = G.trajectory(a.scans, 11)
    raw = [G.scan(scene, poses[i], 70 + i, beams=64, az=2048) for i in range(a.scans)]
    vg = locus_b200.VoxelGridB200(0)
    vg.setLeafSize(0.13)
    tgts = [finite(r.view(np.float32).reshape(-1, 8)[:, :3].copy()) for r in raw]
    srcs = [finite(np.ascontiguousarray(vg.filter(r, 32, locus_b200.xyzi_fields())).view(np.float32).reshape(-1, 8)[:, :3].copy()) for r in raw]
    pairs = [(i - 1, i) for i in range(1, a.scans)]
    nd = locus_b200.NdtB200(0)
    nd.setTransformationEpsilon(a.eps)
    flush = None
    if not a.profile:
        import torch
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.int32, device="cuda")

    def step(k):
        t, s = pairs[k % len(pairs)]
        t0 = time.perf_counter(); nd.setInputTarget(tgts[t])
        t1 = time.perf_counter(); nd.setInputSource(srcs[s])
        t2 = time.perf_counter(); r = nd.align()
        t3 = time.perf_counter()
        return (t1 - t0, t2 - t1, t3 - t2), r

    for w in range(a.warmup):
        step(w)
    tt = np.zeros(3); evals = 0; iters = 0; poses_gpu = {}
    for k in range(a.steps):
        if flush is not None:
            import torch
            flush.fill_(k); torch.cuda.synchronize()
        dt, r = step(k)
        tt += dt; evals += r.n_evaluations; iters += r.nr_iterations
        poses_gpu[k % len(pairs)] = nd.getFinalTransformation().copy()
    tt /= a.steps
    out = {"metric": "NDT scans/sec (scan-to-scan, 131k-ray target, ~30k-point source, 1 m voxels, eps %g)" % a.eps,
           "value": 1.0 / tt.sum(), "unit": "scans/s", "ms_per_scan": 1e3 * tt.sum(),
           "ms_set_target": 1e3 * tt[0], "ms_set_source": 1e3 * tt[1], "ms_align": 1e3 * tt[2],
           "evaluations_per_align": evals / a.steps, "newton_steps_per_align": iters / a.steps,
           "ms_per_evaluation": 1e3 * tt[2] / max(evals / a.steps, 1e-9),
           "source_points": int(np.mean([len(s) for s in srcs[1:]])), "target_points": int(np.mean([len(t) for t in tgts[:-1]])),
           "target_voxels": int(r.n_target_voxels), "timing": "host clock around blocking C-ABI calls, host buffers, L2 flushed between steps",
           "gpu_launches": int(nd.launchCount())}
    if not a.profile:
        from oracle import oracle as O
        threads = min(os.cpu_count() or 1, 64)
        prm = O.ndt_params(num_threads=threads, transformation_epsilon=a.eps)
        ct = np.zeros(2); worst = (0.0, 0.0); n_cmp = 0
        for k, (t, s) in enumerate(pairs):
            c0 = time.perf_counter(); T = O.NdtTarget(tgts[t], prm)
            c1 = time.perf_counter(); o = T.align(srcs[s])
            c2 = time.perf_counter()
            ct += (c1 - c0, c2 - c1)
            if k in poses_gpu:
                d = F.pose_delta(o["T"], poses_gpu[k]); worst = (max(worst[0], float(d[0])), max(worst[1], float(d[1]))); n_cmp += 1
        ct /= len(pairs)
        out["cpu_baseline"] = {"value": 1.0 / ct.sum(), "unit": "scans/s", "cores": threads, "kind": "port",
                               "sample": "%d scan pairs, target build %.1f ms + align %.1f ms" % (len(pairs), 1e3 * ct[0], 1e3 * ct[1])}
        out["speedup_vs_cpu"] = out["value"] / out["cpu_baseline"]["value"]
        out["pose_delta_vs_cpu"] = {"pairs": n_cmp, "max_m": worst[0], "max_rad": worst[1]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
