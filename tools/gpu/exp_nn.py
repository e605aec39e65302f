"""Correspondence search A/B (LB_NN_MODE, read once per process): align() on C2 pairs and on a C3-shaped pair in the
execution modes, results dumped for a bitwise comparison across the search modes.
usage: LB_NN_MODE=m python tools/gpu/exp_nn.py out.npz"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures as F
import locus_b200
from tools import gen_lidar as G
G.WORKERS = 4
leaf = 0.108088
scene, poses, blobs = G.stream(2, 4)
vg = locus_b200.VoxelGridB200(); vg.setLeafSize(leaf); vg.setFilterFieldName("z"); vg.setFilterLimits(-100, 100)
f = [np.ascontiguousarray(vg.filter(b, 32, locus_b200.xyzi_fields())).view(np.float32).reshape(-1, 8)[:, :3].copy() for b in blobs]
cases = [("c2_%d" % i, f[i], f[i - 1], dict(eps=1e-3, corr=1.0, inner=20), None) for i in (1, 2, 3)]
# C3-shaped: scan vs a denser cloud (three scans merged), corr 0.2, tf_eps 1e-5, 50 inner, with a guess
big = np.concatenate([f[0], f[1], f[2]]).astype(np.float32)
guess = np.eye(4, dtype=np.float32); guess[:3, 3] = (0.05, -0.03, 0.01)
cases.append(("c3ish", f[3], big, dict(eps=1e-5, corr=0.2, inner=50), guess))
cases.append(("corr5", f[2], f[1], dict(eps=1e-3, corr=5.0, inner=20), None))
out = {}
for mode in (0, 1, 3):
    g = locus_b200.GicpB200()
    g.setExecution(mode)
    for rep in range(3):
        g.resetKernelTimes(1 if rep == 2 else 0)
        for name, src, tgt, p, gs in cases:
            g.setTransformationEpsilon(p["eps"]); g.setMaxCorrespondenceDistance(p["corr"]); g.setMaximumIterations(50)
            g.setMaximumOptimizerIterations(p["inner"])
            g.setInputSource(src); g.setInputTarget(tgt)
            t0 = time.perf_counter(); r = g.align(gs); dt = (time.perf_counter() - t0) * 1e3
            if rep == 2:
                out["%s_m%d_T" % (name, mode)] = g.getFinalTransformation()
                out["%s_m%d_s" % (name, mode)] = np.array([r.iterations, r.n_correspondences, r.n_objective_evals, g.getFitnessScore()], dtype=np.float64)
                print("mode", mode, name, "host ms %.3f" % dt, "iters", r.iterations, "ncorr", r.n_correspondences, "evals", r.n_objective_evals,
                      "t_iter ms %.3f" % r.t_iterations_ms)
    k = g.kernelTime("align_persistent")
    print("mode", mode, "align kernel avg ms %.3f x%d" % k, "t_corr cycles", g.kernelTime("debug9")[0], "total", g.kernelTime("debug0")[0])
    if mode == 3: print("   loop_nn", g.kernelTime("loop_nn"), "loop_solve", g.kernelTime("loop_solve"))
    if mode == 1: print("   nn_corr", g.kernelTime("nn_corr"))
np.savez(sys.argv[1], **out)
