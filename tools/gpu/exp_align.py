"""align() on one C2 scan pair in the execution modes: time per align, iteration / evaluation counts, the kernels'
cycle counters.  usage: python tools/gpu/exp_align.py [leaf]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures as F
import locus_b200
from tools import gen_lidar as G
leaf = float(sys.argv[1]) if len(sys.argv) > 1 else 0.108088
G.WORKERS = 4
scene, poses, blobs = G.stream(2, 4)
vg = locus_b200.VoxelGridB200(); vg.setLeafSize(leaf); vg.setFilterFieldName("z"); vg.setFilterLimits(-100, 100)
f = [np.ascontiguousarray(vg.filter(b, 32, locus_b200.xyzi_fields())).view(np.float32).reshape(-1, 8)[:, :3].copy() for b in blobs]
print("points", [len(x) for x in f])
Tref = {}
for mode in (0, 3, 3):
    g = locus_b200.GicpB200()
    g.setTransformationEpsilon(1e-3); g.setMaxCorrespondenceDistance(1.0); g.setMaximumIterations(50); g.setExecution(mode)
    for rep in range(3):
        g.resetKernelTimes(1 if rep == 2 else 0)
        ts = []
        for i in (1, 2, 3):
            g.setInputSource(f[i]); g.setInputTarget(f[i - 1])
            t0 = time.perf_counter(); r = g.align(); ts.append((time.perf_counter() - t0) * 1e3)
            T = g.getFinalTransformation()
            if (mode, i) not in Tref: Tref[(mode, i)] = T
        if rep == 2:
            k = g.kernelTime("align_persistent")
            dbg = [g.kernelTime("debug%d" % j)[0] for j in range(10)]
            if mode == 3: print("   loop_nn", g.kernelTime("loop_nn"), "loop_solve", g.kernelTime("loop_solve"))
            cd = [g.kernelTime("dbg%d" % j)[0] for j in range(10, 16)]
            print("   corr: own loop %.0f  nn %.0f  finish %.0f  points %.0f  hits+cache %.0f  allreduce %.0f" % tuple(cd))
            print("mode", mode, "host ms/align", ["%.3f" % t for t in ts], "kernel ms %.3f x%d" % k, "iters", r.iterations, "evals", r.n_objective_evals,
                  "dbg total %.0f acc %.0f sync %.0f ncoll %.0f scalar %.0f d7 %.0f d8 %.0f d9 %.0f" % (dbg[0], dbg[1], dbg[2], dbg[3], dbg[6], dbg[7], dbg[8], dbg[9]))
for i in (1, 2, 3):
    if (2, i) in Tref: print("pair", i, "mode2 vs mode0", F.pose_delta(Tref[(0, i)], Tref[(2, i)]))
    print("pair", i, "mode3 == mode0", np.array_equal(Tref[(0, i)], Tref[(3, i)]))
