"""LB_NNPROF=1 LB_NN_MODE=1: per-warp cycle breakdown of the staged correspondence search (stderr lines from the library)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import locus_b200
from tools import gen_lidar as G
G.WORKERS = 4
scene, poses, blobs = G.stream(2, 4)
vg = locus_b200.VoxelGridB200(); vg.setLeafSize(0.108088); vg.setFilterFieldName("z"); vg.setFilterLimits(-100, 100)
f = [np.ascontiguousarray(vg.filter(b, 32, locus_b200.xyzi_fields())).view(np.float32).reshape(-1, 8)[:, :3].copy() for b in blobs]
big = np.concatenate([f[0], f[1], f[2]]).astype(np.float32)
for name, src, tgt, eps, corr, inner in (("c2", f[1], f[0], 1e-3, 1.0, 20), ("c3ish", f[3], big, 1e-5, 0.2, 50)):
    for mode in (3, 0):
        for maxit in (1, 50):
            g = locus_b200.GicpB200(); g.setExecution(mode)
            g.setTransformationEpsilon(eps); g.setMaxCorrespondenceDistance(corr); g.setMaximumIterations(maxit); g.setMaximumOptimizerIterations(inner)
            g.setInputSource(src); g.setInputTarget(tgt)
            for rep in range(2):
                sys.stderr.write("== %s exec %d max_iterations %d rep %d\n" % (name, mode, maxit, rep)); sys.stderr.flush()
                g.align()
