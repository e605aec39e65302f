"""micro-benchmark: host wall time of lb_gicp_set_source / set_target in steady state, and device allocations"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures as F
import locus_b200
g = locus_b200.GicpB200()
clouds = [F.random_scene(30000 + 700 * i, i) for i in range(8)]
for rep in range(3):
    a0 = g.kernelTime("dbuf_allocs")[0]
    t0 = time.perf_counter()
    for i in range(40):
        g.setInputSource(clouds[i % 8])
    t1 = time.perf_counter()
    for i in range(40):
        g.setInputTarget(clouds[(i + 3) % 8])
    t2 = time.perf_counter()
    for i in range(20):
        g.setInputSource(clouds[i % 8]); g.setInputTarget(clouds[(i + 1) % 8]); g.align()
    t3 = time.perf_counter()
    print("rep", rep, "set_source ms %.3f" % ((t1 - t0) / 40 * 1e3), "set_target ms %.3f" % ((t2 - t1) / 40 * 1e3),
          "src+tgt+align ms %.3f" % ((t3 - t2) / 20 * 1e3), "allocs", g.kernelTime("dbuf_allocs")[0] - a0, "pool", g.kernelTime("pool_clouds")[0])
