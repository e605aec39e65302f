"""the blocking per-scan calls on device-resident buffers, like bench.py's sequential arm: wall time per phase, allocations"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import locus_b200
from locus_b200 import api
from tools import gen_lidar as G
G.WORKERS = 16
scene, poses, blobs = G.stream(2, 30)
G.WORKERS = 1
L = locus_b200.lib()
n = blobs[0].size // 32
stream = torch.cuda.Stream()
fields = locus_b200.xyzi_fields(); fa = api.VoxelGridB200._fields(fields)
vg = locus_b200.VoxelGridB200(0, stream=stream.cuda_stream)
vg.setFilterFieldName("z"); vg.setFilterLimits(-100, 100); vg.setLeafSize(0.108088)
g = locus_b200.GicpB200(0, stream=stream.cuda_stream)
g.setTransformationEpsilon(1e-3); g.setMaxCorrespondenceDistance(1.0); g.setMaximumIterations(50)
with torch.cuda.stream(stream):
    d = [torch.from_numpy(b).cuda() for b in blobs]
    filt = [torch.empty(n * 32, dtype=torch.uint8, device="cuda") for _ in range(2)]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
n_out = C.c_size_t(0); res = api.GicpResult()
for mode in ("timers off", "timers on", "timers off + flush", "timers on + flush"):
    g.resetKernelTimes(1 if "on" in mode else 0)
    tv = ts = tt = ta = 0.0; cnt = 0; n_prev = 0
    a0 = g.kernelTime("dbuf_allocs")[0]
    for rep in range(2):
        for i in range(len(blobs)):
            cur, prv = filt[i & 1], filt[(i + 1) & 1]
            if "flush" in mode:
                with torch.cuda.stream(stream):
                    flush.fill_(i)
            t0 = time.perf_counter()
            L.lb_voxel_filter(vg._h, C.c_void_p(d[i].data_ptr()), n, 32, fa, len(fields), None, 0, C.c_void_p(cur.data_ptr()), n, C.byref(n_out), None, 1, 1)
            t1 = time.perf_counter()
            nc = n_out.value
            if n_prev:
                L.lb_gicp_set_source(g._h, C.c_void_p(cur.data_ptr()), nc, 32, 0, -1, 1)
                t2 = time.perf_counter()
                L.lb_gicp_set_target(g._h, C.c_void_p(prv.data_ptr()), n_prev, 32, 0, -1, 1, None)
                t3 = time.perf_counter()
                L.lb_gicp_align(g._h, None, C.byref(res))
                t4 = time.perf_counter()
                if rep == 1:
                    tv += t1 - t0; ts += t2 - t1; tt += t3 - t2; ta += t4 - t3; cnt += 1
            n_prev = nc
    print(mode, "ms per scan: voxel %.3f set_source %.3f set_target %.3f align %.3f total %.3f" % tuple(1e3 * x / cnt for x in (tv, ts, tt, ta, tv + ts + tt + ta)),
          "allocs", g.kernelTime("dbuf_allocs")[0] - a0, "idx_build", g.kernelTime("index_build"), "knn", g.kernelTime("knn_cov"), flush=True)
