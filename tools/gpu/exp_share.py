"""long-run behaviour of the odometry pipeline with and without cloud sharing: scans/s per batch of 600 scans, device
allocations per batch (must be 0 in steady state), free memory."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import locus_b200
from locus_b200 import api
from tools import gen_lidar as G
G.WORKERS = 16
scene, poses, blobs = G.stream(2, 40)
G.WORKERS = 1
n = blobs[0].size // 32
fields = locus_b200.xyzi_fields(); fa = api.VoxelGridB200._fields(fields)
d = [torch.from_numpy(b).cuda() for b in blobs]
def seq(i, m=len(blobs)):
    p = 2 * (m - 1); r = i % p
    return r if r < m else p - r
for share in (0, 1, 0, 1):
    odo = locus_b200.OdometryB200(0, depth=6, max_points=n, max_point_step=32)
    odo.voxel.setFilterFieldName("z"); odo.voxel.setFilterLimits(-100, 100); odo.voxel.setLeafSize(0.108088)
    odo.setGicpParams(transformation_epsilon=1e-3, max_correspondence_distance=1.0, max_iterations=50, align_points_per_cta=1024)
    odo.setCloudSharing(bool(share))
    g0 = odo.gicp(0)
    i = 0
    rates = []
    for batch in range(6):
        a0 = g0.kernelTime("dbuf_allocs")[0]; free0 = torch.cuda.mem_get_info()[0]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(600):
            odo.submit(d[seq(i)].data_ptr(), n, 32, fa, mem=locus_b200.LB_MEM_DEVICE); i += 1
            r = odo.next(block=False)
            while r is not None:
                r = odo.next(block=False) if odo.pending() else None
        while odo.pending():
            odo.next()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        rates.append(600 / dt)
        print("share", share, "batch", batch, "scans/s %.0f" % (600 / dt), "allocs", g0.kernelTime("dbuf_allocs")[0] - a0,
              "mem taken MB %.1f" % ((free0 - torch.cuda.mem_get_info()[0]) / 1e6), flush=True)
    odo.close()
