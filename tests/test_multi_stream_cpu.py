"""N>1 host logic of bench.py on CPU: world_size-2 gloo.  The path shards only across independent scan
streams (one per GPU, no data-path collective); what is distributed is the start/stop barrier and the
max-over-ranks of the device time."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import conftest

sys.path.insert(0, conftest.ROOT)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    # every rank owns a different stream
    seeds = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(seeds, torch.tensor([bench.stream_seed(rank)], dtype=torch.int64))
    poses, blobs = bench.make_stream(rank, n_scans=2, beams=4, az=64)
    h = int(np.frombuffer(blobs[0].tobytes(), dtype=np.uint8).astype(np.int64).sum())
    hs = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(hs, torch.tensor([h], dtype=torch.int64))
    # rank r takes (r+1)*10 ms per arm: the slowest rank defines the whole-job time
    dist.barrier()
    tmax, vals = bench.aggregate(dist, "cpu", [10.0 * (rank + 1), 20.0 * (rank + 1)], steps=5, world=world)
    if rank == 0:
        out.put({"seeds": [int(s) for s in seeds], "hashes": [int(x) for x in hs], "tmax": tmax, "vals": vals})
    dist.destroy_process_group()


def test_stream_sharding_and_aggregation_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(set(res["seeds"])) == world and len(set(res["hashes"])) == world     # independent streams
    assert res["tmax"] == [20.0, 40.0]                                                # max over ranks
    assert res["vals"][0] == pytest.approx(5 * world / 0.020) and res["vals"][1] == pytest.approx(5 * world / 0.040)


def test_ping_pong_sequence():
    import bench
    n = bench.N_STREAM
    s = [bench.seq(i) for i in range(4 * n)]
    assert all(abs(a - b) == 1 for a, b in zip(s, s[1:]))      # consecutive steps are neighbouring poses
    assert min(s) == 0 and max(s) == n - 1
