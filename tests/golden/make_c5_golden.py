#!/usr/bin/env python
"""Golden pose of the CPU oracle for the C5-shaped parity case of tests/test_gicp_gpu.py
(test_full_size_c5_dense_properties): 200 k-point scan vs 10 M-point map.  The oracle needs minutes for this
(10 M k-NN(20) covariances + a 10 M-point kd-tree), too long for the -m gpu run, so its result is frozen here.

    python tests/golden/make_c5_golden.py        ->  tests/golden/c5_oracle_pose.npz

The inputs are regenerated from seeds inside the test (fixtures.c5_case); the .npz stores the oracle's pose,
iteration / correspondence counts and a checksum of the inputs, so a drifting generator is caught, not trusted.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fixtures as F  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    src, tgt, Tg, cfg = F.c5_case()
    prm = O.default_params(transformation_epsilon=cfg["tf_eps"], corr_dist_threshold=cfg["corr_dist"],
                           max_iterations=cfg["max_iterations"], max_inner_iterations=cfg["max_inner"],
                           num_threads=os.cpu_count() or 1)
    t0 = time.time()
    r = O.gicp_align(src, tgt, prm)
    print("oracle: %.1f s, iterations %d, n_corr %d, converged %s" % (time.time() - t0, r["iterations"], r["n_corr"], r["converged"]))
    print("pose error vs the known offset:", F.pose_delta(Tg, r["T"]))
    np.savez(os.path.join(HERE, "c5_oracle_pose.npz"), T=r["T"], iterations=r["iterations"], n_corr=r["n_corr"],
             converged=int(r["converged"]), n_evals=r["n_evals"], checksum=F.cloud_checksum(src, tgt))


if __name__ == "__main__":
    main()
