"""Child process of test_gicp_gpu.py::test_correspondence_search_modes_agree: align() on a fixed set of cases with the
correspondence search forced to LB_NN_MODE (read once per process by the library); results -> argv[1] (.npz)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import fixtures as F   # noqa: E402
import locus_b200      # noqa: E402


def cases():
    src, tgt = F.garage()                                   # 2277 duplicated points: ties between equidistant candidates
    yield "garage", src, tgt, dict(corr=5.0, eps=5e-4, inner=20), None
    a = F.random_scene(20000, 3)
    T = F.se3([0.08, -0.05, 0.02], [0.002, -0.003, 0.01])
    b = (a @ T[:3, :3].T + T[:3, 3]).astype(np.float32)[::-1].copy()
    yield "scene_corr1", b, a, dict(corr=1.0, eps=1e-3, inner=20), None
    yield "scene_corr02", b, a, dict(corr=0.2, eps=1e-5, inner=50), None          # gate about one cell: many unmatched points
    yield "scene_corr5", b, a, dict(corr=5.0, eps=1e-3, inner=20), None
    far = F.se3([0.9, 0.4, 0.1], [0, 0, 0.05]).astype(np.float32)                # a poor guess: most first-iteration queries undecided
    yield "scene_badguess", b, a, dict(corr=1.0, eps=1e-3, inner=20), far
    part = a[a[:, 0] < np.median(a[:, 0])]                                        # half of the source has no counterpart in the target
    yield "scene_half_target", b, part, dict(corr=0.5, eps=1e-4, inner=20), None
    yield "tiny", a[:37], a[:50], dict(corr=1.0, eps=1e-3, inner=20), None


out = {}
for execution in (0, 3, 1):
    g = locus_b200.GicpB200()
    g.setExecution(execution)
    for name, s, t, p, guess in cases():
        if execution == 1 and name not in ("garage", "scene_badguess"): continue    # host-driven: two cases are enough
        g.setMaxCorrespondenceDistance(p["corr"]); g.setTransformationEpsilon(p["eps"]); g.setMaximumIterations(30)
        g.setMaximumOptimizerIterations(p["inner"])
        g.setInputSource(s); g.setInputTarget(t)
        r = g.align(guess)
        out["%s_e%d_T" % (name, execution)] = g.getFinalTransformation()
        out["%s_e%d_s" % (name, execution)] = np.array([r.iterations, r.n_correspondences, r.n_objective_evals, r.converged, g.getFitnessScore()], dtype=np.float64)
# the same searches behind lb_gicp_nn_target (LB_NN picks the kernel): indices and squared distances of arbitrary queries
a = F.random_scene(20000, 3)
rng = np.random.default_rng(11)
q = np.concatenate([a[rng.integers(0, len(a), 3000)] + rng.normal(0, 0.05, (3000, 3)),       # near the surfaces
                    rng.uniform(-5, 25, (500, 3)),                                           # anywhere, also outside the grid
                    a[:200]]).astype(np.float32)                                             # exact hits
g = locus_b200.GicpB200()
g.setInputTarget(a)
idx, d2 = g.nearestTarget(q)
out["nn_idx"] = idx; out["nn_d2"] = d2
src, tgt = F.garage()
g.setInputTarget(tgt)
idx, d2 = g.nearestTarget(src)
out["nn_garage_idx"] = idx; out["nn_garage_d2"] = d2
np.savez(sys.argv[1], **out)
