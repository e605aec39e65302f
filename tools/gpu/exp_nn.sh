mkdir -p gpurun_out
for m in 0 1 2; do echo "== LB_NN_MODE=$m"; LB_NN_MODE=$m timeout 600 python tools/gpu/exp_nn.py gpurun_out/nn_mode$m.npz 2>&1 | tail -22; done
python - <<'PY'
import numpy as np
a=[np.load('gpurun_out/nn_mode%d.npz'%m) for m in (0,1,2)]
bad=0
for k in a[0].files:
    for m in (1,2):
        if not np.array_equal(a[0][k], a[m][k]): bad+=1; print("DIFF", k, "mode", m, a[0][k].ravel()[:6], a[m][k].ravel()[:6])
print("keys", len(a[0].files), "differences", bad)
PY
