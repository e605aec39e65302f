# first-look radius (cells) and points per solve CTA: t_iter per case (stream-ordered default)
for r in 0.35 0.5 0.75 1.0; do echo "== LB_NN_R0=$r"; LB_NN_R0=$r python tools/gpu/exp_nn.py gpurun_out/tmp.npz 2>&1 | grep "^mode 3 c\|loop_nn"; done
for p in 512 1024; do echo "== LB_PPC=$p"; LB_PPC=$p python tools/gpu/exp_nn.py gpurun_out/tmp.npz 2>&1 | grep "^mode 3 c"; done
