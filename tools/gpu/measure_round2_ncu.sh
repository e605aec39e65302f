# Round-2 ncu set (run under gpurun after measure_round2.sh; needs gpurun_out/r2f_bench_c2.json's leaf or the default)
mkdir -p gpurun_out
LEAF=${LEAF:-0.108088}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches_c2.csv python bench.py --profile --leaf $LEAF --stream-scans 8 > gpurun_out/r2f_ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/r2f_launches_c2_hot.csv python bench.py --profile --leaf $LEAF --stream-scans 8 > gpurun_out/r2f_ncu_l2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'loop_solve|loop_nn|loop_far|knn_cov_quadreg|vg_centroid4|rs_scatter|vg_keys' --launch-skip 60 -c 7 -o gpurun_out/r2f_prof -f python bench.py --profile --leaf $LEAF --stream-scans 8 > gpurun_out/r2f_ncu_f.log 2>&1
LB_NN=staged_tma timeout 600 ncu --set full --clock-control none -k regex:'nn_query_staged|nn_query_far' --launch-skip 2 -c 2 -o gpurun_out/r2f_prof_nn -f python tools/nn_roofline.py --cells 0 --reps 3 > gpurun_out/r2f_ncu_nn.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out | tail -1
