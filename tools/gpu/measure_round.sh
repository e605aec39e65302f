# Round-end measurement set (run under gpurun on one B200): tests, bench lines, ncu launch lists and full captures.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/test_final.log 2>&1; tail -1 gpurun_out/test_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_reference.json 2>/dev/null
python tools/bench_c3.py --steps 20 --warmup 3 --cpu-steps 2 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python tools/nn_roofline.py --cells 0,0.225 > gpurun_out/nn_roofline_final.json 2> gpurun_out/nn_roofline.err
LEAF=$(python -c "import json;print(json.load(open('gpurun_out/bench_r1.json'))['config']['leaf_m'])")
echo leaf $LEAF
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --profile --leaf $LEAF --steps 2 --warmup 3 --scans-per-step 1 > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches_r1_final_hot.csv python bench.py --profile --leaf $LEAF --steps 2 --warmup 3 --scans-per-step 1 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'align_persistent|knn_cov_quadreg|knn_cov_tail|vg_centroid|rs_scatter|vg_gather' --launch-skip 30 -c 18 -o gpurun_out/prof_r1_final -f python bench.py --profile --leaf $LEAF --steps 2 --warmup 3 --scans-per-step 1 > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none -k regex:nn_query_warp --launch-skip 2 -c 2 -o gpurun_out/prof_r1_nn_warp -f python tools/nn_roofline.py --cells 0.225 --reps 3 > gpurun_out/ncu_nn.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/bench_r1.json'))
print('RES', d['value'], d['e2e']['value'], d['sequential']['value'], d.get('cpu_baseline',{}).get('value'), d.get('pose_delta_vs_cpu'), d['pipeline_equals_sequential'])
d=json.load(open('gpurun_out/bench_c3.json')); print('C3', d['value'], d['value_submap_index_reused'], d['cpu_baseline']['value'], d['pose_delta_vs_cpu'])
d=json.load(open('gpurun_out/nn_roofline_final.json')); print('NN', [(r['cell_m'], r['kernel_ms'], r['frac']) for r in d['runs']])
"
