mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/test33.log 2>&1; tail -1 gpurun_out/test33.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_reference.json 2>/dev/null
LEAF=$(python -c "import json;print(json.load(open('gpurun_out/bench_r1.json'))['config']['leaf_m'])")
echo leaf $LEAF
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --profile --leaf $LEAF --steps 2 --warmup 3 > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches_r1_final_hot.csv python bench.py --profile --leaf $LEAF --steps 2 --warmup 3 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'align_persistent|knn_cov_quad|vg_centroid|rs_scatter|vg_gather' --launch-skip 30 -c 16 -o gpurun_out/prof_r1_final -f python bench.py --profile --leaf $LEAF --steps 2 --warmup 3 > gpurun_out/ncu_f.log 2>&1
grep -c align gpurun_out/launches_r1_final.csv gpurun_out/launches_r1_final_hot.csv
python -c "
import json
d=json.load(open('gpurun_out/bench_r1.json'))
print('RES', d['value'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'), d.get('variants'), d.get('pose_delta_vs_cpu'))
"
