mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 150 --warmup 5 --no-cpu-baseline > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_$name.json"))
    print("$name value %.0f e2e %.0f seq %.0f same %s align_ms %.3f inflight %.2f knn_ms %.3f" % (d["value"], d["e2e"]["value"], d["sequential"]["value"], d["pipeline_equals_sequential"], d["roofline"]["avg_launch_ms"], d["roofline"]["aligns_in_flight_mean"], d["per_scan"]["knn_cov_kernel_ms"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/sw_$name.err").read()[-600:])
PY
}
run dyn1 LB_KNN_DYN=1
run dyn0 LB_KNN_DYN=0
run r50 LB_SM_RESERVE=50
run r70 LB_SM_RESERVE=70
run r85 LB_SM_RESERVE=85
run dyn1b LB_KNN_DYN=1
run dyn0b LB_KNN_DYN=0
timeout 300 python tools/bench_c3.py --steps 10 --warmup 2 --cpu-steps 0 > gpurun_out/c3_dyn.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/c3_dyn.json')); print('C3', d['value'], d['value_submap_index_reused'], d['per_scan'])"
