mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_$name.json"))
    print("$name value %.0f e2e %.0f seq %.0f same %s align_ms %.3f seq_align_ms %.3f knn_ms %.3f idx %.3f voxel %.3f launches %d" % (d["value"], d["e2e"]["value"], d["sequential"]["value"], d["pipeline_equals_sequential"], d["roofline"]["avg_launch_ms"], d["roofline"]["avg_launch_ms_sequential"], d["per_scan"]["knn_cov_kernel_ms"], d["per_scan"]["index_build_ms"], d["per_scan"]["voxel_last_call_ms"], d["gpu_launches"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/sw_$name.err").read()[-600:])
PY
}
run a LB_X=1
run b LB_X=1
run ppc512 LB_PIPE_PPC=512
run d8 LB_DEPTH=8
