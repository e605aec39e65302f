mkdir -p gpurun_out
LOCUS_B200_LIB=/root/repo/locus_b200/liblocus_b200_w7.so timeout 600 python -m pytest tests/test_gicp_gpu.py tests/test_odometry_gpu.py -x -q 2>&1 | tail -2
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_$name.json"))
    print("$name value %.0f e2e %.0f seq %.0f same %s align_ms %.3f seq_align_ms %.3f knn_ms %.3f" % (d["value"], d["e2e"]["value"], d["sequential"]["value"], d["pipeline_equals_sequential"], d["roofline"]["avg_launch_ms"], d["roofline"]["avg_launch_ms_sequential"], d["per_scan"]["knn_cov_kernel_ms"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/sw_$name.err").read()[-600:])
PY
}
W7=/root/repo/locus_b200/liblocus_b200_w7.so
run base LB_X=1
run w7_ppc0 LOCUS_B200_LIB=$W7 LB_PIPE_PPC=0
run w7_ppc0_d8 LOCUS_B200_LIB=$W7 LB_PIPE_PPC=0 LB_DEPTH=8
run w7_ppc1792 LOCUS_B200_LIB=$W7 LB_PIPE_PPC=1792
run base2 LB_X=1
run w7_ppc0b LOCUS_B200_LIB=$W7 LB_PIPE_PPC=0
