mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_$name.json"))
    print("$name value %.0f e2e %.0f seq %.0f same %s align_ms %.3f seq_align_ms %.3f probes %s knn_ms %.3f idx_ms %.3f" % (d["value"], d["e2e"]["value"], d["sequential"]["value"], d["pipeline_equals_sequential"], d["roofline"]["avg_launch_ms"], d["roofline"]["avg_launch_ms_sequential"], d["per_scan"]["cell_probe_rounds_total"], d["per_scan"]["knn_cov_kernel_ms"], d["per_scan"]["index_build_ms"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/sw_$name.err").read()[-600:])
PY
}
run kq4 LB_X=1
run kq5 LOCUS_B200_LIB=/root/repo/locus_b200/liblocus_b200_kq5.so
run kq4_d6 LB_DEPTH=6
run kq5_d6 LOCUS_B200_LIB=/root/repo/locus_b200/liblocus_b200_kq5.so LB_DEPTH=6
ncu --set full --clock-control none --import-source on -k regex:knn_cov_quadreg --launch-skip 6 -c 2 -o gpurun_out/prof_knnreg2 -f python bench.py --profile --leaf 0.10808803886175156 --steps 2 --warmup 3 > gpurun_out/ncu_knnreg2.log 2>&1
