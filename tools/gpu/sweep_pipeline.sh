mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_$name.json"))
    print("$name value %.0f e2e %.0f seq %.0f same %s align_ms %.3f seq_align_ms %.3f probes %s" % (d["value"], d["e2e"]["value"], d["sequential"]["value"], d["pipeline_equals_sequential"], d["roofline"]["avg_launch_ms"], d["roofline"]["avg_launch_ms_sequential"], d["per_scan"]["cell_probe_rounds_total"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/sw_$name.err").read()[-600:])
PY
}
for ppc in 512 1024; do for d in 3 4 6; do
run a_${ppc}_$d LB_PIPE_PPC=$ppc LB_DEPTH=$d
done; done
for ppc in 512 1024; do for d in 4 6 8; do
run mb2_${ppc}_$d LOCUS_B200_LIB=/root/repo/locus_b200/liblocus_b200_mb2.so LB_SM_RESERVE=-120 LB_PIPE_PPC=$ppc LB_DEPTH=$d
done; done
run a_1024_4_again LB_PIPE_PPC=1024 LB_DEPTH=4
run a_1024_3_again LB_PIPE_PPC=1024 LB_DEPTH=3
