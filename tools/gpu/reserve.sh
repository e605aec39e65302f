mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 150 --warmup 5 --no-cpu-baseline > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_$name.json"))
    print("$name value %.0f e2e %.0f seq %.0f align_ms %.3f inflight %.2f" % (d["value"], d["e2e"]["value"], d["sequential"]["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["aligns_in_flight_mean"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/sw_$name.err").read()[-600:])
PY
}
run r16 LB_SM_RESERVE=16
run r50 LB_SM_RESERVE=50
run r70 LB_SM_RESERVE=70
run r85 LB_SM_RESERVE=85
run r50_ppc512 LB_SM_RESERVE=28 LB_PIPE_PPC=512
run r16b LB_SM_RESERVE=16
run r50b LB_SM_RESERVE=50
