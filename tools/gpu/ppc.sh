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
run p1024_d6 LB_PIPE_PPC=1024 LB_DEPTH=6
run p2048_d8 LB_PIPE_PPC=2048 LB_DEPTH=8
run p2048_d12 LB_PIPE_PPC=2048 LB_DEPTH=12
run p1536_d8 LB_PIPE_PPC=1536 LB_DEPTH=8
run p1024_d10 LB_PIPE_PPC=1024 LB_DEPTH=10
run p768_d6 LB_PIPE_PPC=768 LB_DEPTH=6
