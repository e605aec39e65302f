mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 150 --warmup 5 --no-cpu-baseline > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_$name.json"))
    st=d["pipeline_stages"]
    print("$name value %.0f e2e %.0f align_ms %.3f inflight %.2f | voxel busy %.3f wait %.3f | worker busy %.3f wait %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["aligns_in_flight_mean"], st["voxel_stage_busy_ms_per_scan"], st["voxel_stage_wait_ms_per_scan"], st["worker_busy_ms_per_scan"], st["worker_wait_ms_per_scan"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/sw_$name.err").read()[-600:])
PY
}
run prio1 LB_VOXEL_PRIO=1
run prio0 LB_VOXEL_PRIO=0
run prio1_d8 LB_VOXEL_PRIO=1 LB_DEPTH=8
run prio1_p2048_d8 LB_VOXEL_PRIO=1 LB_PIPE_PPC=2048 LB_DEPTH=8
run prio1b LB_VOXEL_PRIO=1
run prio0b LB_VOXEL_PRIO=0
