# default execution (persistent, serial search) vs stream-ordered + staged TMA search, sequential seam and pipeline
mkdir -p gpurun_out
for e in -1 3; do
  LB_EXEC_OVERRIDE=$e LB_SHARE_VARIANT=0 timeout 600 python bench.py --no-cpu-baseline --stream-scans 60 > gpurun_out/ex_$e.json 2> gpurun_out/ex_$e.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ex_$e.json'))
print("exec override $e: value %.0f e2e %.0f seq %.0f (%.3f ms/scan) seq_e2e %.0f equal %s" % (d['value'], d['e2e']['value'], d['sequential']['value'], d['sequential']['ms_per_scan'], d['sequential_e2e']['value'], d['pipeline_equals_sequential']))
PY
done
