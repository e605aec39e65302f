# Final check of the tree (what the driver runs at round end): every GPU test, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2z_test.log 2>&1; tail -4 gpurun_out/r2z_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2z_bench_c2.json 2> gpurun_out/r2z_bench_c2.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2z_bench_c2.json'))
    print('C2 value', d['value'], 'e2e', d['e2e']['value'], 'seq', d['sequential']['value'], 'seq_e2e', d['sequential_e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'pipe==seq', d.get('pipeline_equals_sequential'), 'clocks', d.get('clocks'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2z_bench_c2.err').read()[-1500:])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-400
