# round 2, GPU call 2: tests after the parity-envelope rework + submap, bench c2/c3, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_test2.log 2>&1; tail -25 gpurun_out/r2_test2.log
timeout 900 python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; tail -3 gpurun_out/r2_bench_c2.err
timeout 900 python bench.py --config c3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; tail -3 gpurun_out/r2_bench_c3.err
LEAF=$(python -c "import json;print(json.load(open('gpurun_out/r2_bench_c2.json'))['config']['leaf_m'])")
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --profile --leaf $LEAF --stream-scans 8 > gpurun_out/r2_ncu_l1.log 2>&1
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_c2.json'))
    print('C2 value', d['value'], 'e2e', d['e2e']['value'], 'seq', d['sequential']['value'], d['sequential']['ms_per_scan'], 'seq_e2e', d['sequential_e2e']['value'],
          'cpu', d.get('cpu_baseline',{}).get('value'), d['pipeline_equals_sequential'], d['pipeline_scans_compared'])
    print('  pose', json.dumps(d.get('pose_delta_vs_cpu')))
    print('  kernels', d['per_scan']['kernels_sequential'], 'launches', d['gpu_launches'], d['sequential']['gpu_launches'])
    print('  variants', {k:(v.get('value'), v.get('equals_sequential'), v.get('error')) for k,v in d.get('variants',{}).items()})
    print('  speedups', d.get('speedup_vs_cpu'))
except Exception as e: print('C2 parse failed', e)
try:
    d=json.load(open('gpurun_out/r2_bench_c3.json'))
    print('C3 value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('speedup_vs_cpu'))
    print('  pose', json.dumps(d.get('pose_delta_vs_cpu')))
    print('  kernels', d['per_scan']['kernels'], d['per_scan']['outer_iterations_mean'], d['per_scan']['objective_evals_mean'])
    v=d['variants']['submap_rebuilt_every_scan']; print('  rebuilt', v['value'], v['e2e'], v.get('cpu_baseline'), v.get('speedup_vs_cpu'), v['equals_resident_submap'])
    print('  rolling', json.dumps(d['variants'].get('rolling_submap')))
except Exception as e: print('C3 parse failed', e)
PY
