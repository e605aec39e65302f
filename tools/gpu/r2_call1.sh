# round 2, GPU call 1: parity tests, bench lines for c2 / c3, ncu launch list of the blocking per-scan calls
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
nproc
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test1.log 2>&1; tail -15 gpurun_out/r2_test1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; tail -3 gpurun_out/r2_bench_c2.err
timeout 900 python bench.py --config c3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; tail -3 gpurun_out/r2_bench_c3.err
LEAF=$(python -c "import json;print(json.load(open('gpurun_out/r2_bench_c2.json'))['config']['leaf_m'])")
echo leaf $LEAF
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --profile --leaf $LEAF --stream-scans 8 > gpurun_out/r2_ncu_l1.log 2>&1
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_c2.json'))
    print('C2 value', d['value'], 'e2e', d['e2e']['value'], 'seq', d['sequential']['value'], d['sequential']['ms_per_scan'], 'seq_e2e', d['sequential_e2e']['value'],
          'cpu', d.get('cpu_baseline',{}).get('value'), d.get('pose_delta_vs_cpu'), d['pipeline_equals_sequential'], d['pipeline_scans_compared'])
    print('  kernels', d['per_scan']['kernels_sequential'], 'launches', d['gpu_launches'], d['sequential']['gpu_launches'])
    print('  variants', {k:(v.get('value'), v.get('equals_sequential'), v.get('pose_delta_vs_cpu')) for k,v in d.get('variants',{}).items()})
    print('  speedups', d.get('speedup_vs_cpu'), 'setup', d.get('setup_s'))
except Exception as e: print('C2 parse failed', e)
try:
    d=json.load(open('gpurun_out/r2_bench_c3.json'))
    print('C3 value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('pose_delta_vs_cpu'), d.get('speedup_vs_cpu'))
    print('  kernels', d['per_scan']['kernels'], d['per_scan']['outer_iterations_mean'], d['per_scan']['objective_evals_mean'])
    v=d['variants']['submap_rebuilt_every_scan']; print('  rebuilt', v['value'], v['e2e'], v.get('cpu_baseline'), v.get('speedup_vs_cpu'), v['equals_resident_submap'])
    print('  submap', d['submap'], d['setup_s'])
except Exception as e: print('C3 parse failed', e)
PY
