# Round-2 measurement set (run under gpurun on one B200): tests, smoke, bench lines (c2 / reference arm / c3 / c5), NN roofline A/B,
# ncu launch lists and full captures.  Outputs under gpurun_out/r2f_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2f_test.log 2>&1; tail -2 gpurun_out/r2f_test.log
for m in 0 1 2; do LB_NN_MODE=$m timeout 300 python tools/gpu/exp_nn.py gpurun_out/r2f_nnmode$m.npz > gpurun_out/r2f_nnmode$m.log 2>&1; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_c2_reference.json 2>/dev/null
timeout 900 python bench.py --config c3 > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err
timeout 900 python bench.py --config c5 --no-cpu-baseline > gpurun_out/r2f_bench_c5.json 2> gpurun_out/r2f_bench_c5.err
for v in warp staged staged_tma default; do if [ $v = default ]; then unset LB_NN; else export LB_NN=$v; fi; timeout 600 python tools/nn_roofline.py --cells 0 --reps 5 > gpurun_out/r2f_nn_$v.json 2>/dev/null; done; unset LB_NN
python - <<'PY'
import json
def L(f):
    try: return json.load(open(f))
    except Exception as e: print("parse failed", f, e); return None
d=L('gpurun_out/r2f_bench_c2.json')
if d:
    print('C2 value', d['value'], 'e2e', d['e2e']['value'], 'seq', d['sequential']['value'], d['sequential']['ms_per_scan'], 'seq_e2e', d['sequential_e2e']['value'],
          'cpu', d.get('cpu_baseline',{}).get('value'), d['pipeline_equals_sequential'])
    print('  pose', json.dumps(d.get('pose_delta_vs_cpu'))[:600]); print('  speedups', d.get('speedup_vs_cpu')); print('  roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
    print('  variants', {k:(v.get('value'), v.get('equals_sequential'), v.get('error')) for k,v in d.get('variants',{}).items()})
d=L('gpurun_out/r2f_bench_c3.json')
if d:
    print('C3 value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('speedup_vs_cpu'))
    print('  pose', json.dumps(d.get('pose_delta_vs_cpu'))[:600])
    v=d['variants']['submap_rebuilt_every_scan']; print('  rebuilt', v['value'], v['e2e'], v.get('cpu_baseline'), v.get('speedup_vs_cpu'))
    print('  rolling', json.dumps(d['variants'].get('rolling_submap'))[:300])
d=L('gpurun_out/r2f_bench_c5.json')
if d: print('C5 value', d['value'], 'e2e', d['e2e']['value'], 'roofline', {k:d['roofline'].get(k) for k in ('frac','achieved','avg_launch_ms','queries_per_s','candidates_per_query')})
for v in ('warp','staged','staged_tma','default'):
    d=L('gpurun_out/r2f_nn_%s.json'%v)
    if d: r=d['runs'][0]; print('NN', v, r['kernel_ms'], r['queries_per_s'], r['candidates_per_query'], r['frac'])
PY
du -sh gpurun_out | tail -1
