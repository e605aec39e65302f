for v in 0 1 0 1; do
  if [ $v = 1 ]; then export LB_NO_SAMPLER=1; else unset LB_NO_SAMPLER; fi
  timeout 600 python bench.py --no-cpu-baseline --stream-scans 40 > gpurun_out/s_$v.json 2>gpurun_out/s_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/s_$v.json'))
print('no_sampler=$v', 'value %.0f e2e %.0f seq %.0f seq_e2e %.0f' % (d['value'], d['e2e']['value'], d['sequential']['value'], d['sequential_e2e']['value']), d['per_scan']['kernels_sequential']['index_build'], d['clocks'].get('samples'))"
done
