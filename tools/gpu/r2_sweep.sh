# pipeline depth x points-per-align-CTA sweep (sharing on = default); prints value / e2e
for cfg in "6 1024" "8 1024" "10 1024" "6 512" "8 512" "6 2048" "8 2048" "12 2048"; do
  set -- $cfg
  LB_SHARE_VARIANT=0 timeout 300 python bench.py --no-cpu-baseline --stream-scans 40 --depth $1 --pipeline-ppc $2 > gpurun_out/sw_$1_$2.json 2> gpurun_out/sw_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/sw_$1_$2.json'))
print('depth $1 ppc $2', 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['pipeline_stages']['voxel_stage_busy_ms_per_scan'], d['pipeline_stages']['worker_busy_ms_per_scan'], d['roofline']['avg_launch_ms'])"
done
