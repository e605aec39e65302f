mkdir -p gpurun_out
for c in 0.37 0.5 0.65 0.8 1.0; do
LB_CELL=$c timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/c2_$c.json 2> gpurun_out/c2_$c.err
python -c "
import json
d=json.load(open('gpurun_out/c2_$c.json'))
print('cell $c value %.0f seq %.0f knn_ms %.3f align_ms_seq %.3f idx %.3f' % (d['value'], d['sequential']['value'], d['per_scan']['knn_cov_kernel_ms'], d['roofline']['avg_launch_ms_sequential'], d['per_scan']['index_build_ms']))
"
done
