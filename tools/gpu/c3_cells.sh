mkdir -p gpurun_out
for c in 0 0.13 0.16 0.2 0.26; do
timeout 300 python tools/bench_c3.py --steps 6 --warmup 2 --cpu-steps 0 --cell $c > gpurun_out/c3_$c.json 2> gpurun_out/c3_$c.err
python -c "
import json
d=json.load(open('gpurun_out/c3_$c.json'))
print('cell $c value %.1f reuse %.1f knn_ms %.2f align_ms %.2f' % (d['value'], d['value_submap_index_reused'], d['per_scan']['knn_cov_ms_both_clouds'], d['per_scan']['align_kernel_ms']))
"
done
for q in 2 3; do
LB_QSPLIT=$q timeout 300 python tools/bench_c3.py --steps 6 --warmup 2 --cpu-steps 0 > gpurun_out/c3_q$q.json 2> gpurun_out/c3_q$q.err
python -c "
import json
d=json.load(open('gpurun_out/c3_q$q.json'))
print('qsplit $q value %.1f reuse %.1f knn_ms %.2f align_ms %.2f' % (d['value'], d['value_submap_index_reused'], d['per_scan']['knn_cov_ms_both_clouds'], d['per_scan']['align_kernel_ms']))
"
done
