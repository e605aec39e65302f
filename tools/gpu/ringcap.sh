mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gicp_gpu.py -x -q 2>&1 | tail -2
for cap in 2 3 4 6 99; do
LB_RING_CAP=$cap timeout 300 python tools/bench_c3.py --steps 6 --warmup 2 --cpu-steps 0 > gpurun_out/c3_cap$cap.json 2> gpurun_out/c3_cap$cap.err
LB_RING_CAP=$cap timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/c2_cap$cap.json 2> gpurun_out/c2_cap$cap.err
python -c "
import json
d=json.load(open('gpurun_out/c3_cap$cap.json'))
print('cap $cap C3 value %.1f reuse %.1f knn_ms %.2f align_ms %.2f' % (d['value'], d['value_submap_index_reused'], d['per_scan']['knn_cov_ms_both_clouds'], d['per_scan']['align_kernel_ms']))
d=json.load(open('gpurun_out/c2_cap$cap.json'))
print('cap $cap C2 value %.0f e2e %.0f seq %.0f knn_ms %.3f' % (d['value'], d['e2e']['value'], d['sequential']['value'], d['per_scan']['knn_cov_kernel_ms']))
"
done
