# NDT (row f4) on the GPU: parity tests (both evaluation kernels), smoke, bench lines, launch list, ncu captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ndt_gpu.py -x -q 2>&1 | tail -25
LB_NDT_EVAL=thread timeout 600 python -m pytest tests/test_ndt_gpu.py -x -q -k "align or derivatives" 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tools/bench_ndt.py > gpurun_out/r2_bench_ndt.json 2> gpurun_out/r2_bench_ndt.err; tail -c 1500 gpurun_out/r2_bench_ndt.json; tail -3 gpurun_out/r2_bench_ndt.err
LB_NDT_EVAL=thread timeout 300 python tools/bench_ndt.py --profile > gpurun_out/r2_bench_ndt_thread.json 2> gpurun_out/r2_bench_ndt_thread.err; tail -c 700 gpurun_out/r2_bench_ndt_thread.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_ndt.csv python tools/bench_ndt.py --profile --steps 1 --warmup 1 > gpurun_out/r2_ncu_ndt_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ndt_eval_group|ndt_ctl|ndt_gaussians' --launch-skip 2 -c 6 -o gpurun_out/r2_prof_ndt -f python tools/bench_ndt.py --profile --steps 1 --warmup 0 > gpurun_out/r2_ncu_ndt_f.log 2>&1
ls -la gpurun_out/r2_prof_ndt.ncu-rep gpurun_out/r2_launches_ndt.csv
