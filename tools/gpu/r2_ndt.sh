# NDT (row f4) on the GPU: parity tests, smoke, bench line, launch list, one full ncu capture of the evaluation kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ndt_gpu.py -x -q 2>&1 | tail -25
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tools/bench_ndt.py > gpurun_out/r2_bench_ndt.json 2> gpurun_out/r2_bench_ndt.err; tail -c 1500 gpurun_out/r2_bench_ndt.json; tail -3 gpurun_out/r2_bench_ndt.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ndt.csv python tools/bench_ndt.py --profile --steps 1 --warmup 1 > gpurun_out/r2_ncu_ndt_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ndt_eval|ndt_ctl|ndt_gaussians' --launch-skip 6 -c 5 -o gpurun_out/r2_prof_ndt -f python tools/bench_ndt.py --profile --steps 1 --warmup 1 > gpurun_out/r2_ncu_ndt_f.log 2>&1
ls -la gpurun_out/r2_prof_ndt.ncu-rep gpurun_out/r2_launches_ndt.csv
