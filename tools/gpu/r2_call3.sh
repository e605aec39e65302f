mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_test3.log 2>&1; tail -25 gpurun_out/r2_test3.log
