# Last check of the tree: every GPU test and smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2zz_test.log 2>&1; tail -4 gpurun_out/r2zz_test.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
