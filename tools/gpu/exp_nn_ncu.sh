mkdir -p gpurun_out
for m in 0 1; do
LB_NN_MODE=$m timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:loop_nn_kernel -c 60 --csv --log-file gpurun_out/nn_list_m$m.csv python tools/gpu/exp_nn.py gpurun_out/tmp.npz > gpurun_out/nn_ncu_m$m.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/nn_list_m$m.csv') if not l.startswith('=='))]
print("mode $m loop_nn durations (us):", [round(float(r['Metric Value'].replace(',',''))/1000,1) for r in rows][:40])
PY
done
LB_NN_MODE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:loop_nn_kernel --launch-skip 0 -c 3 -o gpurun_out/prof_nn_staged -f python tools/gpu/exp_nn.py gpurun_out/tmp.npz > gpurun_out/nn_ncu_full.log 2>&1
ls -la gpurun_out/prof_nn_staged.ncu-rep
