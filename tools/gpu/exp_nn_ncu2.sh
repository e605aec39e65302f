mkdir -p gpurun_out
for m in 0 2; do
LB_NN_MODE=$m timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none --cache-control none -k regex:'loop_nn_kernel|loop_far_kernel' -c 80 --csv --log-file gpurun_out/nn_list_m$m.csv python tools/gpu/exp_nn.py gpurun_out/tmp.npz > gpurun_out/nn_ncu_m$m.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/nn_list_m$m.csv') if not l.startswith('=='))]
by={}
for r in rows:
    by.setdefault(r['ID'],{})[r['Metric Name']]=float(r['Metric Value'].replace(',','')); by[r['ID']]['k']=r['Kernel Name'][:12]
out=[]
for i in sorted(by,key=int):
    d=by[i]; out.append("%s %.1fus %.0fkcyc %.0fkinst"%(d['k'][5:8], d['gpu__time_duration.sum']/1000, d['sm__cycles_elapsed.max']/1000, d['smsp__inst_executed.sum']/1000))
print("mode $m:", " | ".join(out[:48]))
PY
done
