mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print("N$N value %.0f e2e %.0f seq %.0f n_gpus %d same %s stages %s" % (d["value"], d["e2e"]["value"], d["sequential"]["value"], d["n_gpus"], d["pipeline_equals_sequential"], d["pipeline_stages"]))
PY
nproc; free -g | head -2
