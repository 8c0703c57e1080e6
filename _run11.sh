for n in 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench11_${n}gpu.json 2> gpurun_out/bench11_${n}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench11_${n}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], "pts/s %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "gather us", d.get("residual_gather_us"), "trials", d["trials"]["value"], d.get("population",{}).get("evals_per_s"), d["cpu_affinity"], d["trials"]["population"]["value"])
PY
done
nvidia-smi topo -m | head -14
