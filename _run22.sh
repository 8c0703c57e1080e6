python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench16_8gpu.json 2> gpurun_out/bench16_8gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench16_8gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], "pts/s %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "gather us", d.get("residual_gather_us"), "trials", d["trials"]["value"], d["population"]["evals_per_s"], d["trials"]["population"]["value"], d["clocks"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench16_8gpu_ref.json 2>/dev/null; tail -c 300 gpurun_out/bench16_8gpu_ref.json
