timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
PYXRD_B200_OVERLAP=$v python bench.py --steps 10 --warmup 3 --no-cpu --no-trials --workload c4 > gpurun_out/b17.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/b17.json").read().strip().splitlines()[-1])
print("overlap $v c4", "pts/s %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d["stage_ms"].items()})
PY
done
python bench.py --no-cpu > gpurun_out/b17.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/b17.json").read().strip().splitlines()[-1])
print("c2", "pts/s %.3e"%d["value"], d["stage_ms"]); t=d["trials"]; print("trials %.0f"%t["value"], t["stage_ms"], t["population"]["value"])
PY
