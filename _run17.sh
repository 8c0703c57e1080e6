for w in c1 c3 c5; do python bench.py --steps 10 --warmup 3 --no-cpu --no-trials --workload $w > gpurun_out/b12_$w.json 2>gpurun_out/b12_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/b12_$w.json").read().strip().splitlines()[-1])
print("$w", "K", d["config"]["candidates_per_gpu"], "pts/s %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "roof %.3f"%d["roofline"]["frac"], {k:round(v,3) for k,v in d["stage_ms"].items()}, "e2e %.3e"%d["e2e"]["value"], "pop", d.get("population") and round(d["population"]["evals_per_s"]))
PY
done
python bench.py --steps 10 --warmup 3 --no-cpu --no-trials --workload c3 --candidates 296 > gpurun_out/b12_c3_296.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/b12_c3_296.json").read().strip().splitlines()[-1])
print("c3 K=296", "pts/s %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "roof %.3f"%d["roofline"]["frac"])
PY
