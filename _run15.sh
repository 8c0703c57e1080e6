for k in 1024 4096 16384 65536; do
python bench.py --steps 3 --warmup 3 --no-cpu --workload c1 --candidates 64 --trials-workload c5 --trials-candidates $k > gpurun_out/c5_trials_$k.json 2> gpurun_out/c5_trials_$k.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c5_trials_$k.json").read().strip().splitlines()[-1]); t=d["trials"]
    print("c5 K=$k", "trials/s %.0f"%t["value"], "ms %.2f"%t["ms_per_step"], {a:round(b,3) for a,b in t["stage_ms"].items()}, "pts/s %.3e"%t["pattern_pts_per_s"], "pop", t["population"] and round(t["population"]["value"]))
except Exception as e: print("K=$k failed", e, open("gpurun_out/c5_trials_$k.err").read()[-400:])
PY
done
