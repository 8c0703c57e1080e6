timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c5 c4; do
python bench.py --steps 10 --warmup 3 --no-cpu --no-trials --workload $w > gpurun_out/b13_$w.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/b13_$w.json").read().strip().splitlines()[-1])
print("$w", "pts/s %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "roof %.3f"%d["roofline"]["frac"], {k:round(v,3) for k,v in d["stage_ms"].items()}, "e2e %.3e"%d["e2e"]["value"])
PY
done
python tools/population_sweep.py c5 1024 4096 16384 65536 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 1 -c 1 -o gpurun_out/r1f_small82db python bench.py --steps 2 --warmup 1 --no-cpu --no-trials --workload c5 > gpurun_out/ncu_f1.log 2>&1
