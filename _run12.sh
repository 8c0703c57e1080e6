for mb in 16 14 12; do for w in c2; do PYXRD_B200_MINB22=$mb python bench.py --steps 10 --warmup 3 --no-cpu --no-trials --workload $w > gpurun_out/b8_$w.json 2>gpurun_out/b8_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/b8_$w.json").read().strip().splitlines()[-1])
print("$w minb $mb", "pts/s %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "roof %.3f"%d["roofline"]["frac"], {k:round(v,3) for k,v in d["stage_ms"].items()}, "e2e %.3e"%d["e2e"]["value"])
PY
done; done
for w in c1 c4; do python bench.py --steps 10 --warmup 3 --no-cpu --no-trials --workload $w > gpurun_out/b8_$w.json 2>gpurun_out/b8_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/b8_$w.json").read().strip().splitlines()[-1])
print("$w", "pts/s %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "roof %.3f"%d["roofline"]["frac"], {k:round(v,3) for k,v in d["stage_ms"].items()}, "e2e %.3e"%d["e2e"]["value"])
PY
done
