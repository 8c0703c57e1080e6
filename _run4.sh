set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in c2 c1; do
  python bench.py --steps 5 --warmup 3 --no-cpu --no-trials --workload $w > gpurun_out/p_$w.json 2>gpurun_out/p_$w.err || tail -5 gpurun_out/p_$w.err
done
python - <<'PY'
import json
for n in ['p_c2','p_c1']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
        print(n, 'pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.2f/%.2f=%.3f'%(d['roofline']['achieved'], d['roofline']['peak'], d['roofline']['frac']), {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(n, 'failed', e)
PY
