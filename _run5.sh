for t in 0 1 2 3 4 5; do
for w in c2 c1; do
  PYXRD_B200_TUNE=$t python bench.py --steps 5 --warmup 3 --no-cpu --no-trials --workload $w > gpurun_out/t${t}_$w.json 2>gpurun_out/t${t}_$w.err || tail -5 gpurun_out/t${t}_$w.err
done; done
python - <<'PY'
import json
for t in range(6):
  for w in ['c2','c1']:
    n='t%d_%s'%(t,w)
    try:
        d=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
        print(n, 'pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.2f/%.2f=%.3f'%(d['roofline']['achieved'], d['roofline']['peak'], d['roofline']['frac']), {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(n, 'failed', e)
PY
