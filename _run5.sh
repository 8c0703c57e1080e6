for t in 0 1 2 3 4 5 6; do
for w in c2 c1; do
  PYXRD_B200_TUNE=$t python bench.py --steps 5 --warmup 3 --no-cpu --no-trials --workload $w > gpurun_out/t${t}_$w.json 2>gpurun_out/t${t}_$w.err || tail -5 gpurun_out/t${t}_$w.err
done; done
python - <<'PY'
import json
for t in range(7):
  for w in ['c2','c1']:
    n='t%d_%s'%(t,w)
    try:
        d=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
        print(n, 'pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.2f/%.2f=%.3f'%(d['roofline']['achieved'], d['roofline']['peak'], d['roofline']['frac']), {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(n, 'failed', e)
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-trials --candidates 512"
ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 2 -c 1 -o gpurun_out/r1c_small22 $CMD > gpurun_out/ncu_c1.log 2>&1
