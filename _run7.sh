timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r_c2.json 2>gpurun_out/r_c2.err || tail -5 gpurun_out/r_c2.err
python - <<'PY'
import json
for n in ['r_c2']:
    d=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
    print(n, 'pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.3f'%(d['roofline']['frac']), {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'], 'e2e ms %.3f'%d['e2e']['ms_per_step'])
    t=d['trials']; print('trials/s %.1f e2e %.1f'%(t['value'], t['e2e']['value']), t['stage_ms'])
PY
