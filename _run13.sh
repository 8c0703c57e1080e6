timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu > gpurun_out/bench10.json 2> gpurun_out/bench10.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench10.json').read().strip().splitlines()[-1])
print('pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], d['stage_ms'], 'e2e %.3e'%d['e2e']['value'], 'pop ms', d['population']['ms_per_step'])
t=d['trials']; print('trials/s %.1f e2e %.1f'%(t['value'], t['e2e']['value']), t['stage_ms'], 'pop', t['population']['ms_per_step'], t['population']['value'])
PY
