python bench.py > gpurun_out/bench15.json 2> gpurun_out/bench15.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench15.json').read().strip().splitlines()[-1])
print('pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.3f'%d['roofline']['frac'], d['stage_ms'], 'e2e %.3e'%d['e2e']['value'], 'launches', d['gpu_launches'])
print('population', d['population']['ms_per_step'], d['population']['evals_per_s'], 'cpu', d['cpu_baseline']['value'])
t=d['trials']; print('trials/s %.1f e2e %.1f'%(t['value'], t['e2e']['value']), t['stage_ms'], t['population']['value'], t['population']['ms_per_step'], 'cpu', t['cpu_baseline']['value'])
PY
python tools/single_call_latency.py 2>&1 | tail -5
