timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -c 150 gpurun_out/bench11.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench11_ref.json 2> gpurun_out/bench11_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench11.json').read().strip().splitlines()[-1])
print('pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.3f'%d['roofline']['frac'], d['stage_ms'], 'e2e %.3e'%d['e2e']['value'], 'launches', d['gpu_launches'])
print('population', d['population']['ms_per_step'], d['population']['evals_per_s'], 'cpu', d['cpu_baseline']['value'])
t=d['trials']; print('trials/s %.1f e2e %.1f'%(t['value'], t['e2e']['value']), t['stage_ms'], t['population']['value'], 'cpu', t['cpu_baseline']['value'])
r=json.loads(open('gpurun_out/bench11_ref.json').read().strip().splitlines()[-1])
print('reference arm', r['value'], r['cpu_baseline']['cores'], 'ratio e2e', d['e2e']['value']/r['value'])
PY
