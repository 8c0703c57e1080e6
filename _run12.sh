timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -c 300 gpurun_out/bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench8.json').read().strip().splitlines()[-1])
print('pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.3f'%d['roofline']['frac'], d['stage_ms'], 'e2e %.3e'%d['e2e']['value'], 'pack_s', d['host_pack_s'], d.get('cpu_affinity'))
print('population', d.get('population'))
t=d['trials']; print('trials/s %.1f e2e %.1f'%(t['value'], t['e2e']['value']), t['stage_ms'], t.get('population'))
PY
