timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -c 200 gpurun_out/bench9.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench9.json').read().strip().splitlines()[-1])
print('pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.3f'%d['roofline']['frac'], d['stage_ms'], 'e2e %.3e'%d['e2e']['value'])
print('population', d['population']['ms_per_step'], d['population']['evals_per_s'])
PY
