timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench14.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench14.json').read().strip().splitlines()[-1])
print('pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], d['stage_ms'], 'e2e %.3e'%d['e2e']['value'], 'pop ms', d['population']['ms_per_step'])
PY
