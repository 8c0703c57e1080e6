timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in c2 c1 c5; do
  python bench.py --steps 5 --warmup 3 --no-cpu --no-trials --workload $w > gpurun_out/q_$w.json 2>gpurun_out/q_$w.err || tail -5 gpurun_out/q_$w.err
done
for ch in 1 2 4 8 16; do PYXRD_B200_CHUNKS=$ch python bench.py --steps 5 --warmup 3 --no-cpu --no-trials --workload c2 > gpurun_out/qc${ch}_c2.json 2>gpurun_out/qc${ch}_c2.err; done
python - <<'PY'
import json
for n in ['q_c2','q_c1','q_c5','qc1_c2','qc2_c2','qc4_c2','qc8_c2','qc16_c2']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
        print(n, 'pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.2f/%.2f=%.3f'%(d['roofline']['achieved'], d['roofline']['peak'], d['roofline']['frac']), {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'], 'e2e ms %.3f'%d['e2e']['ms_per_step'])
    except Exception as e: print(n, 'failed', e)
PY
