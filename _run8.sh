timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -c 300 gpurun_out/bench4.err
for w in c1 c3 c5; do python bench.py --steps 5 --warmup 3 --no-cpu --no-trials --workload $w > gpurun_out/bench4_$w.json 2>gpurun_out/bench4_$w.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench4_ref.json 2> gpurun_out/bench4_ref.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches_c2.csv $CMD > gpurun_out/ncu_d0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 2 -c 1 -o gpurun_out/r1d_small22 $CMD --no-trials > gpurun_out/ncu_d1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_phase_mma -s 1 -c 1 -o gpurun_out/r1d_mma27 python bench.py --steps 2 --warmup 1 --no-cpu --no-trials --workload c3 > gpurun_out/ncu_d2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_phase_mma -s 1 -c 1 -o gpurun_out/r1d_mma8 python bench.py --steps 2 --warmup 1 --no-cpu --no-trials --workload c5 > gpurun_out/ncu_d3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 1 -c 1 -o gpurun_out/r1d_small11 python bench.py --steps 2 --warmup 1 --no-cpu --no-trials --workload c1 > gpurun_out/ncu_d4.log 2>&1
ncu --set full --clock-control none -k regex:"k_wavelength_pass|k_mixture_residual" -s 4 -c 3 -o gpurun_out/r1d_stream $CMD --no-trials > gpurun_out/ncu_d5.log 2>&1
python - <<'PY'
import json
for n in ['bench4','bench4_c1','bench4_c3','bench4_c5']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
        print(n, 'pts/s %.3e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'roof %.2f/%.2f=%.3f'%(d['roofline']['achieved'], d['roofline']['peak'], d['roofline']['frac']), {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'])
        if 'trials' in d: t=d['trials']; print('   trials/s %.1f e2e %.1f'%(t['value'], t['e2e']['value']), t['stage_ms'])
    except Exception as e: print(n, 'failed', e)
PY
