set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --candidates 512"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_c2.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 4 -c 2 -o gpurun_out/r1_prof_small $CMD > gpurun_out/ncu2.log 2>&1
CMD3="python bench.py --steps 2 --warmup 1 --no-cpu --no-trials --workload c3 --candidates 32"
$CMD3 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_phase_staged -s 1 -c 2 -o gpurun_out/r1_prof_staged $CMD3 > gpurun_out/ncu3.log 2>&1
cat gpurun_out/plain3.log | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu --no-trials --workload c3 > gpurun_out/bench_c3.json 2>gpurun_out/bench_c3.err; tail -1 gpurun_out/bench_c3.json
python bench.py --steps 5 --warmup 3 --no-cpu --no-trials --workload c5 > gpurun_out/bench_c5.json 2>gpurun_out/bench_c5.err; tail -1 gpurun_out/bench_c5.json
python bench.py --steps 5 --warmup 3 --no-cpu --no-trials --workload c1 > gpurun_out/bench_c1.json 2>gpurun_out/bench_c1.err; tail -1 gpurun_out/bench_c1.json
ls -la gpurun_out
