CMD="python bench.py --steps 2 --warmup 1 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1h_launches_c2.csv $CMD > gpurun_out/ncu_h0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 1 -c 1 -o /tmp/r1h_small82 $CMD --no-trials --workload c5 > gpurun_out/ncu_h1.log 2>&1
python tools/ncu_summary.py /tmp/r1h_small82.ncu-rep > gpurun_out/r1h_small82_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 1 -c 1 -o /tmp/r1h_small273 $CMD --no-trials --workload c3s > gpurun_out/ncu_h2.log 2>&1
python tools/ncu_summary.py /tmp/r1h_small273.ncu-rep > gpurun_out/r1h_small273_summary.txt 2>&1
ls -la gpurun_out/r1h_*
