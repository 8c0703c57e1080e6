set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --candidates 512"
$CMD > gpurun_out/plain_b.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 4 -c 1 -o gpurun_out/r1b_small22 $CMD > gpurun_out/ncu_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mixture_optimize -s 1 -c 1 -o gpurun_out/r1b_optimize $CMD > gpurun_out/ncu_b2.log 2>&1
CMD5="python bench.py --steps 2 --warmup 1 --no-cpu --no-trials --workload c5 --candidates 256"
$CMD5 > gpurun_out/plain_b5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 1 -c 1 -o gpurun_out/r1b_small82 $CMD5 > gpurun_out/ncu_b3.log 2>&1
ls -la gpurun_out
