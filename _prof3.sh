CMD="python bench.py --steps 2 --warmup 1 --no-cpu --candidates 256"
ncu --set full --clock-control none --import-source on -k regex:k_mixture_optimize -s 1 -c 1 -o gpurun_out/r1c_optimize $CMD > gpurun_out/ncu_c2.log 2>&1
