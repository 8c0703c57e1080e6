# ncu recipe of the round (run under gpurun from the repository root): launch list of the default bench command, then
# one --set full capture per kernel family; summaries: python tools/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/<name>_summary.txt
CMD="python bench.py --steps 2 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain_e.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1e_launches_c2.csv $CMD > gpurun_out/ncu_e0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 2 -c 1 -o gpurun_out/r1e_small22 $CMD --no-trials > gpurun_out/ncu_e1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mixture_optimize -s 1 -c 1 -o gpurun_out/r1e_optimize $CMD > gpurun_out/ncu_e2.log 2>&1
ncu --set full --clock-control none -k regex:"k_expand_population|k_phase_planes|k_phase_prologue|k_probabilities" -s 4 -c 4 -o gpurun_out/r1e_prologue $CMD --no-trials > gpurun_out/ncu_e3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
