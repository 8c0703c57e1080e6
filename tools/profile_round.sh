# ncu recipe of the round (run under gpurun from the repository root): launch list of the default bench command, then
# one --set full capture per kernel family; summaries: python tools/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/<name>_summary.txt
TAG=${1:-r1g}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_c2.csv $CMD > gpurun_out/ncu_${TAG}0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_phase_small -s 2 -c 1 -o gpurun_out/${TAG}_small22 $CMD --no-trials > gpurun_out/ncu_${TAG}1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mixture_optimize -s 1 -c 1 -o gpurun_out/${TAG}_optimize $CMD > gpurun_out/ncu_${TAG}2.log 2>&1
ncu --set full --clock-control none -k regex:"k_expand_population|k_phase_planes|k_phase_prologue|k_probabilities" -s 4 -c 4 -o gpurun_out/${TAG}_prologue $CMD --no-trials > gpurun_out/ncu_${TAG}3.log 2>&1
ncu --set full --clock-control none -k regex:"k_wavelength_pass|k_mixture_residual" -s 4 -c 3 -o gpurun_out/${TAG}_stream $CMD --no-trials > gpurun_out/ncu_${TAG}4.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep
