# ncu recipe of the round (run under gpurun from the repository root): launch list of the default bench command, then
# one --set full capture per kernel family.  The reports are summarised on the box (tools/ncu_summary.py) and removed:
# gpurun copies back at most 64 MiB.  Copy gpurun_out/<tag>_*_summary.txt and <tag>_launches_c2.csv into profiles/.
TAG=${1:-r2m}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_c2.csv $CMD > gpurun_out/ncu_${TAG}0.log 2>&1
capture() {   # name, kernel regex, skip, count, extra bench flags
    ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/${TAG}_$1 $CMD $5 > gpurun_out/ncu_${TAG}_$1.log 2>&1
    python tools/ncu_summary.py /tmp/${TAG}_$1.ncu-rep > gpurun_out/${TAG}_$1_summary.txt 2>&1
    rm -f /tmp/${TAG}_$1.ncu-rep
}
capture small22 k_phase_small 2 1 --no-trials
capture optimize k_mixture_optimize 1 1 ""
capture prologue "k_expand_population|k_phase_planes_image|k_phase_prologue|k_probabilities" 4 4 --no-trials
capture stream "k_wavelength_pass|k_mixture_residual|k_patterns_residual|k_residual_finish_mean" 2 2 --no-trials
ls -la gpurun_out/${TAG}_*
