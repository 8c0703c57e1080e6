# ncu --set full captures of the mixture optimiser in its final state of the round (run under gpurun from the
# repository root): the c4 kernel (7 columns, 256 threads, register team solver) and the c2 kernel (3 columns, 128
# threads, single-lane solver).  Summaries only; the reports stay on the box.
TAG=${1:-r1i}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_mixture_optimize -s 1 -c 1 -o /tmp/${TAG}_opt7 $CMD > gpurun_out/ncu_${TAG}_opt7.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_opt7.ncu-rep > gpurun_out/${TAG}_optimize7_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mixture_optimize -s 1 -c 1 -o /tmp/${TAG}_opt3 $CMD --workload c4 --trials-workload c2 > gpurun_out/ncu_${TAG}_opt3.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_opt3.ncu-rep > gpurun_out/${TAG}_optimize3_summary.txt 2>&1
rm -f /tmp/${TAG}_*.ncu-rep
ls -la gpurun_out/${TAG}_*
