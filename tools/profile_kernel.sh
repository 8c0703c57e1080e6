# one ncu --set full capture of one kernel of the default bench command (run under gpurun from the repository root):
#   bash tools/profile_kernel.sh <tag> <name> <kernel regex> <skip> <count> [extra bench flags]
# the report is summarised on the box (tools/ncu_summary.py) and removed: gpurun copies back at most 64 MiB
TAG=$1; NAME=$2; REGEX=$3; SKIP=$4; COUNT=$5; shift 5
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-sweep --no-configs $@"
$CMD > gpurun_out/plain_${TAG}_$NAME.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT -o /tmp/${TAG}_$NAME $CMD > gpurun_out/ncu_${TAG}_$NAME.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_$NAME.ncu-rep --lines 45 > gpurun_out/${TAG}_${NAME}_summary.txt 2>&1
rm -f /tmp/${TAG}_$NAME.ncu-rep
