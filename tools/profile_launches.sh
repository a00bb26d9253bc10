#!/bin/bash
# Launch list only (every launch of the default bench with its device time); run under gpurun:
#   bash tools/profile_launches.sh <tag>   ->  gpurun_out/launches_<tag>.csv
tag=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu_launch_${tag}.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_${tag}.csv
