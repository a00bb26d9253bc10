#!/bin/bash
# Full-set + per-instruction capture of a few named kernels of one steady-state bench step (1 GPU), summarised on the box.
#   bash tools/profile_kernels.sh <tag> '<kernel regex>' <skip> <count> '<source-page regex 1>' ['<source-page regex 2>' ...]
tag=$1; regex=$2; skip=$3; count=$4; shift 4
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:${regex}" -s ${skip} -c ${count} -f -o /tmp/prof_${tag} $CMD > gpurun_out/ncu_full_${tag}.log 2>&1
echo "profile exit $?"
python tools/ncu_summary.py /tmp/prof_${tag}.ncu-rep gpurun_out/ncu_${tag}_kernels.csv > /dev/null
i=0
for pat in "$@"; do
  i=$((i+1))
  ncu -i /tmp/prof_${tag}.ncu-rep --page source --csv -k "regex:${pat}" -c 1 > gpurun_out/ncu_${tag}_source_${i}.csv 2>/dev/null
done
ls -la gpurun_out | grep ${tag}
