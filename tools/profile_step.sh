#!/bin/bash
# Launch list + full-set capture of one steady-state step of the default bench (1 GPU).  Run under gpurun:
#   bash tools/profile_step.sh <tag>
# The .ncu-rep files are summarised ON THE BOX (tools/ncu_summary.py -> CSV) and removed, so that what travels back in
# gpurun_out/ stays small: launches_<tag>.csv, ncu_<tag>_kernels.csv, ncu_<tag>_top_source.csv.
tag=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu_launch_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k 'regex:edge_bwd_main_kernel|edge_fwd_kernel|gemm_tc_kernel|gemm_pair_kernel|edge_bwd_rowdot_kernel|edge_max_kernel|scores_bwd_partial' -s 100 -c 36 \
    -f -o /tmp/prof_${tag} $CMD > gpurun_out/ncu_full_${tag}.log 2>&1
echo "profile exit $?"
python tools/ncu_summary.py /tmp/prof_${tag}.ncu-rep gpurun_out/ncu_${tag}_kernels.csv > /dev/null
# per-instruction page of the dominant kernel (fused source-major backward pass, hidden-layer shape)
ncu -i /tmp/prof_${tag}.ncu-rep --page source --csv -k 'regex:edge_bwd_main_kernel.*bool.0, .bool.1, .bool.1' -c 1 > gpurun_out/ncu_${tag}_top_source.csv 2>/dev/null
ls -la /tmp/prof_${tag}.ncu-rep gpurun_out | tail -12
tail -2 gpurun_out/ncu_full_${tag}.log | cut -c1-300
