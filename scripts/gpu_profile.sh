#!/bin/bash
# Round profile pass (run under gpurun): bench lines, ncu launch lists and one full capture per hot kernel.
# (each bench step launches the table-driven classic instance, then its per-member-coefficient twin, which is an empty
# early-out on the uniform C4 ensemble: -s 2 lands on the table-driven launch of the second step)
# usage: scripts/gpu_profile.sh <tag>     outputs -> gpurun_out/<tag>_*
tag=${1:-r1}
O=gpurun_out
CL="python bench.py --members 16384 --years 5 --steps 2 --warmup 1 --no-e2e --no-cpu"
MZ="python bench.py --workload miz --members 8192 --years 1 --steps 2 --warmup 1 --no-e2e --no-cpu"
set -x
timeout 200 $CL > $O/${tag}_classic_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${tag}_classic_launches.csv $CL > $O/${tag}_classic_ncu1.log 2>&1
timeout 200 $CL > /dev/null 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:classic_uniform -s 2 -c 1 -f -o $O/${tag}_classic_full $CL > $O/${tag}_classic_ncu2.log 2>&1
timeout 200 $MZ > $O/${tag}_miz_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${tag}_miz_launches.csv $MZ > $O/${tag}_miz_ncu1.log 2>&1
timeout 200 $MZ > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:miz_ -s 1 -c 1 -f -o $O/${tag}_miz_full $MZ > $O/${tag}_miz_ncu2.log 2>&1
tail -2 $O/${tag}_classic_plain.log $O/${tag}_miz_plain.log
