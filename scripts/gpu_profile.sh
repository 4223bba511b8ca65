#!/bin/bash
# Round profile pass (run under gpurun): ncu launch lists, one full capture per hot kernel, DRAM traffic of the default
# launch shapes.  Every command runs without ncu first (the recipe's rule); numbers printed under ncu are never bench values.
# usage: scripts/gpu_profile.sh <tag>     outputs -> gpurun_out/<tag>_*
tag=${1:-r2}
O=gpurun_out
CL="python bench.py --members 21312 --years 4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra"
MZ="python bench.py --workload miz --members 8192 --years 1 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra"
CLD="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-extra"
MZD="python bench.py --workload miz --members 131072 --years 5 --steps 1 --warmup 0 --no-e2e --no-cpu --no-extra"
set -x
timeout 300 $CL > $O/${tag}_classic_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${tag}_classic_launches.csv $CL > $O/${tag}_classic_ncu1.log 2>&1
timeout 300 $CL > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:classic_uniform_kernel -s 2 -c 1 -f -o $O/${tag}_classic_full $CL > $O/${tag}_classic_ncu2.log 2>&1
timeout 300 $MZ > $O/${tag}_miz_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${tag}_miz_launches.csv $MZ > $O/${tag}_miz_ncu1.log 2>&1
timeout 300 $MZ > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:miz_ -s 1 -c 1 -f -o $O/${tag}_miz_full $MZ > $O/${tag}_miz_ncu2.log 2>&1
# DRAM traffic of one launch of the default shapes (one metrics pass, no replay of the 13 s kernel 40 times)
timeout 300 $CLD > $O/${tag}_classic_default_plain.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:classic_uniform_kernel -c 1 --csv --log-file $O/${tag}_classic_default_traffic.csv $CLD > $O/${tag}_classic_default_ncu.log 2>&1
timeout 300 $MZD > $O/${tag}_miz_default_plain.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:miz_ -c 1 --csv --log-file $O/${tag}_miz_default_traffic.csv $MZD > $O/${tag}_miz_default_ncu.log 2>&1
tail -n 2 $O/${tag}_classic_plain.log $O/${tag}_miz_plain.log
