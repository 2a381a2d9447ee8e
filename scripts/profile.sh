#!/bin/bash
# Profiling recipe (run under gpurun, 1 GPU).  Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu"
TAG=${1:-r01}
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 300 --csv --log-file gpurun_out/${TAG}_launches_config2.csv $B --steps 200 --warmup 5 > /dev/null 2>>gpurun_out/prof_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 300 --csv --log-file gpurun_out/${TAG}_launches_config2_e2e.csv python bench.py --no-cpu --steps 60 --warmup 3 > /dev/null 2>>gpurun_out/prof_err.log
# the top kernel, full set
ncu --set full --clock-control none --import-source on -k regex:fanout -s 8 -c 2 -o gpurun_out/${TAG}_fanout_config2 $B --steps 20 --warmup 5 > /dev/null 2>>gpurun_out/prof_err.log
ncu --set full --clock-control none --import-source on -k regex:fanout -s 4 -c 1 -o gpurun_out/${TAG}_fanout_config3 $B --workload config3 --steps 8 --warmup 3 > /dev/null 2>>gpurun_out/prof_err.log
ncu --set full --clock-control none --import-source on -k regex:fanout -s 4 -c 1 -o gpurun_out/${TAG}_fanout_config5 $B --workload config5 --steps 8 --warmup 3 > /dev/null 2>>gpurun_out/prof_err.log
tail -3 gpurun_out/prof_err.log
ls -la gpurun_out/${TAG}_*
