#!/bin/bash
set -u
T=${1:-r2y}
: > gpurun_out/${T}_ab.txt
for subs in 131072 262144; do
for rep in 1 2; do
for spw in 8 7 6 5 4; do echo "subs=$subs spw=$spw" | tee -a gpurun_out/${T}_ab.txt; AB_SUBS5=$subs CPBUS_SUBS_PER_WARP=$spw AB_REPS=1 AB_CONFIGS=config5 python scripts/ab_bench.py ab/p6_default.so 2>&1 | tee -a gpurun_out/${T}_ab.txt; done
done
done
