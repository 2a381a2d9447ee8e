#!/bin/bash
set -u
T=${1:-r3c}
: > gpurun_out/${T}_ab.txt
for subs in 131072 262144 1048576; do
for rep in 1 2; do
for hv in 0 1; do echo "subs=$subs heavy_first=$hv" | tee -a gpurun_out/${T}_ab.txt; AB_SUBS5=$subs CPBUS_ORDER_HEAVY=$hv AB_REPS=1 AB_CONFIGS=config5 python scripts/ab_bench.py ab/p7_heavy.so 2>&1 | tee -a gpurun_out/${T}_ab.txt; done
done
done
CPBUS_ORDER_HEAVY=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "zipf or config5" 2>&1 | tail -2 | tee -a gpurun_out/${T}_ab.txt
