#!/bin/bash
# new defaults (planar staging on the dense builds, 2x unroll) vs the record-major build; then the locality of the mask order
set -u
T=${1:-r2w}
AB_REPS=2 AB_CONFIGS=config5,config2,config3 python scripts/ab_bench.py ab/p0_aos.so ab/p4_new_default.so 2>&1 | tee gpurun_out/${T}_ab.txt
for blk in 4096 16384 65536 262144; do echo "order_block=$blk" | tee -a gpurun_out/${T}_ab.txt; CPBUS_ORDER_BLOCK=$blk AB_REPS=1 AB_CONFIGS=config5 python scripts/ab_bench.py ab/p4_new_default.so 2>&1 | tee -a gpurun_out/${T}_ab.txt; done
for subs in 524288 262144; do
  for blk in 0 16384; do echo "subs=$subs order_block=$blk" | tee -a gpurun_out/${T}_ab.txt; AB_SUBS5=$subs CPBUS_ORDER_BLOCK=$blk AB_REPS=1 AB_CONFIGS=config5 python scripts/ab_bench.py ab/p4_new_default.so 2>&1 | tee -a gpurun_out/${T}_ab.txt; done
done
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x 2>&1 | tail -n 4 | tee gpurun_out/${T}_pytest.txt
