#!/bin/bash
set -u
T=${1:-r2x}
: > gpurun_out/${T}_ab.txt
for rep in 1 2; do
for blk in -1 131072 262144 524288; do echo "subs=1048576 order_block=$blk" | tee -a gpurun_out/${T}_ab.txt; CPBUS_ORDER_BLOCK=$blk AB_REPS=1 AB_CONFIGS=config5 python scripts/ab_bench.py ab/p5_order_blocks.so 2>&1 | tee -a gpurun_out/${T}_ab.txt; done
done
echo "subs=1048576 default policy" | tee -a gpurun_out/${T}_ab.txt; AB_REPS=1 AB_CONFIGS=config5 python scripts/ab_bench.py ab/p5_order_blocks.so 2>&1 | tee -a gpurun_out/${T}_ab.txt
for subs in 524288 131072; do
  for blk in -1 262144 65536; do echo "subs=$subs order_block=$blk" | tee -a gpurun_out/${T}_ab.txt; AB_SUBS5=$subs CPBUS_ORDER_BLOCK=$blk AB_REPS=1 AB_CONFIGS=config5 python scripts/ab_bench.py ab/p5_order_blocks.so 2>&1 | tee -a gpurun_out/${T}_ab.txt; done
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 600 -x 2>&1 | tail -n 4 | tee gpurun_out/${T}_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/${T}_bench_n1.json
