#!/bin/bash
# bench at N GPUs (and N/2, N/4) with the final library + the reference arm at N
set -u
T=${1:-r2u}; N=${2:-8}
for g in $N $((N / 2)) $((N / 4)); do
  [ $g -ge 2 ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) \
    bench.py --gpus $g --steps 20 --warmup 5 > gpurun_out/${T}_bench_n${g}.json 2> gpurun_out/${T}_bench_n${g}.err; echo "bench N=$g rc=$?"
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench N=1 rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_ref_n1.json 2> gpurun_out/${T}_ref_n1.err
python scripts/bench_brief.py gpurun_out/${T}_bench_n*.json
