#!/bin/bash
# final multi-GPU evidence (run under `gpurun --gpus 8`): sharded-bus parity at every G <= 8, then the bench at 8/4/2/1 + reference arm
set -u
mkdir -p gpurun_out
T=${1:-r3b}; N=${2:-8}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${T}_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 900 -p no:cacheprovider -rs > gpurun_out/${T}_pytest_multi.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest_multi.txt
bash scripts/gpu_bench8.sh $T $N
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29480 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_ref_n${N}.json 2> gpurun_out/${T}_ref_n${N}.err; echo "ref N=$N rc=$?"
