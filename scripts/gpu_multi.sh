#!/bin/bash
# multi-GPU call: scripts/gpu_multi.sh <tag> <N>   (run under `gpurun --gpus N`): on-hardware parity of the sharded bus at every
# G <= N (real CUDA-IPC / NVLink peer mappings), then the bench at N GPUs (+ N/2 when N >= 4)
set -u
mkdir -p gpurun_out
T=${1:-r2m}; N=${2:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${T}_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${T}_gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_stream.py tests/test_gpu_lossless.py -m gpu -q --timeout 900 -p no:cacheprovider -rs > gpurun_out/${T}_pytest_multi.txt 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${T}_pytest_multi.txt
run_bench() {   # $1 = gpus
  local g=$1
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) \
    bench.py --gpus $g --steps 20 --warmup 5 > gpurun_out/${T}_bench_n${g}.json 2> gpurun_out/${T}_bench_n${g}.err; echo "bench N=$g rc=$?"
  tail -2 gpurun_out/${T}_bench_n${g}.err | cut -c1-300
}
run_bench $N
if [ $N -ge 4 ]; then run_bench $((N / 2)); fi
if [ "${3:-}" = "with1" ]; then timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench N=1 rc=$?"; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29480 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_ref_n${N}.json 2> gpurun_out/${T}_ref_n${N}.err; echo "ref N=$N rc=$?"
python scripts/bench_brief.py gpurun_out/${T}_bench_n*.json
