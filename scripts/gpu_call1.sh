#!/bin/bash
# round 2, GPU call 1: smoke, the whole -m gpu suite, the default bench (+ reference arm), ncu of configs 3/5, PAIRS sanitizer
set -u
mkdir -p gpurun_out
T=r2a
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/${T}_gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_pytest.txt 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${T}_pytest.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench.err
B="python bench.py --no-e2e --no-cpu --no-verify"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout -s 4 -c 1 -o gpurun_out/${T}_fanout_config3 $B --workload config3 --steps 8 --warmup 3 > /dev/null 2>>gpurun_out/prof_err.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout -s 4 -c 1 -o gpurun_out/${T}_fanout_config5 $B --workload config5 --steps 8 --warmup 3 > /dev/null 2>>gpurun_out/prof_err.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout -s 8 -c 1 -o gpurun_out/${T}_fanout_config2 $B --workload config2 --steps 20 --warmup 5 > /dev/null 2>>gpurun_out/prof_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_default.csv python bench.py --no-cpu --steps 20 --warmup 5 > gpurun_out/${T}_bench_under_ncu.log 2>&1
timeout 600 bash scripts/profile_pairs.sh ${T} > gpurun_out/${T}_profile_pairs.log 2>&1
ls -la gpurun_out | grep ${T}
