#!/bin/bash
# round 2, GPU call 2 (1 GPU): whole -m gpu suite on the reworked prologue, default bench, C++ host tests, ncu 3/5/2
set -u
mkdir -p gpurun_out
T=${1:-r2b}
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_pytest.txt 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${T}_pytest.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --gpus 1 --steps 400 --warmup 5 --no-cpu > gpurun_out/${T}_bench400.json 2> gpurun_out/${T}_bench400.err; echo "bench400 rc=$?"
B="python bench.py --no-e2e --no-cpu --no-verify"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout -s 4 -c 1 -o gpurun_out/${T}_fanout_config3 $B --workload config3 --steps 8 --warmup 3 > /dev/null 2>>gpurun_out/prof_err.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout -s 4 -c 1 -o gpurun_out/${T}_fanout_config5 $B --workload config5 --steps 8 --warmup 3 > /dev/null 2>>gpurun_out/prof_err.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout -s 8 -c 1 -o gpurun_out/${T}_fanout_config2 $B --workload config2 --steps 20 --warmup 5 > /dev/null 2>>gpurun_out/prof_err.log
cp containerpilot_b200/libcpbus.so gpurun_out/${T}_libcpbus.so
ls -la gpurun_out | grep ${T}
