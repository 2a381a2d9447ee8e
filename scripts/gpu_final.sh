#!/bin/bash
# round 2, final 1-GPU evidence call: whole -m gpu suite, sanitizer on the new paths, ncu of the three configurations with
# the final kernel, the PAIRS build's sanitizer + ncu, the default bench (short and long)
set -u
mkdir -p gpurun_out
T=${1:-r2z}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -rs > gpurun_out/${T}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/${T}_pytest.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_stream.py tests/test_gpu_lossless.py -q -p no:cacheprovider \
  -k 'not timeout and not ephemeral' > gpurun_out/${T}_memcheck_stream.txt 2>&1; tail -n 4 gpurun_out/${T}_memcheck_stream.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_stream.py -q -p no:cacheprovider \
  -k 'shards_match and (1-0-0 or 2-1-3) or descriptor_itself' > gpurun_out/${T}_racecheck_stream.txt 2>&1; tail -n 4 gpurun_out/${T}_racecheck_stream.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider \
  -k 'config1 or random_mixed_traces and (1-1 or 2-4) or timer_heavy or zipf' > gpurun_out/${T}_memcheck_parity.txt 2>&1; tail -n 4 gpurun_out/${T}_memcheck_parity.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider \
  -k 'config1 or timer_heavy or zipf_filter_sweep_scaled and 64 or oversize' > gpurun_out/${T}_racecheck_parity.txt 2>&1; tail -n 4 gpurun_out/${T}_racecheck_parity.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --gpus 1 --steps 2000 --warmup 5 --no-cpu > gpurun_out/${T}_bench2000.json 2> gpurun_out/${T}_bench2000.err; echo "bench2000 rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
B="python bench.py --no-e2e --no-cpu --no-verify --no-extras"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout -s 4 -c 1 -o gpurun_out/${T}_fanout_config3 $B --workload config3 --steps 8 --warmup 3 > /dev/null 2>>gpurun_out/prof_err.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout -s 4 -c 1 -o gpurun_out/${T}_fanout_config5 $B --workload config5 --steps 8 --warmup 3 > /dev/null 2>>gpurun_out/prof_err.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout -s 8 -c 1 -o gpurun_out/${T}_fanout_config2 $B --workload config2 --steps 20 --warmup 5 > /dev/null 2>>gpurun_out/prof_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_default.csv python bench.py --no-cpu --steps 20 --warmup 5 > gpurun_out/${T}_bench_under_ncu.log 2>&1
timeout 900 bash scripts/profile_pairs.sh ${T} > gpurun_out/${T}_profile_pairs.log 2>&1
cp containerpilot_b200/libcpbus.so gpurun_out/${T}_libcpbus.so
python scripts/bench_brief.py gpurun_out/${T}_bench.json gpurun_out/${T}_bench2000.json
ls gpurun_out | grep ${T}
