#!/bin/bash
# PAIRS build after the triage change: parity, sanitizer, fleet timing (launch list), one full ncu capture of a PAIRS launch
set -u
T=${1:-r2p}
timeout 900 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_parity.py tests/test_gpu_events_api.py tests/test_gpu_lossless.py -q -p no:cacheprovider --timeout 600 2>&1 | tail -n 6 | tee gpurun_out/${T}_pytest.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pairs.py -q -p no:cacheprovider > gpurun_out/${T}_pairs_memcheck.txt 2>&1; tail -n 3 gpurun_out/${T}_pairs_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_pairs.py -q -p no:cacheprovider -k 'semantics or (random_traces and 3-2-128) or fleet' > gpurun_out/${T}_pairs_racecheck.txt 2>&1; tail -n 3 gpurun_out/${T}_pairs_racecheck.txt
N_JOBS=32768 timeout 300 python scripts/diag_pairs.py 2>&1 | tee gpurun_out/${T}_diag_pairs.txt
N_JOBS=32768 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fanout -c 700 --csv --log-file gpurun_out/${T}_launches_pairs_fleet.csv python scripts/diag_pairs.py > /dev/null 2>&1
N_JOBS=32768 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanout --launch-skip 520 -c 1 -o gpurun_out/${T}_fanout_pairs python scripts/diag_pairs.py > /dev/null 2>>gpurun_out/prof_err.log
cp containerpilot_b200/libcpbus.so gpurun_out/${T}_libcpbus.so
ls -la gpurun_out | grep ${T}
