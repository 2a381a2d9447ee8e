#!/bin/bash
# Round-2 first call (not run in round 1: GPU budget spent): sanitizer + ncu evidence for the PAIRS build of the fan-out
# kernel (second-level {code, source} filter).  Run under gpurun, 1 GPU:  gpurun --timeout 600 -- 'bash scripts/profile_pairs.sh r02'
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
# 1. memcheck + racecheck over the pair-filter parity tests (small shapes; the filter in shared memory is the new part)
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pairs.py -q -p no:cacheprovider \
  -k 'semantics or (random_traces and (1-0-32 or 3-2-128 or 5-8-512)) or lossless or one_call' > gpurun_out/${TAG}_pairs_memcheck.txt 2>&1
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_pairs.py -q -p no:cacheprovider \
  -k 'semantics or (random_traces and 3-2-128)' > gpurun_out/${TAG}_pairs_racecheck.txt 2>&1
tail -n 4 gpurun_out/${TAG}_pairs_memcheck.txt; tail -n 4 gpurun_out/${TAG}_pairs_racecheck.txt
# 2. launch list + one full capture of a PAIRS launch on the Job-shaped fleet (scripts/diag_pairs.py; the third bus is the PAIRS one)
N_JOBS=32768 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fanout -c 700 --csv \
  --log-file gpurun_out/${TAG}_launches_pairs_fleet.csv python scripts/diag_pairs.py > gpurun_out/${TAG}_diag_pairs_under_ncu.log 2>&1
N_JOBS=32768 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fanout_kernel<2, 1, 1, 0, 1>' -s 30 -c 1 \
  -o gpurun_out/${TAG}_fanout_pairs python scripts/diag_pairs.py > /dev/null 2>>gpurun_out/prof_err.log
ls -la gpurun_out/${TAG}_*
