#!/bin/bash
# same-box A/B of library builds: scripts/gpu_ab.sh <tag> <libA> <libB> ...   (each lib: bench configs 3, 5, 2 twice, interleaved)
set -u
mkdir -p gpurun_out
T=$1; shift
B="python bench.py --no-e2e --no-cpu --steps 60 --warmup 5"
for rep in 1 2; do
  for lib in "$@"; do
    name=$(basename $lib .so)
    for wl in config3 config5 config2; do
      CPBUS_LIB=$PWD/$lib timeout 300 $B --workload $wl > gpurun_out/${T}_${name}_${wl}_${rep}.json 2>> gpurun_out/${T}_err.log
    done
  done
done
# config 5: mailboxes per warp (ORDERED build) with the newest library
last="${@: -1}"
for spw in 4 16 32; do
  CPBUS_SUBS_PER_WARP=$spw CPBUS_LIB=$PWD/$last timeout 300 $B --workload config5 > gpurun_out/${T}_spw${spw}_config5.json 2>> gpurun_out/${T}_err.log
done
python scripts/bench_brief.py gpurun_out/${T}_*.json | grep -v "^==" 
tail -3 gpurun_out/${T}_err.log
