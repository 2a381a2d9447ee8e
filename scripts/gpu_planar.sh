#!/bin/bash
# planar staging: full GPU suite on the default (planar) library, then same-box A/B against the record-major build
set -u
T=${1:-r2v}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x 2>&1 | tail -n 8 | tee gpurun_out/${T}_pytest.txt
AB_REPS=2 AB_CONFIGS=config5,config2,config3 python scripts/ab_bench.py ab/p0_aos.so ab/p1_planar.so ab/p2_planar_unroll2.so ab/p3_planar_noidxpf.so 2>&1 | tee gpurun_out/${T}_ab.txt
for spw in 4 6 12; do echo "spw=$spw" | tee -a gpurun_out/${T}_ab.txt; CPBUS_SUBS_PER_WARP=$spw AB_REPS=1 AB_CONFIGS=config5 python scripts/ab_bench.py ab/p1_planar.so 2>&1 | tee -a gpurun_out/${T}_ab.txt; done
