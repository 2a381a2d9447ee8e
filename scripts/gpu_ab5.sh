for h in 0 1; do echo "hints=$h" | tee -a gpurun_out/r2k_ab.txt; CPBUS_HINTS=$h AB_REPS=1 AB_CONFIGS=config3,config5 python scripts/ab_bench.py ab/cur.so 2>&1 | tee -a gpurun_out/r2k_ab.txt; done
echo "auto" | tee -a gpurun_out/r2k_ab.txt; AB_REPS=1 python scripts/ab_bench.py ab/cur.so 2>&1 | tee -a gpurun_out/r2k_ab.txt
