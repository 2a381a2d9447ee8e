AB_REPS=1 python scripts/ab_bench.py ab/y0_round1.so ab/y1_defaults.so ab/y2_occ5.so 2>&1 | tee gpurun_out/r2f_ab.txt
for spw in 2 4 16; do echo "spw=$spw" | tee -a gpurun_out/r2f_ab.txt; CPBUS_SUBS_PER_WARP=$spw AB_REPS=1 AB_CONFIGS=config3,config5 python scripts/ab_bench.py ab/y1_defaults.so 2>&1 | tee -a gpurun_out/r2f_ab.txt; done
echo "pdl=0" | tee -a gpurun_out/r2f_ab.txt; CPBUS_PDL=0 AB_REPS=1 python scripts/ab_bench.py ab/y1_defaults.so 2>&1 | tee -a gpurun_out/r2f_ab.txt
