timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pairs.py tests/test_gpu_stream.py -q -p no:cacheprovider --timeout 600 2>&1 | tail -n 5 | tee gpurun_out/r2l_ab.txt
AB_REPS=2 AB_CONFIGS=config5,config2,config3 python scripts/ab_bench.py ab/r0_noruns.so ab/r1_runs.so 2>&1 | tee -a gpurun_out/r2l_ab.txt
for spw in 4 16 32; do echo "spw=$spw" | tee -a gpurun_out/r2l_ab.txt; CPBUS_SUBS_PER_WARP=$spw AB_REPS=1 AB_CONFIGS=config5 python scripts/ab_bench.py ab/r1_runs.so 2>&1 | tee -a gpurun_out/r2l_ab.txt; done
