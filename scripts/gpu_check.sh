#!/bin/bash
# closing check on one GPU: whole -m gpu suite + the default bench
set -u
T=${1:-r3d}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 2>&1 | tail -n 3 | tee gpurun_out/${T}_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/${T}_bench_n1.json
