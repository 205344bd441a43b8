#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "flat or assign or kmeans" > gpurun_out/j_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j_tests.log
tail -25 gpurun_out/j_tests.log
timeout 300 python bench.py --workload flat --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/j_flat_1k.json 2> gpurun_out/j_flat_1k.err
tail -c 700 gpurun_out/j_flat_1k.json
timeout 600 python bench.py --workload kmeans --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/j_km_bench.json 2> gpurun_out/j_km_bench.err
tail -c 900 gpurun_out/j_km_bench.json
