#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "assign or kmeans" > gpurun_out/c_km_tests.log 2>&1
echo "km tests rc=$?" >> gpurun_out/c_km_tests.log
tail -5 gpurun_out/c_km_tests.log
timeout 600 python bench.py --workload kmeans --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/c_km_bench.json 2> gpurun_out/c_km_bench.err
tail -c 900 gpurun_out/c_km_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_assign1 -s 1 -c 1 -o gpurun_out/prof_assign1_r2b -f \
  python bench.py --workload kmeans --km-rows 4000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu.log 2>&1
