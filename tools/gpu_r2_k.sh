#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "flat" > gpurun_out/k_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/k_tests.log
tail -25 gpurun_out/k_tests.log
timeout 300 python bench.py --workload flat --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/k_flat_1k.json 2> gpurun_out/k_flat_1k.err
tail -c 700 gpurun_out/k_flat_1k.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_flat_kernel -s 2 -c 1 -o gpurun_out/prof_tcflat_r2b -f \
  python bench.py --workload flat --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/k_ncu.log 2>&1
