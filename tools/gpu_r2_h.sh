#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "flat" > gpurun_out/h_flat_tests.log 2>&1
echo "flat tests rc=$?" >> gpurun_out/h_flat_tests.log
tail -25 gpurun_out/h_flat_tests.log
timeout 300 python bench.py --workload flat --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/h_flat_1k.json 2> gpurun_out/h_flat_1k.err
tail -c 1200 gpurun_out/h_flat_1k.json
timeout 300 python bench.py --workload flat --steps 20 --warmup 5 --no-cpu-baseline --flat-mode 2 > gpurun_out/h_flat_1k_m2.json 2>> gpurun_out/h_flat_1k.err
