#!/bin/bash
# round 2, call N: fp16 candidate copy (mode 4): parity tests, then the bench at 1 GPU in both modes
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "h16 or ivf_search_bit_exact or ivf_add or tensor_core_probe or candidate_path" > gpurun_out/n_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/n_tests.log
tail -25 gpurun_out/n_tests.log
timeout 900 python bench.py --mode 4 --no-kmeans --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/n_bench_m4.json 2> gpurun_out/n_bench_m4.err
echo "bench rc=$?"
tail -c 2500 gpurun_out/n_bench_m4.json
tail -5 gpurun_out/n_bench_m4.err
