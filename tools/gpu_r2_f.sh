#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "flat" > gpurun_out/f_flat_tests.log 2>&1
echo "flat tests rc=$?" >> gpurun_out/f_flat_tests.log
tail -8 gpurun_out/f_flat_tests.log
for nq in 1 2 4 8; do
  timeout 300 python bench.py --workload flat --nq $nq --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/f_flat_nq$nq.json 2> gpurun_out/f_flat_nq$nq.err
done
timeout 300 python bench.py --workload flat --nq 1 --flat-dim 768 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/f_flat_nq1_d768.json 2>> gpurun_out/f_flat_nq1.err
timeout 300 python bench.py --workload flat --nq 1 --flat-dim 128 --flat-rows 4000000 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/f_flat_nq1_d128.json 2>> gpurun_out/f_flat_nq1.err
