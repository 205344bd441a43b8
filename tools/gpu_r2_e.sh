#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "flat" > gpurun_out/e_flat_tests.log 2>&1
echo "flat tests rc=$?" >> gpurun_out/e_flat_tests.log
tail -8 gpurun_out/e_flat_tests.log
for nq in 1 2 4 8; do
  timeout 300 python bench.py --workload flat --nq $nq --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/e_flat_nq$nq.json 2> gpurun_out/e_flat_nq$nq.err
  tail -c 600 gpurun_out/e_flat_nq$nq.json
done
timeout 300 python bench.py --workload flat --nq 1 --flat-dim 768 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/e_flat_nq1_d768.json 2>> gpurun_out/e_flat_nq1.err
timeout 300 python bench.py --workload lsh --steps 10 --warmup 3 > gpurun_out/e_lsh.json 2> gpurun_out/e_lsh.err
tail -c 1500 gpurun_out/e_lsh.json
