#!/bin/bash
# round 2, call A: new k-means kernel parity + first timing, then the whole GPU suite
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "assign or kmeans" > gpurun_out/a_km_tests.log 2>&1
echo "km tests rc=$?" >> gpurun_out/a_km_tests.log
tail -5 gpurun_out/a_km_tests.log
timeout 600 python bench.py --workload kmeans --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/a_km_bench.json 2> gpurun_out/a_km_bench.err
tail -c 1500 gpurun_out/a_km_bench.json
timeout 300 python bench.py --workload kmeans --steps 3 --warmup 1 --no-cpu-baseline --km-mode 2 --km-rows 10000000 > gpurun_out/a_km_bench_m2.json 2>> gpurun_out/a_km_bench.err
timeout 300 python bench.py --workload kmeans --steps 3 --warmup 1 --no-cpu-baseline --km-mode 0 --km-rows 10000000 > gpurun_out/a_km_bench_m0.json 2>> gpurun_out/a_km_bench.err
for nq in 1 8; do
  timeout 300 python bench.py --workload flat --nq $nq --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/a_flat_nq$nq.json 2> gpurun_out/a_flat_nq$nq.err
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/a_gpu_tests.log
tail -15 gpurun_out/a_gpu_tests.log
