#!/bin/bash
# round 2, call P: HNSW distance offload parity + bench, C++ host mirror, racecheck experiment on the small-batch scan
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pair_distances or cpp_host" > gpurun_out/p_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/p_tests.log
tail -25 gpurun_out/p_tests.log
timeout 600 python bench.py --workload hnswdist --steps 10 --warmup 3 > gpurun_out/p_hnswdist.json 2> gpurun_out/p_hnswdist.err
echo "bench rc=$?"
tail -c 1800 gpurun_out/p_hnswdist.json
tail -5 gpurun_out/p_hnswdist.err
