#!/bin/bash
# LSH load/save tests + C++ host mirror, then compute-sanitizer (memcheck / synccheck / racecheck) over small-shape tests
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lsh or cpp_host or comm_world1" > gpurun_out/l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/l_tests.log
tail -15 gpurun_out/l_tests.log
SUB='assign_tf32_first or kmeans_fit_and_cost or ivf_search_bit_exact or query_block_kernel_bit_exact or small_batch_streaming or tensor_core_probe_bit_exact or lsh_search_identical or ivf_add_then_search or update_centroids'
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$SUB" > gpurun_out/san_$tool.log 2>&1
  echo "sanitizer $tool rc=$?" >> gpurun_out/san_$tool.log
  tail -6 gpurun_out/san_$tool.log
done
