#!/bin/bash
# round 2, call Z: compute-sanitizer over the small-shape GPU tests with the final kernels (scheduler-warp list scan, fp16
# candidate copy, fp16 k-means assign, pipelined rerank, fused redo grouping, pair distances)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
SUB='assign_tf32_first or assign_f16 or kmeans_fit_and_cost or ivf_search_bit_exact or h16 or query_block_kernel_bit_exact or small_batch_streaming or tensor_core_probe_bit_exact or lsh_search_identical or ivf_add_then_search or update_centroids or pair_distances or candidate_path_falls_back'
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$SUB" > gpurun_out/z_san_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"
  tail -4 gpurun_out/z_san_$tool.log
done
RSUB='ivf_search_bit_exact and (10-4 or 1-16) or h16_candidate_copy_bit_exact or assign_f16 or small_batch_streaming or pair_distances or candidate_path_falls_back or (assign_tf32_first and 5000)'
timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$RSUB" > gpurun_out/z_san_racecheck.log 2>&1
echo "sanitizer racecheck rc=$?"
tail -12 gpurun_out/z_san_racecheck.log
