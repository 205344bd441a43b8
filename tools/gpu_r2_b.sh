#!/bin/bash
# round 2, call B: ncu --set full of tc_assign1_kernel at the C5 shape (4M-row slice)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_assign1 -s 1 -c 1 -o gpurun_out/prof_assign1_r2 -f \
  python bench.py --workload kmeans --km-rows 4000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
ls -la gpurun_out/
