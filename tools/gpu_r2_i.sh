#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_flat_kernel -s 2 -c 1 -o gpurun_out/prof_tcflat_r2 -f \
  python bench.py --workload flat --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/i_ncu.log 2>&1
tail -3 gpurun_out/i_ncu.log
