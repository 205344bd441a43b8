#!/bin/bash
# round 2: ncu full capture of the candidate scan of a 32-query exhaustive search (1M x 300)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_list_scan_kernel -s 3 -c 1 -o gpurun_out/prof_tc_flat32 -f python bench.py --workload flat --nq 32 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/mid_ncu_full.log 2>&1
echo "ncu rc=$?"
