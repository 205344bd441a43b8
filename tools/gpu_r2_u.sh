#!/bin/bash
# round 2, call U: ncu full captures of the small latency-bound kernels of a step (rerank, candidate merge, probe select,
# probe scan) on the per-rank-of-8 emulation, to see what their time is made of
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --rows 1250000 --nlist 512 --nprobe 4 --n-centers 8192 --no-kmeans --no-cpu-baseline --no-spotcheck --recall-queries 0 --no-graph --steps 2 --warmup 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rerank_certify|cand_merge|probe_select" -s 12 -c 4 -o gpurun_out/prof_small_kernels -f $B > gpurun_out/u_ncu.log 2>&1
echo "ncu rc=$?"
B2="python bench.py --no-kmeans --no-cpu-baseline --no-spotcheck --recall-queries 0 --no-graph --steps 2 --warmup 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rerank_certify|cand_merge|probe_select" -s 12 -c 4 -o gpurun_out/prof_small_kernels_full -f $B2 > gpurun_out/u_ncu2.log 2>&1
echo "ncu2 rc=$?"
