#!/bin/bash
# round 2, call O: whole GPU suite with the fp16 candidate path, default bench, ncu full capture of the fp16 scan + launch list
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/o_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/o_gpu_tests.log
tail -6 gpurun_out/o_gpu_tests.log
timeout 1500 python bench.py > gpurun_out/o_bench_n1.json 2> gpurun_out/o_bench_n1.err
echo "bench rc=$?"
head -c 600 gpurun_out/o_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/o_launches.csv \
  python bench.py --no-cpu-baseline --no-kmeans --no-spotcheck --no-graph --steps 2 --warmup 1 --recall-queries 0 > gpurun_out/o_bench_under_ncu.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_list_scan_kernel -s 7 -c 1 -o gpurun_out/prof_tc_h16 -f \
  python bench.py --no-cpu-baseline --no-kmeans --no-spotcheck --no-graph --steps 2 --warmup 1 --recall-queries 0 > gpurun_out/o_ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/o_ncu_full.log
