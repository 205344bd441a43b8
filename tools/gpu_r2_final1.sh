#!/bin/bash
# round 2, final 1-GPU record: whole GPU suite, the driver's default bench invocation, reference arm, the other workloads,
# launch list and ncu full capture of the dominant kernel
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/f_gpu_tests.log
tail -4 gpurun_out/f_gpu_tests.log
timeout 1500 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
timeout 600 python bench.py --workload flat > gpurun_out/f_flat_1k.json 2> gpurun_out/f_flat_1k.err
for nq in 1 8; do
  timeout 300 python bench.py --workload flat --nq $nq --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_flat_nq$nq.json 2> gpurun_out/f_flat_nq$nq.err
done
timeout 600 python bench.py --workload lsh > gpurun_out/f_lsh.json 2> gpurun_out/f_lsh.err
timeout 600 python bench.py --workload hnswdist > gpurun_out/f_hnswdist.json 2> gpurun_out/f_hnswdist.err
timeout 900 python bench.py --workload kmeans --steps 20 --warmup 2 > gpurun_out/f_kmeans.json 2> gpurun_out/f_kmeans.err
timeout 900 python bench.py --mode 0 --no-kmeans --no-cpu-baseline > gpurun_out/f_bench_n1_mode0.json 2> gpurun_out/f_bench_n1_mode0.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches.csv \
  python bench.py --no-cpu-baseline --no-kmeans --no-spotcheck --no-graph --steps 2 --warmup 1 --recall-queries 0 > gpurun_out/f_bench_under_ncu.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_list_scan_kernel -s 7 -c 1 -o gpurun_out/prof_tc_h16_final -f \
  python bench.py --no-cpu-baseline --no-kmeans --no-spotcheck --no-graph --steps 2 --warmup 1 --recall-queries 0 > gpurun_out/f_ncu_full.log 2>&1
echo "ncu full rc=$?"
