#!/bin/bash
# round 2, last 1-GPU record after the shared-bound change: whole GPU suite, smoke, the driver's default bench invocation,
# the exhaustive-search workloads, ncu full capture of the query-block kernel
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/g_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/g_gpu_tests.log
tail -4 gpurun_out/g_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/g_smoke.log 2>&1; tail -1 gpurun_out/g_smoke.log
timeout 600 python bench.py > gpurun_out/g_bench_n1.json 2> gpurun_out/g_bench_n1.err
echo "bench rc=$?"
timeout 300 python bench.py --workload flat > gpurun_out/g_flat_1k.json 2> gpurun_out/g_flat_1k.err
echo "flat rc=$?"
for nq in 16 32 64 128; do
  timeout 200 python bench.py --workload flat --nq $nq --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/g_flat_nq$nq.json 2> gpurun_out/g_flat_nq$nq.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_flat_kernel -s 3 -c 1 -o gpurun_out/prof_tc_flat_c2_final -f \
  python bench.py --workload flat --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/g_ncu_full.log 2>&1
echo "ncu full rc=$?"
