#!/bin/bash
# round 2, call M: state-of-the-tree record at 1 GPU: whole GPU suite, the driver's default bench invocation, launch list
# under ncu, and the three other workloads (C2 flat, C3 lsh, C5 k-means) with their CPU legs
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/m_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/m_gpu_tests.log
tail -6 gpurun_out/m_gpu_tests.log
timeout 1500 python bench.py > gpurun_out/m_bench_n1.json 2> gpurun_out/m_bench_n1.err
echo "bench rc=$?"
tail -c 800 gpurun_out/m_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/m_bench_ref.json 2> gpurun_out/m_bench_ref.err
tail -c 600 gpurun_out/m_bench_ref.json
timeout 600 python bench.py --workload flat > gpurun_out/m_flat_1k.json 2> gpurun_out/m_flat_1k.err
for nq in 1 8; do
  timeout 300 python bench.py --workload flat --nq $nq --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/m_flat_nq$nq.json 2> gpurun_out/m_flat_nq$nq.err
done
timeout 600 python bench.py --workload lsh > gpurun_out/m_lsh.json 2> gpurun_out/m_lsh.err
timeout 900 python bench.py --workload kmeans --steps 20 --warmup 2 > gpurun_out/m_kmeans.json 2> gpurun_out/m_kmeans.err
tail -c 600 gpurun_out/m_kmeans.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/m_launches.csv \
  python bench.py --no-cpu-baseline --no-kmeans --no-spotcheck --no-graph --steps 2 --warmup 1 --recall-queries 0 > gpurun_out/m_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
