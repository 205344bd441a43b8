#!/bin/bash
# round 2, call R (N GPUs): multi-GPU parity check + the bench at N GPUs (default mode: fp16 candidate copy)
set -x
N=${1:-8}
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/r_mgpu$N.log 2>&1
echo "mgpu rc=$?" >> gpurun_out/r_mgpu$N.log
grep -v '^\*\|^$\|OMP_NUM' gpurun_out/r_mgpu$N.log | tail -8 | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --trace 3 > gpurun_out/r_bench_n$N.json 2> gpurun_out/r_bench_n$N.err
echo "bench rc=$?"
mv gpurun_out/trace_n$N.txt gpurun_out/r_trace_n$N.txt
head -c 400 gpurun_out/r_bench_n$N.json
tail -3 gpurun_out/r_bench_n$N.err
