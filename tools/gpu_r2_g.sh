#!/bin/bash
# N GPUs: vers_comm / vers_sharded_* parity, then the bench at N
set -x
N=${1:-2}
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/g_mgpu$N.log 2>&1
echo "mgpu rc=$?" >> gpurun_out/g_mgpu$N.log
grep -v "^\*\|^$\|OMP_NUM" gpurun_out/g_mgpu$N.log | tail -12 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --trace 3 > gpurun_out/g_bench_n$N.json 2> gpurun_out/g_bench_n$N.err
echo "bench rc=$?"
tail -c 600 gpurun_out/g_bench_n$N.json
