#!/bin/bash
# round 2, call D (2 GPUs): vers_comm / vers_sharded_* parity, then the bench at N=2
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/d_mgpu2.log 2>&1
echo "mgpu rc=$?" >> gpurun_out/d_mgpu2.log
tail -15 gpurun_out/d_mgpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/d_bench_n2.json
tail -5 gpurun_out/d_bench_n2.err
