#!/bin/bash
# round 2: exhaustive search batches through the tensor-core candidate paths (shared 4-slot bounds) + parity tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for nq in 32 95 96 128 256 1000; do
  timeout 200 python bench.py --workload flat --nq $nq --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/mid_${nq}.json 2> gpurun_out/mid_${nq}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/mid_${nq}.json') if l.startswith('{')][-1])
    print('nq $nq ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'], d['roofline'].get('uncertified_queries_last_step'), d['roofline']['kernel'][:40])
except Exception as e:
    print('nq $nq failed', e)
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "flat or exhaustive" > gpurun_out/mid_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/mid_tests.log
