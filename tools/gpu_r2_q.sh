#!/bin/bash
# round 2, call Q: regrouped pair-distance kernel (parity + bench), small-batch scan with per-lane arrives (parity + nq=1/8
# bench + racecheck of just that kernel's tests)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pair_distances or cpp_host or small_batch or flat" > gpurun_out/q_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/q_tests.log
tail -8 gpurun_out/q_tests.log
timeout 600 python bench.py --workload hnswdist --steps 10 --warmup 3 > gpurun_out/q_hnswdist.json 2> gpurun_out/q_hnswdist.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/q_hnswdist.json') if l.startswith('{')][-1])
print('hnswdist', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_spotcheck'], d['e2e']['value'])
PY
for nq in 1 8; do
  timeout 300 python bench.py --workload flat --nq $nq --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q_flat_nq$nq.json 2> gpurun_out/q_flat_nq$nq.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/q_flat_nq$nq.json') if l.startswith('{')][-1])
print('flat nq$nq', d['value'], d['ms_per_step'], d['roofline']['frac'])
PY
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "small_batch_streaming" > gpurun_out/q_racecheck_stream.log 2>&1
echo "racecheck rc=$?"
tail -12 gpurun_out/q_racecheck_stream.log
