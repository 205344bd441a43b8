#!/bin/bash
# round 2, call V: pipelined rerank + candidate merge: parity tests, launch list of a step (1 GPU full size + per-rank-of-8
# emulation), bench
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ivf or flat_search or lsh_search" > gpurun_out/v_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/v_tests.log
tail -5 gpurun_out/v_tests.log
timeout 900 python bench.py --no-kmeans --no-cpu-baseline --steps 20 --warmup 5 --trace 2 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
mv gpurun_out/trace_n1.txt gpurun_out/v_trace_full.txt
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/v_bench.json') if l.startswith('{')][-1])
r=d['roofline']
print('full', d['value'], 'step', round(d['ms_per_step'],4), 'eager', round(d['eager_ms_per_step'],4), 'scan', round(r['avg_launch_ms'],4), r['family_ms_per_step'], d['parity_spotcheck']['ids_equal_oracle'], d['parity_spotcheck']['distance_bits_equal_oracle'], d['e2e']['value'])
PY
B="python bench.py --rows 1250000 --nlist 512 --nprobe 4 --n-centers 8192 --no-kmeans --no-cpu-baseline --no-spotcheck --recall-queries 0 --steps 20 --warmup 5"
timeout 300 $B --trace 2 > gpurun_out/v_emul.json 2> gpurun_out/v_emul.err
mv gpurun_out/trace_n1.txt gpurun_out/v_trace_emul.txt
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/v_emul.json') if l.startswith('{')][-1])
r=d['roofline']
print('emul', 'step', round(d['ms_per_step'],4), 'eager', round(d['eager_ms_per_step'],4), 'scan', round(r['avg_launch_ms'],4), r['family_ms_per_step'])
PY
