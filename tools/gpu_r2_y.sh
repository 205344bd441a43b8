#!/bin/bash
# round 2, call Y: fp16 assign kernel as default: k-means parity tests, C5 bench (full, with CPU leg), ncu full capture on a
# 2M-row slice
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "assign or kmeans or ivf_build" > gpurun_out/y_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/y_tests.log
tail -4 gpurun_out/y_tests.log
timeout 900 python bench.py --workload kmeans --steps 20 --warmup 2 > gpurun_out/y_kmeans.json 2> gpurun_out/y_kmeans.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/y_kmeans.json') if l.startswith('{')][-1])
r=d['roofline']
print('km', d['value'], 'ms/iter', round(d['ms_per_step'],2), 'assign ms', round(r['avg_launch_ms'],2), 'frac', round(r['frac'],3), 'redo', r.get('exact_redo_ms_per_pass'), d.get('uncertified_rows_last_pass'), d['cpu_baseline'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_assign1 -s 1 -c 1 -o gpurun_out/prof_assign1_f16 -f python bench.py --workload kmeans --km-rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/y_ncu.log 2>&1
echo "ncu rc=$?"
