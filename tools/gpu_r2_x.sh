#!/bin/bash
# round 2, call X: kind::f16 variant of the tf32-first k-means assign kernel: parity, then the C5 build in both modes
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "assign or kmeans" > gpurun_out/x_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/x_tests.log
tail -25 gpurun_out/x_tests.log
for m in 3 0; do
timeout 600 python bench.py --workload kmeans --steps 5 --warmup 1 --no-cpu-baseline --km-mode $m > gpurun_out/x_km_m$m.json 2> gpurun_out/x_km_m$m.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/x_km_m$m.json') if l.startswith('{')][-1])
r=d['roofline']
print('km mode $m', d['value'], 'ms/iter', round(d['ms_per_step'],2), 'assign ms', round(r['avg_launch_ms'],2), 'frac', round(r['frac'],3), 'redo ms', r.get('exact_redo_ms_per_pass'), d.get('uncertified_rows_last_pass'), d.get('parity_spotcheck'))
PY
tail -2 gpurun_out/x_km_m$m.err
done
