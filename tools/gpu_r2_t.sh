#!/bin/bash
# round 2, call T: scheduler-warp version of the list scan: parity tests, the emulated per-rank 8-GPU scan (item-size
# sweep), and the full-size 1-GPU bench
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ivf or flat_search" > gpurun_out/t_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/t_tests.log
tail -5 gpurun_out/t_tests.log
B="python bench.py --rows 1250000 --nlist 512 --nprobe 4 --n-centers 8192 --no-kmeans --no-cpu-baseline --no-spotcheck --recall-queries 0 --steps 20 --warmup 5"
for cr in 0 512 1024 2048 4096; do
  VERS_TC_CHUNK_ROWS=$cr timeout 300 $B > gpurun_out/t_emul_cr$cr.json 2> gpurun_out/t_emul_cr$cr.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/t_emul_cr$cr.json') if l.startswith('{')][-1])
r=d['roofline']
print('cr=$cr', 'step', round(d['ms_per_step'],4), 'eager', round(d['eager_ms_per_step'],4), 'scan', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), r['family_ms_per_step'])
PY
done
for m in 4 0; do
timeout 900 python bench.py --mode $m --no-kmeans --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/t_bench_m$m.json 2> gpurun_out/t_bench_m$m.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/t_bench_m$m.json') if l.startswith('{')][-1])
r=d['roofline']
print('full mode $m', d['value'], 'step', round(d['ms_per_step'],4), 'scan', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4), d['parity_spotcheck']['ids_equal_oracle'], d['parity_spotcheck']['distance_bits_equal_oracle'], d['e2e']['value'])
PY
done
