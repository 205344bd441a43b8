#!/bin/bash
# round 2, call Z2: per-lane arrives + graph-replayed host calls: whole GPU suite, racecheck subset again, bench
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/z2_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/z2_gpu_tests.log
tail -6 gpurun_out/z2_gpu_tests.log
timeout 900 python bench.py --no-kmeans --steps 20 --warmup 5 > gpurun_out/z2_bench.json 2> gpurun_out/z2_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/z2_bench.json') if l.startswith('{')][-1])
r=d['roofline']
print('full', d['value'], 'step', round(d['ms_per_step'],4), 'eager', round(d['eager_ms_per_step'],4), 'scan', round(r['avg_launch_ms'],4), 'e2e', d['e2e']['value'], d['e2e']['ids_match_device_path'], d['parity_spotcheck']['ids_equal_oracle'], d['parity_spotcheck']['distance_bits_equal_oracle'])
PY
for w in flat kmeans; do
timeout 900 python bench.py --workload $w --no-cpu-baseline --steps 5 --warmup 2 > gpurun_out/z2_$w.json 2> gpurun_out/z2_$w.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/z2_$w.json') if l.startswith('{')][-1])
print('$w', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'])
PY
done
RSUB='ivf_search_bit_exact and (10-4 or 1-16) or h16_candidate_copy_bit_exact or assign_f16 or small_batch_streaming or pair_distances or candidate_path_falls_back or (assign_tf32_first and 5000) or (kmeans_fit_and_cost and 300) or flat_search_tensor_core_path'
timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$RSUB" > gpurun_out/z2_san_racecheck.log 2>&1
echo "sanitizer racecheck rc=$?"
tail -12 gpurun_out/z2_san_racecheck.log | cut -c1-300
