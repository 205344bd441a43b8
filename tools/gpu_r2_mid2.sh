#!/bin/bash
# round 2: what bounds the candidate scan of a mid-size exhaustive batch: row alignment (300 vs 320 dims) or the
# selection (1 / 4 / 32 queries through the same kernel)?
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for dim in 300 320; do
for nq in 1 4 32; do
  VERS_TC_MIN_NQ=1 timeout 200 python bench.py --workload flat --flat-dim $dim --nq $nq --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/mid2_${dim}_${nq}.json 2> gpurun_out/mid2_${dim}_${nq}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/mid2_${dim}_${nq}.json') if l.startswith('{')][-1])
    print('dim $dim nq $nq ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'GB', d['roofline']['algorithmic_bytes_per_launch']/1e9)
except Exception as e:
    print('dim $dim nq $nq failed', e)
PY
done
done
