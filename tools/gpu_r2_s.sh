#!/bin/bash
# round 2, call S: what one rank of an 8-GPU step scans, emulated on 1 GPU (1.25M rows, 512 lists, nprobe 4: the same
# 4000 (query, list) pairs over 512 lists of ~2441 rows).  Item-size sweep + ncu full capture of the scan.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --rows 1250000 --nlist 512 --nprobe 4 --n-centers 8192 --no-kmeans --no-cpu-baseline --no-spotcheck --recall-queries 0 --steps 20 --warmup 5"
for cr in 0 512 1024 2048 4096; do
  VERS_TC_CHUNK_ROWS=$cr timeout 300 $B > gpurun_out/s_emul_cr$cr.json 2> gpurun_out/s_emul_cr$cr.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/s_emul_cr$cr.json') if l.startswith('{')][-1])
r=d['roofline']
print('cr=$cr', 'step', round(d['ms_per_step'],4), 'eager', round(d['eager_ms_per_step'],4), 'scan', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), r['family_ms_per_step'], 'rows', r['algorithmic_bytes_per_launch']/1536)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_list_scan_kernel -s 7 -c 1 -o gpurun_out/prof_tc_h16_emul8 -f $B --no-graph --steps 2 --warmup 1 > gpurun_out/s_ncu.log 2>&1
echo "ncu rc=$?"
