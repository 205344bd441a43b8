#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics (raw page) + a windowed view of the warp-stall samples (source page).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [window]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__inst_executed_pipe_tensor", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:100])
    for k in hdr:
        if any(k == s or k.startswith(s) for s in KEYS):
            print(f"   {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
si, so, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = [r for r in rows[hi + 1:] if len(r) > si]
tot = sum(int(r[si] or 0) for r in data) or 1
print(f"-- warp-stall samples: {tot} over {len(data)} SASS instructions (windows of {win})")
for s in range(0, len(data), win):
    w = data[s:s + win]
    t = sum(int(r[si] or 0) for r in w)
    if t < tot * 0.002:
        continue
    ex = max(int(r[ie] or 0) for r in w)
    ops = set()
    for r in w:
        f = r[so].split()
        if not f:
            continue
        op = f[1] if f[0].startswith("@") and len(f) > 1 else f[0]
        if any(x in op for x in ("SYNCS", "UTC", "LDTM", "STTM", "UTMA", "ATOM", "SHFL", "BAR", "LDS", "STS", "LDG", "STG", "VOTE", "CALL")):
            ops.add(op.split(".")[0] + ("." + op.split(".")[1] if "." in op else ""))
    print(f"   [{s:5d}] {100 * t / tot:5.1f}%  maxexec={ex:<11d} {' '.join(sorted(ops))}")
print("-- top instructions")
for i, r in sorted(enumerate(data), key=lambda x: -int(x[1][si] or 0))[:25]:
    print(f"   [{i:5d}] {100 * int(r[si] or 0) / tot:5.1f}%  exec={r[ie]:>11s}  {r[so][:100]}")
