#!/usr/bin/env python
"""bench.py — IVFFlat QPS @ recall@10 on 10M x 768 synthetic vectors (BASELINE.json configs[3]), batch 1k,
nlist 4096, nprobe 32, k 10, rows sharded over N B200s with an NCCL all-gather + merge of the per-GPU top-k.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      the CPU arm: the oracle port of vers' own search path

One "step" = one search_approximate batch (1000 queries) over the whole index.
  value  : QPS with the query batch already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e    : QPS through the reference-facing C-ABI call with HOST (pinned) buffers: H2D of the queries and D2H of
           ids+distances inside the timed region
  roofline: the list-scan kernel's algorithmic bytes (rows of the DISTINCT probed lists x dim x 4 B, per launch)
           over its CUDA-event duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the oracle (kind "port": the Rust reference cannot be compiled in this image),
           all host cores over queries, on a bounded sample (see `sample`)
Synthetic data (include/vers_synth.h), random-init k-means (a few Lloyd iterations, reported as build_s).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED_DATA, SEED_QUERY, SEED_INIT, SEED_CENTERS = 1, 2, 3, 7
CPU_SCALE = 32  # the CPU arms run on rows/32 with nlist/32: same rows per list, same nprobe => same scan work per query


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nlist", type=int, default=4096)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--nq", type=int, default=1000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--n-centers", type=int, default=65536,
                    help="natural clusters of the synthetic data (65536 => ~150 rows each: k-means lists come out "
                         "balanced and every probe opens a full-size list, ~78k rows scanned per query)")
    ap.add_argument("--kmeans-iters", type=int, default=2, help="max Lloyd iterations of the index build")
    ap.add_argument("--reduce", default="chained", choices=["chained", "allreduce"])
    ap.add_argument("--shard-by", default="lists", choices=["lists", "rows"],
                    help="N > 1: every GPU owns whole inverted lists (rows exchanged once after k-means) or keeps its "
                         "row block (1/N of every list)")
    ap.add_argument("--recall-queries", type=int, default=100)
    ap.add_argument("--cpu-queries", type=int, default=256, help="queries in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pairs", type=int, default=1 << 22, help="--workload hnswdist: (query, neighbour) pairs per step")
    ap.add_argument("--workload", default="ivf", choices=["ivf", "kmeans", "flat", "lsh", "hnswdist"],
                    help="ivf (default): the QPS line; kmeans: BASELINE.json configs[4], k-means build seconds on "
                         "50M x 128, 16384 centroids, --steps Lloyd iterations (default 20), rows sharded over the GPUs")
    ap.add_argument("--flat-rows", type=int, default=1_000_000, help="--workload flat: BASELINE.json configs[1]")
    ap.add_argument("--flat-dim", type=int, default=300)
    ap.add_argument("--flat-mode", type=int, default=0, help="0 tensor-core candidate path, 1 exact-order engine")
    ap.add_argument("--km-rows", type=int, default=50_000_000)
    ap.add_argument("--km-dim", type=int, default=128)
    ap.add_argument("--km-clusters", type=int, default=16384)
    ap.add_argument("--km-mode", type=int, default=0, help="0 tcgen05 candidate argmin + certificate (single-MMA fp16 kernel "
                    "for dim <= 128), 1 exact order only, 2 split-precision tcgen05 kernel (3 MMAs per K step)")
    ap.add_argument("--no-graph", action="store_true",
                    help="time eager launches of the step instead of CUDA-graph replays (the default at every N: the "
                         "step has no NCCL call and no host synchronisation in it)")
    ap.add_argument("--no-spotcheck", action="store_true", help="skip the oracle parity spot check of 4 queries")
    ap.add_argument("--no-kmeans", action="store_true",
                    help="skip the kmeans_c5 sub-record (BASELINE.json configs[4] build seconds) of the default run")
    ap.add_argument("--km-cpu-rows", type=int, default=50_000, help="rows of the CPU assign sample (kmeans cpu_baseline)")
    ap.add_argument("--trace", type=int, default=0, help="diagnostic: profile this many extra steps with torch.profiler "
                    "(CUPTI) on rank 0 after the timed region and write the kernel timeline to gpurun_out/")
    ap.add_argument("--mode", type=int, default=4, help="candidate pass: 4 fp16 candidate copy (tcgen05 kind::f16) + "
                    "exact fp32 rerank + certificate (default: BASELINE.json configs[3] '16-bit candidate / fp32 "
                    "rerank'), 0 tcgen05 split-TF32 over the fp32 rows, 1 exact order only, 2 fp32 FMA SIMT, "
                    "3 tcgen05 plain TF32")
    return ap.parse_args()


def committed_traffic(args, ws):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json);
    None when no capture exists for this workload"""
    key = (f"ivf_{args.rows}x{args.dim}_nlist{args.nlist}_nprobe{args.nprobe}_nq{args.nq}_k{args.k}_gpus{ws}"
           f"_mode{args.mode}")
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[key]["dram_bytes_per_launch"], t[key]["source"]
    except Exception:
        return None, None


def committed_traffic_key(key):
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[key]["dram_bytes_per_launch"], t[key]["source"]
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def use_all_host_threads(vo):
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms are meant to use every host core this process
    may run on (the reference's rayon pool defaults to all logical CPUs)"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    vo.set_threads(max(1, n))


def cpu_arm(args, steps, warmup, nq_sample):
    """the reference's CPU search path (oracle port) on the bounded sample; returns (qps, ms_per_step, info)"""
    import oracle as vo

    use_all_host_threads(vo)
    rows_n = max(args.rows // CPU_SCALE, 1000)
    nlist = max(args.nlist // CPU_SCALE, 1)
    nprobe = min(args.nprobe, nlist)
    cores = vo.num_threads()
    rows = vo.synth(SEED_DATA, rows_n, args.dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS)
    q = vo.synth(SEED_QUERY, nq_sample, args.dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS)
    init = vo.init_rows(SEED_INIT, 1, nlist, rows_n)
    t0 = time.perf_counter()
    cents, assign, _ = vo.kmeans_fit(rows, init[0], args.kmeans_iters)
    build_s = time.perf_counter() - t0
    off, lr = vo.ivf_lists(assign, nlist)
    for _ in range(warmup):
        vo.ivf_search(rows, cents, off, lr, q[: max(8, nq_sample // 8)], args.k, nprobe=nprobe)
    t0 = time.perf_counter()
    for _ in range(steps):
        vo.ivf_search(rows, cents, off, lr, q, args.k, nprobe=nprobe)
    dt = time.perf_counter() - t0
    qps = nq_sample * steps / dt
    sample = (f"{nq_sample} queries/step on rows/{CPU_SCALE} ({rows_n}x{args.dim}) with nlist/{CPU_SCALE} ({nlist}) and "
              f"nprobe {nprobe}: the same ~{nprobe * rows_n // nlist} rows scanned per query as the full config "
              f"(nprobe*rows/nlist), centroid probe {CPU_SCALE}x smaller (<5% of a query's work); OpenMP over queries")
    return qps, dt / steps * 1e3, dict(cores=cores, kind="port", sample=sample, cpu_build_s=round(build_s, 2))


def config_dict(args, n_gpus):
    return {"workload": f"IVFFlat search_approximate batch: {args.rows}x{args.dim} fp32 synthetic clustered+normalized, "
                        f"nlist {args.nlist}, nprobe {args.nprobe}, top_k {args.k}, {args.nq}-query batch "
                        f"(BASELINE.json configs[3])",
            "rows": args.rows, "dim": args.dim, "nlist": args.nlist, "nprobe": args.nprobe, "top_k": args.k,
            "batch": args.nq,
            "sharding": ("1 GPU" if n_gpus == 1 else
                         (f"inverted lists balanced over {n_gpus} GPUs (rows exchanged once after k-means)"
                          if args.shard_by == "lists" else f"rows/{n_gpus} per GPU (1/{n_gpus} of every list)") +
                         ", probe split over ranks, probe lists + per-GPU top-k exchanged over NVLink peer memory"),
            "kmeans_iters": args.kmeans_iters, "synthetic_natural_clusters": args.n_centers,
            "candidate_pass": {4: "fp16 candidate copy of the lists (tcgen05 kind::f16), exact fp32 rerank, certificate, "
                                  "exact redo (results bit-identical to the fp32 reference arithmetic)",
                               0: "fp32 rows, split-precision tf32 (tcgen05), exact fp32 rerank, certificate",
                               1: "exact order only", 2: "fp32 FMA SIMT candidates", 3: "plain tf32 candidates"}[args.mode],
            "l2": "per-step scan (>= 1.7 GB per GPU at 8 GPUs, 13.8 GB at 1) is far larger than the 126 MB L2; no flush "
                  "needed"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    qps, ms, info = cpu_arm(args, args.steps, args.warmup, args.cpu_queries)
    line = {"impl": "reference", "metric": "IVFFlat QPS@recall10 (10Mx768, batch 1k)", "value": qps, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": qps, "unit": "queries/s", **info},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def setup_ranks():
    """torch.distributed is plumbing here: it carries the NCCL unique id of the library's own communicator (vers_comm,
    csrc/comm.cu) and the object gathers of the parity spot check.  Every collective of the data path is the library's."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return rank, ws, local_rank


def parity_spotcheck(args, index, ids_dev, d_dev, rank, ws):
    """Outside the timed region, at every N: 4 queries of the timed batch are searched by the CPU oracle
    (ivfflat.rs:153-198 with the nprobe extension) over the rows of the lists it probes — rows REGENERATED from the
    synthetic-data specification, list membership taken from the ranks that own the lists and itself re-checked by
    re-assigning a sample of those rows on the CPU — and must give the ids and distance bits of the timed path."""
    import torch.distributed as dist

    import oracle as vo

    C, dim, k = args.nlist, args.dim, args.k
    npb = min(args.nprobe, C)
    qsel = sorted({0, args.nq // 3, (2 * args.nq) // 3, args.nq - 1})
    q = vo.synth(SEED_QUERY, args.nq, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS)[qsel]
    cents = index.ivf.centroids  # replicated on every rank
    probed = []
    for i in range(len(qsel)):
        dc = np.array([vo.l2sq(q[i], cents[c]) for c in range(C)], np.float32)
        probed.append(np.lexsort((np.arange(C), dc))[:npb])  # stable sort by distance (ivfflat.rs:155-161)
    need = sorted(set(int(c) for p in probed for c in p))
    sizes = index.ivf.list_sizes
    mine = {c: index.ivf.get_list(c) for c in need if sizes[c] > 0}
    parts = [mine]
    if ws > 1:
        parts = [None] * ws
        dist.all_gather_object(parts, mine)
    if rank != 0:
        return None
    lists = {c: np.sort(np.concatenate([p[c] for p in parts if c in p] or [np.empty(0, np.uint64)])) for c in need}
    ok_ids = ok_bits = True
    for i, qi in enumerate(qsel):
        sub_ids = np.concatenate([lists[int(c)] for c in probed[i]])
        sub_assign = np.concatenate([np.full(lists[int(c)].shape[0], int(c), np.uint64) for c in probed[i]])
        order = np.argsort(sub_ids, kind="stable")
        sub_ids, sub_assign = sub_ids[order], sub_assign[order]
        rows = vo.synth_rows(SEED_DATA, sub_ids, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS)
        off, lr = vo.ivf_lists(sub_assign, C)
        oi, od, _ = vo.ivf_search(rows, cents, off, lr, q[i:i + 1], k, nprobe=npb)
        got_ids = sub_ids[oi[0].astype(np.int64)]
        ok_ids = ok_ids and bool(np.array_equal(got_ids, ids_dev[qi]))
        ok_bits = ok_bits and bool(np.array_equal(od[0].view(np.uint32), d_dev[qi].view(np.uint32)))
    rng = np.random.default_rng(11)
    pool_ids = np.concatenate([lists[c] for c in need])
    pool_lists = np.concatenate([np.full(lists[c].shape[0], c, np.uint64) for c in need])
    pick = rng.choice(pool_ids.shape[0], min(256, pool_ids.shape[0]), replace=False)
    prow = vo.synth_rows(SEED_DATA, pool_ids[pick], dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS)
    ok_assign = bool(np.array_equal(vo.assign(prow, cents), pool_lists[pick]))
    return {"queries": len(qsel), "query_rows": [int(x) for x in qsel], "ids_equal_oracle": ok_ids,
            "distance_bits_equal_oracle": ok_bits, "membership_rows_reassigned_by_oracle": int(pick.shape[0]),
            "membership_equal_oracle": ok_assign,
            "how": "oracle search over the probed lists' rows regenerated from include/vers_synth.h; list membership "
                   "from the owning ranks, re-checked by vo.assign on a sample"}


def main_ours(args):
    import ctypes as C

    import torch

    import vers_b200 as vb
    from vers_b200 import _abi
    from vers_b200.sharded import Comm, ShardedIVFFlat, device_view, shard_bounds

    rank, ws, local_rank = setup_ranks()
    dev = torch.device("cuda", local_rank)
    ctx = vb.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    comm = Comm(ctx)  # NCCL communicator + peer-memory exchange buffers of the library (connections made here)

    def barrier():
        comm.barrier()
        torch.cuda.synchronize()

    max_over_ranks = comm.max_over_ranks

    # ---- data + index build (not in the QPS timed region; build seconds reported separately)
    row0, n_local = shard_bounds(args.rows, rank, ws)
    ds = vb.Dataset.synth(ctx, SEED_DATA, n_local, args.dim, kind=1, n_centers=args.n_centers,
                          center_seed=SEED_CENTERS, row0=row0, normalize=True)
    init = vb.synth_init_rows(SEED_INIT, 1, args.nlist, args.rows)[0]
    barrier()
    t0 = time.perf_counter()
    index = ShardedIVFFlat.build(comm, ds, args.nlist, args.kmeans_iters, init, reduce=args.reduce,
                                 shard_by=args.shard_by)
    barrier()
    build_s = max_over_ranks(time.perf_counter() - t0)
    exchange_s = max_over_ranks(comm.last_exchange_s)

    index.ivf.set_mode(args.mode)
    qds = vb.Dataset.synth(ctx, SEED_QUERY, args.nq, args.dim, kind=1, n_centers=args.n_centers,
                           center_seed=SEED_CENTERS, row0=0, normalize=True)
    d_q = device_view(qds.device_ptr, (args.nq, qds.ld))
    h_q = torch.empty((args.nq, qds.ld), dtype=torch.float32).pin_memory()
    h_q.copy_(d_q)
    h_ids = torch.empty((args.nq, args.k), dtype=torch.int64).pin_memory()
    h_d = torch.empty((args.nq, args.k), dtype=torch.float32).pin_memory()
    h_c = torch.empty((args.nq,), dtype=torch.int32).pin_memory()

    # ---- recall@10 against the exhaustive ground truth (flat scan of every row shard + merge by (distance, id))
    nrec = min(args.recall_queries, args.nq)
    recall = None
    if nrec > 0:
        g_ids = torch.empty((nrec, args.k), dtype=torch.int64, device=dev)
        g_d = torch.empty((nrec, args.k), dtype=torch.float32, device=dev)
        g_c = torch.empty((nrec,), dtype=torch.int32, device=dev)
        _abi.check(vb.lib().vers_flat_search_dev(ds.h, C.c_void_p(d_q.data_ptr()), nrec, args.k, 0,
                                                 C.c_void_p(g_ids.data_ptr()), C.c_void_p(g_d.data_ptr()),
                                                 C.c_void_p(g_c.data_ptr())))
        if ws > 1:
            import torch.distributed as dist

            a_ids = torch.empty((ws, nrec, args.k), dtype=torch.int64, device=dev)
            a_d = torch.empty((ws, nrec, args.k), dtype=torch.float32, device=dev)
            torch.cuda.synchronize()
            dist.all_gather_into_tensor(a_ids, g_ids)
            dist.all_gather_into_tensor(a_d, g_d)
            torch.cuda.synchronize()
            _abi.check(vb.lib().vers_topk_merge_dev(ctx.h, C.c_void_p(a_ids.data_ptr()), C.c_void_p(a_d.data_ptr()),
                                                    ws, 0, 0, nrec, args.k, C.c_void_p(g_ids.data_ptr()),
                                                    C.c_void_p(g_d.data_ptr()), C.c_void_p(g_c.data_ptr())))
        ids, _, _ = index.search_dev(d_q[:nrec].contiguous(), args.k, args.nprobe)
        torch.cuda.synchronize()
        gt = g_ids.cpu().numpy()
        got = ids.cpu().numpy()
        recall = float(np.mean([len(set(got[i]) & set(gt[i])) / args.k for i in range(nrec)]))
    ds.close()  # the row-major rows are not needed any more (the index holds its list-major copy)

    # ---- device-timed QPS (queries resident in HBM).  The step (probe -> peer all-gather of the probe lists -> list
    #      scan -> rerank -> peer gather+merge) has no NCCL call and no host synchronisation in it, so at any N it is
    #      captured once into a CUDA graph and the K timed steps are K replays (--no-graph: eager launches).  The
    #      dominant kernel's duration for the roofline comes from K eager steps with per-family CUDA events after.
    def eager_step():
        return index.search_dev(d_q, args.k, args.nprobe)

    graph = None
    if not args.no_graph:
        try:
            graph, _ = index.capture_search(d_q, args.k, args.nprobe)
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {rank}: CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches",
                  file=sys.stderr)
            graph = None
    if ws > 1:  # every rank must take the same path
        ok = -max_over_ranks(-(1.0 if graph is not None else 0.0))
        if ok < 0.5:
            graph = None
    step = graph.replay if graph is not None else eager_step

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    qps = args.nq * args.steps / (dev_ms * 1e-3)
    peer_us = None
    if ws > 1:  # where the fused top-k exchange of the LAST timed step spent its time on this rank (device timestamps)
        t = (C.c_uint64 * 4)()
        _abi.check(vb.lib().vers_debug_peer_times(comm.h, t))
        mine = [(t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[3] - t[2]) / 1e3]
        peer_us = {"publish": max_over_ranks(mine[0]), "wait_for_peers_max": max_over_ranks(mine[1]),
                   "wait_for_peers_min": -max_over_ranks(-mine[1]), "merge": max_over_ranks(mine[2])}

    # eager steps with per-family events: kernel durations for the roofline, launch count
    for _ in range(2):
        eager_step()
    barrier()
    ctx.enable_timing(True)
    launches0 = ctx.launch_count
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for _ in range(args.steps):
        eager_step()
    ee1.record()
    barrier()
    eager_ms = max_over_ranks(ee0.elapsed_time(ee1))
    launches = ctx.launch_count - launches0
    fam = {name: ctx.kernel_ms(getattr(_abi, "KF_" + name.upper())) for name in
           ("list_scan", "probe", "cand_scan", "rerank")}
    ctx.enable_timing(False)
    stats = index.ivf.last_search_stats()
    per_rank_cand = None
    if ws > 1:
        import torch.distributed as dist

        mine = (fam["cand_scan"][0] / max(fam["cand_scan"][1], 1), int(stats["distinct_list_rows"]))
        allp = [None] * ws
        dist.all_gather_object(allp, mine)
        per_rank_cand = {"cand_scan_ms": [round(x[0], 4) for x in allp], "distinct_list_rows": [x[1] for x in allp]}

    if args.trace > 0:
        from torch.profiler import ProfilerActivity, profile

        barrier()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(args.trace):
                index.search_dev(d_q, args.k, args.nprobe)
            torch.cuda.synchronize()
        if rank == 0:
            evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
            evs.sort(key=lambda e: e.time_range.start)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"trace_n{ws}.txt"), "w") as f:
                prev_end = None
                for e in evs:
                    st, en = e.time_range.start, e.time_range.end
                    gap = (st - prev_end) if prev_end is not None else 0.0
                    f.write(f"{st:14.1f} gap {gap:8.1f} dur {en - st:9.1f} us  {e.name[:100]}\n")
                    prev_end = en if prev_end is None else max(prev_end, en)
        barrier()

    # ---- e2e: the reference-facing C-ABI call with HOST (pinned) buffers at every N: vers_sharded_ivf_search copies the
    #      queries in, runs the sharded step, copies ids + distances + counts out and synchronises
    def e2e_step():
        _abi.check(vb.lib().vers_sharded_ivf_search(comm.h, index.ivf.h, h_q.data_ptr(), args.nq, qds.ld, args.k,
                                                    args.nprobe, h_ids.data_ptr(), h_d.data_ptr(), h_c.data_ptr()))

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_qps = args.nq * args.steps / e2e_s
    # the e2e result must equal the device-resident result
    ids_dev, d_dev, _ = index.search_dev(d_q, args.k, args.nprobe)
    torch.cuda.synchronize()
    e2e_matches = bool(torch.equal(ids_dev.cpu(), h_ids))

    # ---- the reference's own call shape: ONE query per call (Index::search_approximate(query, top_k), ivfflat.rs:153),
    #      host buffers in and out.  nprobe 0 = the reference's nearest-list-plus-spill semantics (exact order everywhere),
    #      nprobe = args.nprobe = the batch path's semantics for one query.  Latency-bound: a call streams the centroid
    #      table and one (or nprobe) lists, 20 MB or so; reported as microseconds per call.
    single = None
    if ws == 1:
        single = {}
        one_ids = torch.empty((1, args.k), dtype=torch.int64).pin_memory()
        one_d = torch.empty((1, args.k), dtype=torch.float32).pin_memory()
        one_c = torch.empty((1,), dtype=torch.int32).pin_memory()
        for name, npb in (("reference_semantics_nprobe0", 0), (f"nprobe{args.nprobe}", args.nprobe)):
            def one(i, npb=npb):
                _abi.check(vb.lib().vers_ivf_search(index.ivf.h, h_q.data_ptr() + (i % args.nq) * qds.ld * 4, 1, qds.ld,
                                                    args.k, npb, one_ids.data_ptr(), one_d.data_ptr(), one_c.data_ptr()))
            for i in range(20):
                one(i)
            t0 = time.perf_counter()
            for i in range(200):
                one(i)
            single[name + "_us_per_call"] = (time.perf_counter() - t0) / 200 * 1e6
        single["call"] = "vers_ivf_search(nq = 1) with host buffers, 200 different queries one after the other"

    spot = None
    if not args.no_spotcheck:
        try:
            spot = parity_spotcheck(args, index, ids_dev.cpu().numpy().view(np.uint64), d_dev.cpu().numpy(), rank, ws)
        except Exception as e:  # noqa: BLE001
            spot = {"error": f"{type(e).__name__}: {e}"}

    # ---- roofline of the dominant kernel (list scan)
    peak, peak_src = measured_peaks()
    # bytes the candidate pass must stream: the distinct probed rows once, fp32 or (mode 4) the fp16 candidate copy
    alg_bytes = stats["distinct_list_rows"] * (((args.dim + 7) // 8 * 8) * 2 if args.mode == 4 else args.dim * 4)
    roof = None
    dom = "cand_scan" if fam["cand_scan"][1] else "list_scan"
    dom_ms, dom_n = fam[dom]
    if dom_n:
        avg_ms = dom_ms / dom_n
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
        cand_names = {0: "tc_list_scan_kernel<1> (candidate pass: TMA + tcgen05 kind::tf32 split hi/lo, fp32 rows "
                         "streamed once)",
                      3: "tc_list_scan_kernel<0> (candidate pass: TMA + tcgen05 kind::tf32)",
                      4: "tc_list_scan_kernel<2> (candidate pass: TMA + tcgen05 kind::f16 over the fp16 candidate copy "
                         "of the lists x [q_hi; q_lo], streamed once; exact fp32 rerank + Cauchy-Schwarz certificate)",
                      2: "list_scan_kernel<StreamCfg,1> (candidate pass: fp32 FMA SIMT)"}
        kname = {"cand_scan": cand_names.get(args.mode, "candidate pass"),
                 "list_scan": "list_scan_kernel<NarrowCfg,0> (exact-order scan)"}[dom]
        traffic, traffic_src = committed_traffic(args, ws)
        roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms,
                "kernel_share_of_step": dom_ms / eager_ms if ws == 1 else None,
                "timed_in": "K eager steps right after the timed region (CUDA events around every launch of the family)",
                "family_ms_per_step": {k_: (v[0] / args.steps) for k_, v in fam.items()},
                "pair_rows_per_launch": stats["pair_rows"], "lists_touched": stats["lists_touched"],
                "per_rank": per_rank_cand,
                "uncertified_queries_last_step": stats["uncertified_queries"],
                "uncertified_probe_queries_last_step": stats["uncertified_probe_queries"],
                "reranked_candidates_last_step": stats["reranked"],
                "max_candidate_error_last_step": stats["max_candidate_error"]}

    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline:
        cqps, _, info = cpu_arm(args, 1, 1, args.cpu_queries)
        cpu = {"value": cqps, "unit": "queries/s", **info}

    # ---- the metric's second half, in the same record: k-means build seconds (BASELINE.json configs[4])
    km_rec = None
    if not args.no_kmeans:
        index.ivf.close()
        qds.close()
        torch.cuda.empty_cache()
        try:
            km_rec = kmeans_record(args, comm, ctx, rank, ws, local_rank, iters=20, warm=1,
                                   cpu=(ws == 1 and not args.no_cpu_baseline), sample_clocks=False)
        except Exception as e:  # noqa: BLE001
            km_rec = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {"metric": "IVFFlat QPS@recall10 (10Mx768, batch 1k)", "value": qps, "unit": "queries/s", "n_gpus": ws,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_dict(args, ws), "recall_at_10": recall,
                "build_s": build_s, "build_exchange_s": exchange_s if ws > 1 else None, "kmeans_reduce": args.reduce,
                "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": args.nq * qds.ld * 4,
                        "d2h_bytes_per_step": args.nq * args.k * 12 + args.nq * 4, "ids_match_device_path": e2e_matches,
                        "call": "vers_sharded_ivf_search (host buffers in and out)"},
                "exchange": ("probe lists: peer-memory all-gather (remote stores + flags); top-k: ONE fused peer "
                             "gather+merge kernel; no NCCL call inside a step" if ws > 1 else None),
                "peer_exchange_us": peer_us, "single_query": single,
                "parity_spotcheck": spot,
                "gpu_launches": launches, "launch_mode": "cuda_graph_replay" if graph is not None else "eager",
                "eager_ms_per_step": eager_ms / args.steps, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "kmeans_c5": km_rec}
        print(json.dumps(line))
    comm.close()
    if ws > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def kmeans_cpu_sample(args, passes):
    """the reference's assign_to_clusters (ivfflat.rs:29-46, rayon over rows => all host cores) timed on a row sample of
    the same shape, extrapolated linearly to the full build (assign is > 99 % of the reference's build time)"""
    try:
        import oracle as vo

        use_all_host_threads(vo)
        n_s = max(1000, min(args.km_rows, args.km_cpu_rows))
        rows = vo.synth(SEED_DATA, n_s, args.km_dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS,
                        normalize=False)
        cents = vo.synth(SEED_INIT, args.km_clusters, args.km_dim, kind=1, n_centers=args.n_centers,
                         center_seed=SEED_CENTERS, normalize=False)
        t0 = time.perf_counter()
        vo.assign(rows, cents)
        dt = time.perf_counter() - t0
        est = dt * (args.km_rows / n_s) * passes
        return {"value": est, "unit": "s", "cores": vo.num_threads(), "kind": "port",
                "sample": f"one assign pass of {n_s} rows x {args.km_clusters} centroids x {args.km_dim} dims took {dt:.2f} s; "
                          f"extrapolated linearly to {args.km_rows} rows x {passes} passes (update/cost not included)"}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "s", "cores": 0, "kind": "port", "sample": f"failed: {type(e).__name__}: {e}"}


def kmeans_record(args, comm, ctx, rank, ws, local_rank, iters, warm, cpu, sample_clocks=True):
    """k-means build seconds (BASELINE.json configs[4]): `iters` Lloyd iterations (assign + ordered update + bitwise
    convergence test) + the final assign of build_kmeans (ivfflat.rs:73-100), rows sharded over the GPUs, through
    vers_sharded_kmeans_fit.  Returns the record (rank 0) / None."""
    import torch

    import vers_b200 as vb
    from vers_b200 import _abi
    from vers_b200.sharded import kmeans_fit_sharded, shard_bounds

    def barrier():
        comm.barrier()
        torch.cuda.synchronize()

    row0, n_local = shard_bounds(args.km_rows, rank, ws)
    ds = vb.Dataset.synth(ctx, SEED_DATA, n_local, args.km_dim, kind=1, n_centers=args.n_centers,
                          center_seed=SEED_CENTERS, row0=row0, normalize=False)
    init = vb.synth_init_rows(SEED_INIT, 1, args.km_clusters, args.km_rows)[0]
    km = vb.KMeans(ds, args.km_clusters)
    km.set_mode(args.km_mode)
    kmeans_fit_sharded(comm, km, init, warm, reduce=args.reduce)  # warm-up iterations (allocations, clocks)
    barrier()
    ctx.enable_timing(True)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0 and sample_clocks:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    t0 = time.perf_counter()
    ran = kmeans_fit_sharded(comm, km, init, iters, reduce=args.reduce)
    ev1.record()
    barrier()
    wall = comm.max_over_ranks(time.perf_counter() - t0)
    dev_s = comm.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    a_ms, a_n = ctx.kernel_ms(_abi.KF_ASSIGN)
    r_ms, r_n = ctx.kernel_ms(_abi.KF_LIST_SCAN)  # the exact-order redo of uncertified rows is timed in this family
    s_ms, s_n = ctx.kernel_ms(_abi.KF_SUMS)
    ctx.enable_timing(False)
    launches = ctx.launch_count - launches0
    flagged = km.last_uncertified_rows
    km.close()
    ds.close()
    if rank != 0:
        return None
    cpu_rec = kmeans_cpu_sample(args, ran + 1) if cpu else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # Roofline of the assign kernel on ALGORITHMIC flops (the GEMM form 2·N·C·D of one pass), against
    # half of the measured cuBLAS bf16 rate (tf32 runs at half the bf16 rate; MEASURED_PEAKS.json has no tf32
    # figure): the sustained number, because the kernel is timed inside a seconds-long build under the power cap.
    bf16_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    # the default kernel for dim <= 128 issues kind::f16 MMAs (fp16 operands, the bf16 rate); every other path kind::tf32
    f16_mma = args.km_mode == 0 and args.km_dim <= 128
    tf32_peak = bf16_sust if f16_mma else bf16_sust / 2
    peak_src = (("measured (MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else
                 "fallback (B200_PROFILING.md: 1.4 PFLOP/s sustained bf16") +
                (": the kernel issues kind::f16 MMAs)" if f16_mma else " / 2: the kernel issues kind::tf32 MMAs)"))
    passes = ran + 1
    flop_pass = 2.0 * args.km_rows * args.km_clusters * args.km_dim  # the GEMM form of one assign pass
    mma_per_kstep = {0: 1 if args.km_dim <= 128 else 3, 1: 0, 2: 3, 3: 1 if args.km_dim <= 128 else 3}[args.km_mode]
    avg_assign_ms = a_ms / max(a_n, 1)
    achieved = flop_pass / ws / (avg_assign_ms * 1e-3) / 1e12 if a_n else None
    kname = {3: ("tc_assign1_kernel (tcgen05 kind::tf32, 1 MMA per K step, rows resident in TMEM, top-4 + exact "
                 "rerank + certificate in the epilogue)" if args.km_dim <= 128 else
                 "tc_assign_kernel (tcgen05 kind::tf32, split hi/lo: 3 MMAs per K step)"),
             0: ("tc_assign1_kernel<F16> (tcgen05 kind::f16, 1 MMA per 16 dims, rows resident in TMEM as packed fp16, "
                 "top-4 + exact rerank + certificate in the epilogue)" if args.km_dim <= 128 else
                 "tc_assign_kernel (tcgen05 kind::tf32, split hi/lo: 3 MMAs per K step)"),
             1: "assign_kernel (exact order, fp32 pipe)",
             2: "tc_assign_kernel (tcgen05 kind::tf32, split hi/lo: 3 MMAs per K step)"}[args.km_mode]
    return {"metric": "k-means build seconds (50Mx128, 16384 centroids, 20 iterations)", "value": dev_s, "unit": "s",
            "n_gpus": ws, "steps": ran, "warmup": warm, "ms_per_step": dev_s / max(ran, 1) * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"IVFFlatIndex::build_kmeans: {args.km_rows}x{args.km_dim} fp32 synthetic clustered "
                                   f"(unnormalized), {args.km_clusters} centroids, {ran} Lloyd iterations + final "
                                   f"assign (BASELINE.json configs[4])",
                       "rows": args.km_rows, "dim": args.km_dim, "clusters": args.km_clusters, "iterations": ran,
                       "reduce": args.reduce, "sharding": f"rows/{ws} per GPU",
                       "l2": "every assign pass streams the row shard (>= 3 GB) once: far larger than L2"},
            "wall_s": wall, "assign_passes": passes, "uncertified_rows_last_pass": flagged,
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": tf32_peak,
                         "unit": "TFLOP/s", "frac": (achieved / tf32_peak) if achieved else None, "traffic": None,
                         "peak_source": peak_src, "algorithmic_flop_per_launch": flop_pass / ws,
                         "mma_per_k_step": mma_per_kstep, "avg_launch_ms": avg_assign_ms,
                         "kernel_share_of_step": a_ms * 1e-3 / dev_s if dev_s else None,
                         "sums_ms_per_iteration": s_ms / max(s_n, 1),
                         "exact_redo_ms_per_pass": (r_ms / r_n) if r_n else 0.0},
            "cpu_baseline": cpu_rec}


def main_kmeans(args):
    import torch

    import vers_b200 as vb
    from vers_b200.sharded import Comm

    rank, ws, local_rank = setup_ranks()
    ctx = vb.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    comm = Comm(ctx)
    iters = args.steps if args.steps != 10 else 20
    rec = kmeans_record(args, comm, ctx, rank, ws, local_rank, iters, min(args.warmup, 3),
                        cpu=(ws == 1 and not args.no_cpu_baseline))
    if rank == 0:
        print(json.dumps(rec))
    comm.close()
    if ws > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def main_flat(args):
    """BASELINE.json configs[1]: exhaustive top-10 over 1M x 300 normalized vectors, 1000-query batch, 1 GPU
    (utils::search_exhaustive semantics, squared L2)."""
    import torch

    import vers_b200 as vb
    from vers_b200 import _abi
    from vers_b200.sharded import device_view
    import ctypes as C

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = vb.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    n, dim = args.flat_rows, args.flat_dim
    ds = vb.Dataset.synth(ctx, SEED_DATA, n, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS, row0=0,
                          normalize=True)
    ds.set_flat_mode(args.flat_mode)
    qds = vb.Dataset.synth(ctx, SEED_QUERY, args.nq, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS,
                           row0=0, normalize=True)
    d_q = device_view(qds.device_ptr, (args.nq, qds.ld))
    ids = torch.empty((args.nq, args.k), dtype=torch.int64, device=dev)
    dd = torch.empty((args.nq, args.k), dtype=torch.float32, device=dev)
    cc = torch.empty((args.nq,), dtype=torch.int32, device=dev)

    def step():
        _abi.check(vb.lib().vers_flat_search_dev(ds.h, C.c_void_p(d_q.data_ptr()), args.nq, args.k, 0,
                                                 C.c_void_p(ids.data_ptr()), C.c_void_p(dd.data_ptr()),
                                                 C.c_void_p(cc.data_ptr())))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    sampler = ClockSampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    # the scan kernel's own duration: the same K steps again with CUDA events around every launch of the family
    ctx.enable_timing(True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    f_ms, f_n = ctx.kernel_ms(_abi.KF_FLAT_SCAN)
    ctx.enable_timing(False)
    st = ds.last_flat_search_stats()
    # e2e through the host-buffer C-ABI call
    h_q = torch.empty((args.nq, qds.ld), dtype=torch.float32).pin_memory()
    h_q.copy_(d_q)
    h_ids = torch.empty((args.nq, args.k), dtype=torch.int64).pin_memory()
    h_d = torch.empty((args.nq, args.k), dtype=torch.float32).pin_memory()
    h_c = torch.empty((args.nq,), dtype=torch.int32).pin_memory()

    def e2e():
        _abi.check(vb.lib().vers_flat_search(ds.h, h_q.data_ptr(), args.nq, qds.ld, args.k, 0, h_ids.data_ptr(),
                                             h_d.data_ptr(), h_c.data_ptr()))

    for _ in range(args.warmup):
        e2e()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e()
    e2e_s = time.perf_counter() - t0
    match = bool(torch.equal(ids.cpu(), h_ids))
    peak, peak_src = measured_peaks()
    alg = n * dim * 4
    avg_ms = f_ms / max(f_n, 1)
    cpu = None
    if not args.no_cpu_baseline:
        import oracle as vo

        use_all_host_threads(vo)
        rows_s = vo.synth(SEED_DATA, n // 8, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS)
        qs = vo.synth(SEED_QUERY, 64, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS)
        t0 = time.perf_counter()
        vo.exhaustive(rows_s, qs, args.k, 0)
        dt = time.perf_counter() - t0
        cpu = {"value": 64 / dt / 8, "unit": "queries/s", "cores": vo.num_threads(), "kind": "port",
               "sample": f"64 queries over rows/8 ({n // 8}x{dim}), QPS divided by 8 (scan work is linear in rows); "
                         f"OpenMP over queries"}
    flat_traffic, flat_traffic_src = committed_traffic_key(f"flat_{n}x{dim}_nq{args.nq}_k{args.k}_mode{args.flat_mode}")
    line = {"metric": f"exhaustive top-{args.k} QPS ({n}x{dim}, batch {args.nq})", "value": args.nq * args.steps / (dev_ms * 1e-3),
            "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"search_exhaustive batch: {n}x{dim} fp32 synthetic clustered+normalized, top_k {args.k}, "
                                   f"{args.nq}-query batch (BASELINE.json configs[1])", "rows": n, "dim": dim,
                       "top_k": args.k, "batch": args.nq, "flat_mode": args.flat_mode,
                       "l2": "the dataset (1.2 GB) is far larger than L2"},
            "e2e": {"value": args.nq * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": args.nq * qds.ld * 4,
                    "d2h_bytes_per_step": args.nq * args.k * 12 + args.nq * 4, "ids_match_device_path": match},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": ("flat_stream_kernel (bulk-copy staged row tiles, lane = row, exact "
                                                    "order)" if args.nq <= 8 and args.flat_mode == 0 else
                                                    "flat search step (tcgen05 candidate scan + merge + exact rerank)"
                                                    if args.flat_mode == 0 else "flat_scan_kernel (exact order, fp32 pipe)"),
                         "achieved": alg / (avg_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (avg_ms * 1e-3) / 1e9 / peak, "traffic": flat_traffic,
                         "traffic_source": flat_traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "avg_launch_ms": avg_ms,
                         "note": ("algorithmic bytes = the dataset streamed once per batch" if args.nq <= 8 else
                                  "algorithmic bytes = the dataset streamed once per batch; the candidate scan re-streams "
                                  "it once per 32-query group below 96 queries, once per 128-query block above (through "
                                  "L2: DRAM traffic stays ~1.1x the table), so frac is far below 1 by construction"),
                         "uncertified_queries_last_step": st["uncertified_queries"],
                         "max_candidate_error_last_step": st["max_candidate_error"]},
            "cpu_baseline": cpu}
    print(json.dumps(line))


def main_hnswdist(args):
    """HNSW distance offload.  One step = --pairs (query, neighbour id) pairs evaluated with
    Vector::cosine_similarity_simd's association (base.rs:158-223) over the 1M x 300 table of BASELINE.json configs[1];
    neighbour ids are uniform random (a graph frontier gathers rows from all over the table), 1000 queries."""
    import ctypes as C

    import torch

    import vers_b200 as vb
    from vers_b200 import _abi
    from vers_b200.sharded import device_view

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = vb.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    n, dim, npairs = args.flat_rows, args.flat_dim, args.pairs
    ds = vb.Dataset.synth(ctx, SEED_DATA, n, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS, row0=0,
                          normalize=True)
    qds = vb.Dataset.synth(ctx, SEED_QUERY, args.nq, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS,
                           row0=0, normalize=True)
    d_q = device_view(qds.device_ptr, (args.nq, qds.ld))
    rng = np.random.default_rng(5)
    pr = rng.integers(0, n, npairs).astype(np.int64)
    pq = (np.arange(npairs) % args.nq).astype(np.int32)
    h_pr, h_pq = torch.from_numpy(pr).pin_memory(), torch.from_numpy(pq).pin_memory()
    d_pr, d_pq = h_pr.to(dev), h_pq.to(dev)
    d_out = torch.empty((npairs,), dtype=torch.float32, device=dev)
    d_bad = torch.zeros((1,), dtype=torch.int32, device=dev)

    def step():
        _abi.check(vb.lib().vers_pair_distances_simd_dev(ds.h, C.c_void_p(d_q.data_ptr()), args.nq, qds.ld,
                                                         C.c_void_p(d_pq.data_ptr()), C.c_void_p(d_pr.data_ptr()), npairs,
                                                         1, C.c_void_p(d_out.data_ptr()), C.c_void_p(d_bad.data_ptr())))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    sampler = ClockSampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    assert int(d_bad.item()) == 0
    # e2e through the host-buffer call: queries, pair lists in; distances out
    h_q = torch.empty((args.nq, qds.ld), dtype=torch.float32).pin_memory()
    h_q.copy_(d_q)
    h_out = torch.empty((npairs,), dtype=torch.float32).pin_memory()

    def e2e():
        _abi.check(vb.lib().vers_pair_distances_simd(ds.h, h_q.data_ptr(), args.nq, qds.ld, h_pq.data_ptr(),
                                                     h_pr.data_ptr(), npairs, 1, h_out.data_ptr()))

    for _ in range(args.warmup):
        e2e()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e()
    e2e_s = (time.perf_counter() - t0) / args.steps
    match = bool(torch.equal(d_out.cpu().view(torch.int32), h_out.view(torch.int32)))
    # parity spot check + CPU leg: the oracle on a bounded sample of the same pairs
    import oracle as vo

    use_all_host_threads(vo)
    rows = ds.download()
    q = qds.download()
    ns = min(npairs, 1 << 20)
    vo.pair_distances_simd(rows, q, pr[:4096].astype(np.uint64), pq[:4096].astype(np.uint32), 1)
    t0 = time.perf_counter()
    want = vo.pair_distances_simd(rows, q, pr[:ns].astype(np.uint64), pq[:ns].astype(np.uint32), 1)
    cdt = time.perf_counter() - t0
    bits_equal = bool(np.array_equal(want.view(np.uint32), h_out.numpy()[:ns].view(np.uint32)))
    peak, peak_src = measured_peaks()
    alg = npairs * dim * 4
    line = {"metric": "HNSW distance offload: cosine_similarity_simd pairs/s (1Mx300 table, random neighbour ids)",
            "value": npairs / (dev_ms * 1e-3), "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"Vector::cosine_similarity_simd over {npairs} (query, neighbour) pairs per step, "
                                   f"{n}x{dim} table, {args.nq} queries (SURVEY.md §8f HNSW distance offload)",
                       "rows": n, "dim": dim, "pairs": npairs,
                       "l2": f"the pairs gather {alg / 1e9:.1f} GB of rows per step from a 1.2 GB table: every step re-reads "
                             f"the table ~{alg / (n * dim * 4):.1f}x in random order, far beyond the 126 MB L2"},
            "e2e": {"value": npairs / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": npairs * 12 + args.nq * qds.ld * 4,
                    "d2h_bytes_per_step": npairs * 4, "bits_match_device_path": match,
                    "call": "vers_pair_distances_simd (host buffers in and out)"},
            "parity_spotcheck": {"pairs": ns, "distance_bits_equal_oracle": bits_equal},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "pair_distances_simd_kernel (one warp per pair, lane = SIMD chunk)",
                         "achieved": alg / (dev_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (dev_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "avg_launch_ms": dev_ms,
                         "note": "row gathers of dim*4 = 1200 B at random positions; the query rows stay in L2"},
            "cpu_baseline": {"value": ns / cdt, "unit": "pairs/s", "cores": vo.num_threads(), "kind": "port",
                             "sample": f"the first {ns} pairs of the step, OpenMP over pairs (the reference evaluates them "
                                       f"one at a time inside the traversal, hnsw.rs:258-273)"}}
    print(json.dumps(line))


def main_lsh(args):
    """BASELINE.json configs[2]: the hyperplane forest ("LSH", indexes/lsh.rs) on 1M x 300, 16 trees, max_size 100,
    1000-query batches, 1 GPU.  build = level-synchronous split of every tree on the device; search = batched
    traversal + leaf scan + exact rerank.  The search entry point takes host buffers, so value == e2e here."""
    import vers_b200 as vb

    n, dim = args.flat_rows, args.flat_dim
    ctx = vb.Context(0)
    ds = vb.Dataset.synth(ctx, SEED_DATA, n, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS, row0=0,
                          normalize=True)
    rows = ds.download()
    qds = vb.Dataset.synth(ctx, SEED_QUERY, args.nq, dim, kind=1, n_centers=args.n_centers, center_seed=SEED_CENTERS,
                           row0=0, normalize=True)
    q = qds.download()
    t0 = time.perf_counter()
    idx = vb.ANNIndex.build_index(16, 100, rows, None, seed=4, ctx=ctx)
    ctx.sync()
    build_s = time.perf_counter() - t0
    for _ in range(args.warmup):
        idx.search_batch(q, args.k)
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = ctx.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids, d, cnt = idx.search_batch(q, args.k)
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    nrec = min(args.recall_queries, args.nq)
    gi, _, _ = vb.search_exhaustive_batch(ds, q[:nrec], args.k, 0)
    recall = float(np.mean([len(set(ids[i][: cnt[i]]) & set(gi[i])) / args.k for i in range(nrec)]))
    qps = args.nq * args.steps / dt
    info = idx.info()
    trees = 16
    # algorithmic bytes per query: trees x (depth x 4(D+1) plane bytes + leaf rows x 4D) + candidates x 4D
    # for the rerank.  depth and leaf size are the forest's own averages (a tree of L leaves over n rows).
    leaves = (info["num_nodes"] / trees + 1) / 2
    depth = float(np.log2(max(leaves, 1.0)))
    leaf_rows = n / max(leaves, 1.0)
    cand = min(trees * args.k, n)
    bytes_q = trees * (depth * 4 * (dim + 1) + leaf_rows * 4 * dim) + cand * 4 * dim
    peak, peak_src = measured_peaks()
    achieved = bytes_q * args.nq / (dt / args.steps) / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        import oracle as vo

        use_all_host_threads(vo)
        n_s = n // 8
        t0 = time.perf_counter()
        o = vo.LSH(rows[:n_s], None, trees, 100, 4)
        cpu_build = time.perf_counter() - t0
        o.search(q[:64], args.k)
        t0 = time.perf_counter()
        oi, od, oc = o.search(q, args.k)
        cdt = time.perf_counter() - t0
        cpu = {"value": args.nq / cdt, "unit": "queries/s", "cores": vo.num_threads(), "kind": "port",
               "cpu_build_s": round(cpu_build, 2),
               "sample": f"{args.nq} queries on a forest over rows/8 ({n_s}x{dim}, {trees} trees, max_size 100: 3 levels "
                         f"shallower than the full forest, same leaf size => ~15 % less work per query than the full "
                         f"config); OpenMP over queries (the reference parallelises over trees, lsh.rs:268)"}
    line = {"metric": "hyperplane-forest (LSH) search QPS (1Mx300, 16 trees, batch 1k)", "value": qps, "unit": "queries/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"ANNIndex build_index + search_approximate batch: {n}x{dim}, 16 trees, max_size 100, "
                                   f"top_k {args.k}, {args.nq}-query batch (BASELINE.json configs[2])",
                       "rows": n, "dim": dim, "trees": trees, "max_size": 100, "nodes": info["num_nodes"],
                       "l2": "per step the batch touches ~1.7 GB of planes, leaf rows and candidates scattered over the "
                             "1.2 GB table + 0.6 GB of planes: larger than L2"},
            "build_s": build_s, "recall_at_10": recall,
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": args.nq * dim * 4,
                    "d2h_bytes_per_step": args.nq * args.k * 12 + args.nq * 4,
                    "call": "vers_lsh_search (host buffers in and out): value == e2e for this workload"},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "lsh search step (batched traversal + leaf scan + exact rerank)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_q * args.nq,
                         "avg_tree_depth": depth, "avg_leaf_rows": leaf_rows,
                         "note": "whole step incl. H2D/D2H and the host synchronise; the accesses are data-dependent "
                                 "gathers (one plane per level per tree, one leaf per tree), so the bound is latency "
                                 "and sector efficiency rather than streaming bandwidth"},
            "cpu_baseline": cpu}
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    elif a.workload == "kmeans":
        main_kmeans(a)
    elif a.workload == "flat":
        main_flat(a)
    elif a.workload == "hnswdist":
        main_hnswdist(a)
    elif a.workload == "lsh":
        main_lsh(a)
    else:
        main_ours(a)
