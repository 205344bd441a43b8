// flat.cu — exhaustive search (utils::search_exhaustive, utils.rs:68-82): exact-order scan of every row,
// per-CTA private top-k, then a merge by (distance, id).  Also hosts the generic merge / scan helpers.
#include "scan.cuh"

namespace vers {

template <class Cfg, int OP, int XF>
__global__ void __launch_bounds__(Cfg::NT, 2) flat_scan_kernel(FlatScanParams p) {
    extern __shared__ __align__(16) float smem[];
    float* list_d = smem + Cfg::TILE_FLOATS;
    uint32_t* list_p = reinterpret_cast<uint32_t*>(list_d + Cfg::NLISTS * p.kpad);
    const uint64_t chunk = blockIdx.x;
    const uint64_t b0 = (uint64_t)blockIdx.y * Cfg::TB;
    RowSrc B = p.B;
    if (p.nB_dev) {
        B.n = *p.nB_dev;
        if (b0 >= B.n) return;
    }
    const uint64_t r0 = chunk * p.rows_per_chunk;
    const uint64_t r1 = min(p.A.n, r0 + p.rows_per_chunk);
    lists_init<Cfg>(list_d, list_p, p.kpad);
    for (uint64_t a0 = r0; a0 < r1; a0 += Cfg::TA) {
        float acc[Cfg::MA][Cfg::MB];
        tile_compute<Cfg, OP>(acc, p.A, a0, B, b0, p.ld, smem);
        tile_select_topk<Cfg, XF>(acc, a0, r1, b0, B.n, p.k, p.kpad, list_d, list_p, 0);
    }
    __syncwarp();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int SLOTS_PER_WARP = Cfg::TBS_PER_WARP * Cfg::MB;
    for (int s = 0; s < SLOTS_PER_WARP; ++s) {
        int slot = warp * SLOTS_PER_WARP + s, col, split;
        slot_to_col<Cfg>(slot, col, split);
        uint64_t q = b0 + (uint64_t)col;
        if (q >= B.n) continue;
        uint64_t base = (q * p.nparts + chunk * Cfg::NSPLIT + split) * p.k;
        for (uint32_t e = lane; e < p.k; e += 32) {
            p.part_d[base + e] = list_d[slot * p.kpad + e];
            p.part_p[base + e] = list_p[slot * p.kpad + e];
        }
    }
}

constexpr int MERGE_WARPS = 4;

__global__ void __launch_bounds__(MERGE_WARPS * 32) merge_topk_kernel(MergeParams p) {
    extern __shared__ __align__(16) unsigned char msm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * MERGE_WARPS + warp;
    if (q >= p.nq) return;
    if (p.nq_dev && q >= *p.nq_dev) return;
    if (p.skip_if_zero && *p.skip_if_zero == 0) return;
    if (p.qmask && p.qmask[q] == 0) return;
    uint64_t* sp = reinterpret_cast<uint64_t*>(msm) + (size_t)warp * p.k;
    float* sd = reinterpret_cast<float*>(msm + (size_t)MERGE_WARPS * p.k * 8) + (size_t)warp * p.k;
    for (uint32_t e = lane; e < p.k; e += 32) {
        sd[e] = __int_as_float(0x7f800000);
        sp[e] = 0xffffffffffffffffull;
    }
    __syncwarp();
    uint64_t beg, end;
    if (p.seg) {
        beg = p.seg[(uint64_t)q * p.seg_stride] * p.seg_scale;
        end = p.seg[(uint64_t)(q + 1) * p.seg_stride] * p.seg_scale;
    } else {
        beg = (uint64_t)q * p.per_query;
        end = beg + p.per_query;
    }
    const int k = (int)p.k;
    for (uint64_t e0 = beg; e0 < end; e0 += 32) {
        uint64_t e = e0 + lane;
        float v = 0.f;
        uint64_t id = 0;
        bool live = false;
        if (e < end) {
            uint32_t pp = p.part_p[e];
            if (pp != 0xffffffffu) {
                live = true;
                v = p.part_d[e];
                id = p.map ? p.map[pp] : p.id_base + pp;
            }
        }
        while (true) {
            bool pass = live && entry_less<uint64_t>(v, id, sd[k - 1], sp[k - 1]);
            unsigned m = __ballot_sync(FULL_MASK, pass);
            if (!m) break;
            int src = __ffs(m) - 1;
            float bv = __shfl_sync(FULL_MASK, v, src);
            uint64_t bid = __shfl_sync(FULL_MASK, id, src);
            warp_topk_insert<uint64_t>(sd, sp, k, bv, bid, lane);
            if (lane == src) live = false;
        }
    }
    uint32_t cnt = 0;
    for (uint32_t e0 = 0; e0 < p.k; e0 += 32) {
        uint32_t e = e0 + lane;
        bool have = false;
        if (e < p.k) {
            p.out_ids[(uint64_t)q * p.k + e] = sp[e];
            p.out_d[(uint64_t)q * p.k + e] = sd[e];
            have = sp[e] != 0xffffffffffffffffull;
        }
        cnt += __popc(__ballot_sync(FULL_MASK, have));
    }
    if (p.out_cnt && lane == 0) p.out_cnt[q] = cnt;
}

int32_t launch_merge(vers_ctx* ctx, const MergeParams& mp) {
    if (mp.nq == 0) return VERS_OK;
    size_t smem = (size_t)MERGE_WARPS * mp.k * 12;
    merge_topk_kernel<<<(unsigned)ceil_div(mp.nq, MERGE_WARPS), MERGE_WARPS * 32, smem, ctx->stream>>>(mp);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

// single block, 1024 threads x IT consecutive elements per tile: thread-serial scan -> warp scan -> block scan
template <int IT>
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const uint32_t* in, uint64_t n, uint64_t* out,
                                                              const uint32_t* skip_if_zero) {
    if (skip_if_zero && *skip_if_zero == 0) return;
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024 * IT) {
        const uint64_t i0 = base + (uint64_t)threadIdx.x * IT;
        uint32_t v[IT];
        if (i0 + IT <= n && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
#pragma unroll
            for (int k = 0; k < IT; k += 4) {
                const uint4 a = *reinterpret_cast<const uint4*>(in + i0 + k);
                v[k] = a.x; v[k + 1] = a.y; v[k + 2] = a.z; v[k + 3] = a.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < IT; ++k) v[k] = i0 + k < n ? in[i0 + k] : 0u;
        }
        uint64_t tsum = 0;
#pragma unroll
        for (int k = 0; k < IT; ++k) tsum += v[k];
        uint64_t x = tsum;
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(FULL_MASK, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t y = __shfl_up_sync(FULL_MASK, w, o);
                if (lane >= o) w += y;
            }
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        uint64_t run = carry + (warp ? warp_tot[warp - 1] : 0) + (x - tsum);
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            if (i0 + k < n) out[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

// several blocks, one 8192-element tile each, carries chained through flags[] (flags[b] = inclusive total of the
// tiles 0..b, plus 1; 0 = not published yet).  The grid never exceeds the SM count, so every block is resident and
// the spin on the predecessor terminates.
__global__ void __launch_bounds__(1024) exclusive_scan_chained_kernel(const uint32_t* in, uint64_t n, uint64_t* out,
                                                                      unsigned long long* flags,
                                                                      const uint32_t* skip_if_zero) {
    constexpr int IT = 8;
    if (skip_if_zero && *skip_if_zero == 0) return;
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t i0 = ((uint64_t)blockIdx.x * 1024 + threadIdx.x) * IT;
    uint32_t v[IT];
    if (i0 + IT <= n && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < IT; k += 4) {
            const uint4 a = *reinterpret_cast<const uint4*>(in + i0 + k);
            v[k] = a.x; v[k + 1] = a.y; v[k + 2] = a.z; v[k + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < IT; ++k) v[k] = i0 + k < n ? in[i0 + k] : 0u;
    }
    uint64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < IT; ++k) tsum += v[k];
    uint64_t x = tsum;
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(FULL_MASK, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = warp_tot[lane];
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(FULL_MASK, w, o);
            if (lane >= o) w += y;
        }
        warp_tot[lane] = w;  // inclusive over warps
        if (lane == 31) {
            uint64_t carry = 0;
            if (blockIdx.x > 0) {
                volatile unsigned long long* f = flags + blockIdx.x - 1;
                unsigned long long got;
                while ((got = *f) == 0ull) {
                }
                carry = got - 1;
                *f = 0ull;  // consumed exactly once: the flags are all zero again when the kernel ends
            }
            carry_s = carry;
            if (blockIdx.x == gridDim.x - 1) {
                out[n] = carry + w;
            } else {
                __threadfence();
                *(volatile unsigned long long*)(flags + blockIdx.x) = carry + w + 1;
            }
        }
    }
    __syncthreads();
    uint64_t run = carry_s + (warp ? warp_tot[warp - 1] : 0) + (x - tsum);
#pragma unroll
    for (int k = 0; k < IT; ++k) {
        if (i0 + k < n) out[i0 + k] = run;
        run += v[k];
    }
}

int32_t launch_exclusive_scan(vers_ctx* ctx, const uint32_t* d_in, uint64_t n, uint64_t* d_out,
                              const uint32_t* skip_if_zero) {
    const uint64_t tiles = ceil_div(n, 8192);
    if (tiles > 1 && tiles <= (uint64_t)std::min(ctx->sm_count, 256)) {
        if (!ctx->scan_flags) {
            VERS_CUDA(cudaMalloc(&ctx->scan_flags, 256 * 8));
            VERS_CUDA(cudaMemsetAsync(ctx->scan_flags, 0, 256 * 8, ctx->stream));  // the kernel leaves them zero
        }
        exclusive_scan_chained_kernel<<<(unsigned)tiles, 1024, 0, ctx->stream>>>(d_in, n, d_out, ctx->scan_flags,
                                                                                skip_if_zero);
        VERS_LAUNCH_CHECK(ctx);
        return VERS_OK;
    }
    // one tile when it fits: 1024 threads x 8 or 32 consecutive elements
    if (n <= 8192)
        exclusive_scan_kernel<8><<<1, 1024, 0, ctx->stream>>>(d_in, n, d_out, skip_if_zero);
    else
        exclusive_scan_kernel<32><<<1, 1024, 0, ctx->stream>>>(d_in, n, d_out, skip_if_zero);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

template <class Cfg, int OP, int XF>
static int32_t launch_flat_scan(vers_ctx* ctx, FlatScanParams& p, uint32_t nchunks, uint32_t nq) {
    auto kern = flat_scan_kernel<Cfg, OP, XF>;
    size_t smem = scan_smem_bytes(Cfg::TILE_FLOATS, Cfg::NLISTS, p.kpad);
    VERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nchunks, (unsigned)ceil_div(nq, Cfg::TB));
    kern<<<grid, Cfg::NT, smem, ctx->stream>>>(p);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

ScanPlan scan_topk_plan(const vers_ctx* ctx, uint64_t nA, uint32_t nq, uint32_t k) {
    ScanPlan pl;
    pl.narrow = nq <= 24;
    const uint32_t TA = pl.narrow ? NarrowCfg::TA : WideCfg::TA;
    const uint32_t TB = pl.narrow ? NarrowCfg::TB : WideCfg::TB;
    const uint32_t NSPLIT = pl.narrow ? NarrowCfg::NSPLIT : WideCfg::NSPLIT;
    const uint64_t nqtiles = ceil_div(nq ? nq : 1, TB);
    uint64_t target = (uint64_t)ctx->sm_count * 4;
    uint64_t nchunks = ceil_div(target, nqtiles);
    uint64_t max_chunks = nA ? ceil_div(nA, TA) : 1;
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks > 2048) nchunks = 2048;
    if (nchunks < 1) nchunks = 1;
    uint64_t rpc = ceil_div(nA ? nA : 1, nchunks);
    rpc = ceil_div(rpc, TA) * TA;
    nchunks = nA ? ceil_div(nA, rpc) : 1;
    pl.nchunks = nchunks;
    pl.rows_per_chunk = rpc;
    pl.nparts = (uint32_t)(nchunks * NSPLIT);
    pl.entries = (size_t)nq * pl.nparts * k;
    ScratchCarver c(nullptr);
    c.plan<float>(pl.entries);
    c.plan<uint32_t>(pl.entries);
    pl.bytes = c.off + 256;
    return pl;
}

int32_t scan_topk_run(vers_ctx* ctx, const ScanPlan& pl, void* scratch, const RowSrc& A, const RowSrc& B, uint32_t nq,
                      uint32_t ld, uint32_t k, uint32_t metric, const uint64_t* id_map, uint64_t id_base,
                      uint64_t* d_ids, float* d_d, uint32_t* d_cnt, int family, const uint32_t* nq_dev) {
    if (k == 0 || nq == 0) return VERS_OK;
    if (k > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "top_k %u > %u", k, VERS_MAX_TOPK);
    if (A.n >= 0xffffffffull) return fail(VERS_ERR_UNSUPPORTED, "more than 2^32-2 rows per GPU shard");
    FlatScanParams p;
    p.A = A;
    p.B = B;
    p.ld = ld;
    p.k = k;
    p.kpad = round_up(k, 32);
    p.rows_per_chunk = pl.rows_per_chunk;
    p.nparts = pl.nparts;
    p.nB_dev = nq_dev;
    ScratchCarver sc(scratch);
    p.part_d = sc.take<float>(pl.entries);
    p.part_p = sc.take<uint32_t>(pl.entries);
    {
        FamilyTimer ft(ctx, family);
        if (pl.narrow) {
            if (metric == VERS_METRIC_L2SQ)
                VERS_TRY((launch_flat_scan<NarrowCfg, OP_L2SQ, 0>(ctx, p, (uint32_t)pl.nchunks, nq)));
            else
                VERS_TRY((launch_flat_scan<NarrowCfg, OP_DOT, 1>(ctx, p, (uint32_t)pl.nchunks, nq)));
        } else {
            if (metric == VERS_METRIC_L2SQ)
                VERS_TRY((launch_flat_scan<WideCfg, OP_L2SQ, 0>(ctx, p, (uint32_t)pl.nchunks, nq)));
            else
                VERS_TRY((launch_flat_scan<WideCfg, OP_DOT, 1>(ctx, p, (uint32_t)pl.nchunks, nq)));
        }
    }
    MergeParams mp;
    mp.part_d = p.part_d;
    mp.part_p = p.part_p;
    mp.seg = nullptr;
    mp.seg_scale = 1;
    mp.seg_stride = 1;
    mp.per_query = (uint64_t)p.nparts * k;
    mp.map = id_map;
    mp.id_base = id_base;
    mp.nq = nq;
    mp.k = k;
    mp.out_ids = d_ids;
    mp.out_d = d_d;
    mp.out_cnt = d_cnt;
    mp.qmask = nullptr;
    mp.nq_dev = nq_dev;
    return launch_merge(ctx, mp);
}

int32_t scan_topk_dev(vers_ctx* ctx, const RowSrc& A, const RowSrc& B, uint32_t nq, uint32_t ld, uint32_t k,
                      uint32_t metric, const uint64_t* id_map, uint64_t id_base, uint64_t* d_ids, float* d_d,
                      uint32_t* d_cnt, int family) {
    if (k == 0 || nq == 0) return VERS_OK;
    ScanPlan pl = scan_topk_plan(ctx, A.n, nq, k);
    VERS_TRY(scratch_reserve(ctx, pl.bytes));
    return scan_topk_run(ctx, pl, ctx->scratch, A, B, nq, ld, k, metric, id_map, id_base, d_ids, d_d, d_cnt, family);
}

// merge of already-global (id, dist) lists: [parts][nq][k] -> [nq][k]; the step after the all-gather of per-GPU top-k
__global__ void __launch_bounds__(MERGE_WARPS * 32)
    merge_ids_kernel(const uint64_t* __restrict__ ids_all, const float* __restrict__ d_all, uint32_t parts,
                     uint64_t stride_ids, uint64_t stride_d, uint32_t nq, uint32_t k, uint64_t* out_ids, float* out_d,
                     uint32_t* out_cnt) {
    extern __shared__ __align__(16) unsigned char msm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * MERGE_WARPS + warp;
    if (q >= nq) return;
    uint64_t* sp = reinterpret_cast<uint64_t*>(msm) + (size_t)warp * k;
    float* sd = reinterpret_cast<float*>(msm + (size_t)MERGE_WARPS * k * 8) + (size_t)warp * k;
    for (uint32_t e = lane; e < k; e += 32) {
        sd[e] = __int_as_float(0x7f800000);
        sp[e] = 0xffffffffffffffffull;
    }
    __syncwarp();
    const uint32_t total = parts * k;
    for (uint32_t e0 = 0; e0 < total; e0 += 32) {
        uint32_t e = e0 + lane;
        float v = 0.f;
        uint64_t id = 0xffffffffffffffffull;
        if (e < total) {
            uint64_t part = e / k, at = (uint64_t)q * k + (e % k);
            id = ids_all[part * stride_ids + at];
            v = d_all[part * stride_d + at];
        }
        bool live = id != 0xffffffffffffffffull;
        while (true) {
            bool pass = live && entry_less<uint64_t>(v, id, sd[k - 1], sp[k - 1]);
            unsigned m = __ballot_sync(FULL_MASK, pass);
            if (!m) break;
            int src = __ffs(m) - 1;
            float bv = __shfl_sync(FULL_MASK, v, src);
            uint64_t bid = __shfl_sync(FULL_MASK, id, src);
            warp_topk_insert<uint64_t>(sd, sp, (int)k, bv, bid, lane);
            if (lane == src) live = false;
        }
    }
    uint32_t cnt = 0;
    for (uint32_t e0 = 0; e0 < k; e0 += 32) {
        uint32_t e = e0 + lane;
        bool have = false;
        if (e < k) {
            out_ids[(uint64_t)q * k + e] = sp[e];
            out_d[(uint64_t)q * k + e] = sd[e];
            have = sp[e] != 0xffffffffffffffffull;
        }
        cnt += __popc(__ballot_sync(FULL_MASK, have));
    }
    if (out_cnt && lane == 0) out_cnt[q] = cnt;
}

}  // namespace vers

using namespace vers;

// caller holds ctx->mu
static int32_t flat_search_dev_locked(vers_dataset* ds, const float* d_queries, uint32_t nq, uint32_t top_k,
                                      uint32_t metric, uint64_t* d_ids, float* d_dists, uint32_t* d_counts) {
    vers_ctx* ctx = ds->ctx;
    if (metric == VERS_METRIC_L2SQ && ds->flat_mode == 0 && nq >= 32 && top_k >= 1 && top_k <= 64) {
        // large batches: tensor-core candidate keys -> exact-order rerank -> certificate -> exact redo (ivf.cu)
        if (!ds->d_norm) {
            VERS_CUDA(cudaMalloc(&ds->d_norm, (ds->n ? ds->n : 1) * 4));
            VERS_CUDA(cudaMalloc(&ds->d_nmax, 4));
            VERS_CUDA(cudaMalloc(&ds->d_stats, 128));
            VERS_TRY(launch_rownorm(ctx, ds->d_rows, ds->ld, ds->n, ds->d_norm, ds->d_nmax));
        }
        bool used = false;
        int32_t rc = flat_search_tc_plan_and_run(ctx, ds->d_rows, ds->n, ds->ld, ds->d_norm, ds->d_nmax, ds->id_base,
                                                 ds->d_stats, d_queries, nq, top_k, d_ids, d_dists, d_counts, &used);
        if (used || rc != VERS_ERR_UNSUPPORTED) return rc;
    }
    if (ds->d_stats) VERS_CUDA(cudaMemsetAsync(ds->d_stats, 0, 64, ctx->stream));  // the exact-order engine ran
    RowSrc A{ds->d_rows, nullptr, ds->ld, ds->n};
    RowSrc B{d_queries, nullptr, ds->ld, nq};
    return scan_topk_dev(ctx, A, B, nq, ds->ld, top_k, metric, nullptr, ds->id_base, d_ids, d_dists, d_counts,
                         KF_FLAT_SCAN);
}

extern "C" int32_t vers_flat_search_dev(vers_dataset* ds, const float* d_queries, uint32_t nq, uint32_t top_k,
                                        uint32_t metric, uint64_t* d_ids, float* d_dists, uint32_t* d_counts) {
    if (!ds || (!d_queries && nq) || !d_ids || !d_dists) return fail(VERS_ERR_ARG, "flat_search_dev: null argument");
    if (metric > VERS_METRIC_COSINE) return fail(VERS_ERR_ARG, "flat_search_dev: unknown metric %u", metric);
    std::lock_guard<std::mutex> lk(ds->ctx->mu);
    VERS_CUDA(cudaSetDevice(ds->ctx->device));
    return flat_search_dev_locked(ds, d_queries, nq, top_k, metric, d_ids, d_dists, d_counts);
}

extern "C" int32_t vers_flat_set_mode(vers_dataset* ds, int32_t mode) {
    if (!ds) return fail(VERS_ERR_ARG, "flat_set_mode: null");
    if (mode < 0 || mode > 1) return fail(VERS_ERR_ARG, "flat_set_mode: mode %d", mode);
    ds->flat_mode = mode;
    return VERS_OK;
}

extern "C" int32_t vers_flat_last_search_stats(const vers_dataset* ds, uint64_t out[8]) {
    if (!ds || !out) return fail(VERS_ERR_ARG, "flat_last_search_stats: null");
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if (!ds->d_stats) return VERS_OK;
    VERS_CUDA(cudaSetDevice(ds->ctx->device));
    VERS_CUDA(cudaMemcpyAsync(out, ds->d_stats, 64, cudaMemcpyDeviceToHost, ds->ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ds->ctx->stream));
    return VERS_OK;
}

namespace vers {
// host buffers -> device staging shared by flat / ivf / lsh host entry points
int32_t upload_queries(vers_ctx* ctx, const float* q, uint32_t nq, uint32_t stride, uint32_t dim, uint32_t ld,
                       float** d_q);
}

extern "C" int32_t vers_flat_search(vers_dataset* ds, const float* queries, uint32_t nq, uint32_t q_stride_floats,
                                    uint32_t top_k, uint32_t metric, uint64_t* ids, float* dists, uint32_t* counts) {
    if (!ds || (!queries && nq) || (!ids && nq && top_k) || (!dists && nq && top_k))
        return fail(VERS_ERR_ARG, "flat_search: null argument");
    if (q_stride_floats < ds->dim) return fail(VERS_ERR_ARG, "flat_search: query stride < dim");
    if (top_k > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "top_k %u > %u", top_k, VERS_MAX_TOPK);
    if (nq == 0) return VERS_OK;
    if (top_k == 0) {
        if (counts) memset(counts, 0, sizeof(uint32_t) * nq);
        return VERS_OK;
    }
    vers_ctx* ctx = ds->ctx;
    VERS_CUDA(cudaSetDevice(ctx->device));
    const size_t nk = (size_t)nq * top_k;
    float* d_q;
    uint64_t* d_ids;
    float* d_d;
    uint32_t* d_c;
    if (metric > VERS_METRIC_COSINE) return fail(VERS_ERR_ARG, "flat_search: unknown metric %u", metric);
    std::lock_guard<std::mutex> lk(ctx->mu);
    {   // device staging from the context's grow-only I/O arena: no allocation on the steady-state path
        ScratchCarver plan(nullptr);
        plan.plan<float>((size_t)nq * ds->ld);
        plan.plan<uint64_t>(nk);
        plan.plan<float>(nk);
        plan.plan<uint32_t>(nq);
        VERS_TRY(io_reserve(ctx, plan.off + 256));
        ScratchCarver io(ctx->io);
        d_q = io.take<float>((size_t)nq * ds->ld);
        d_ids = io.take<uint64_t>(nk);
        d_d = io.take<float>(nk);
        d_c = io.take<uint32_t>(nq);
        if (q_stride_floats == ds->ld && ds->ld == ds->dim) {
            VERS_CUDA(cudaMemcpyAsync(d_q, queries, (size_t)nq * ds->ld * 4, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            if (ds->ld != ds->dim) VERS_CUDA(cudaMemsetAsync(d_q, 0, (size_t)nq * ds->ld * 4, ctx->stream));
            VERS_CUDA(cudaMemcpy2DAsync(d_q, (size_t)ds->ld * 4, queries, (size_t)q_stride_floats * 4,
                                        (size_t)ds->dim * 4, nq, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    int32_t rc = flat_search_dev_locked(ds, d_q, nq, top_k, metric, d_ids, d_d, d_c);
    if (rc == VERS_OK) {
        cudaError_t e = cudaMemcpyAsync(ids, d_ids, nk * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dists, d_d, nk * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && counts)
            e = cudaMemcpyAsync(counts, d_c, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "flat_search copy back: %s", cudaGetErrorString(e));
    }
    return rc;
}

extern "C" int32_t vers_topk_merge_dev(vers_ctx* ctx, const uint64_t* d_ids_all, const float* d_dists_all,
                                       uint32_t parts, uint64_t part_stride_ids, uint64_t part_stride_dists,
                                       uint32_t nq, uint32_t top_k, uint64_t* d_ids, float* d_dists,
                                       uint32_t* d_counts) {
    if (!ctx || !d_ids_all || !d_dists_all || !d_ids || !d_dists) return fail(VERS_ERR_ARG, "topk_merge_dev: null");
    if (top_k > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "top_k %u > %u", top_k, VERS_MAX_TOPK);
    if (nq == 0 || top_k == 0) return VERS_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    merge_ids_kernel<<<(unsigned)ceil_div(nq, MERGE_WARPS), MERGE_WARPS * 32, (size_t)MERGE_WARPS * top_k * 12,
                       ctx->stream>>>(d_ids_all, d_dists_all, parts, part_stride_ids ? part_stride_ids : (uint64_t)nq * top_k,
                                      part_stride_dists ? part_stride_dists : (uint64_t)nq * top_k, nq, top_k, d_ids,
                                      d_dists, d_counts);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}
