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

// =====================================================================================================================
// flat_stream_kernel — the single-query / small-batch scan (nq <= 8): the reference's own API shape (one query per
// call, utils.rs:68-82; ivfflat.rs:153).  The tile engine above pads one query to 8 columns and stages 32-float
// k-chunks with cp.async; here the table is STREAMED: one producer warp issues ONE bulk copy (cp.async.bulk, SASS
// UBLKCP) per row tile — rows are contiguous, so a tile is a single linear rows x ld x 4-byte transfer — into a ring
// of six shared-memory stages (mbarrier full/empty).  Three consumer GROUPS own two stages each (one being computed,
// one in flight), lane = row, and every lane walks ITS row's dimensions in order with one exact-order chain per query
// (rounded sub/mul/add, no FMA: base.rs:91-93, :119-126) — independent chains per lane, no padding columns, queries
// read from shared memory.  For 4 and 8 queries a group is two warps that split the queries (the fp32 pipe, not HBM,
// is the bound there: 3 instructions per row, query and dimension).
// A stage is always consumed by the same group, so every waiter observes every phase of its barriers in order (a
// parity wait cannot tell "not yet" from "two phases ago").
// Bank conflicts: a lane reads its row with LDS.128 at stride ld/4 float4s; when that stride is even the 8 lanes of a
// phase would collide, so lane i runs i mod 8 float4s BEHIND (its chain is still strictly sequential; the loop runs 7
// extra iterations) — then the 8 lanes of a phase always hit 8 distinct 16-byte bank groups, for the row AND for the
// query.  Selection: per (consumer warp, query) sorted top-k list in shared memory, ballot + insertion like the merge
// kernels; the lists of a query are folded per CTA at the end, the per-CTA lists go through merge_topk_kernel.
constexpr int FS_GROUPS = 3, FS_STAGES = 2 * FS_GROUPS, FS_MAX_ROWS = 32;

struct FlatStreamParams {
    const float* rows;     // [n][ld]
    uint64_t n;
    uint32_t ld;
    const float* queries;  // [nq][ld]
    uint32_t k, kpad, tile_rows;
    float* part_d;         // [nq][gridDim.x][k]
    uint32_t* part_p;
    float* dense_out;      // optional [NQ][n]: every distance is written out, nothing is selected (small-batch probe)
};

__device__ __forceinline__ bool fs_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t fs_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fs_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fs_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fs_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(fs_smem(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// one float4 of the row against the same float4 of NQW queries
template <int NQW, int OP>
__device__ __forceinline__ void fs_step(float (&acc)[NQW], const float4* __restrict__ x4, const float4* __restrict__ q4,
                                        uint32_t nf4, uint32_t j) {
    const float4 x = x4[j];
#pragma unroll
    for (int q = 0; q < NQW; ++q) {
        const float4 b = q4[(uint32_t)q * nf4 + j];
        pair_step<OP>(acc[q], x.x, b.x);
        pair_step<OP>(acc[q], x.y, b.y);
        pair_step<OP>(acc[q], x.z, b.z);
        pair_step<OP>(acc[q], x.w, b.w);
    }
}

// fold (v, pp) of the live lanes into the sorted list (sd, sp) of k entries
__device__ __forceinline__ void fs_fold(float* sd, uint32_t* sp, uint32_t k, float v, uint32_t pp, bool live, int lane) {
    while (true) {
        const bool pass = live && entry_less<uint32_t>(v, pp, sd[k - 1], sp[k - 1]);
        const unsigned m = __ballot_sync(FULL_MASK, pass);
        if (!m) break;
        const int src = __ffs(m) - 1;
        const float bv = __shfl_sync(FULL_MASK, v, src);
        const uint32_t bp = __shfl_sync(FULL_MASK, pp, src);
        warp_topk_insert<uint32_t>(sd, sp, (int)k, bv, bp, lane);
        if (lane == src) live = false;
    }
}

template <int NQW, int NH, int OP>  // NQW queries per consumer warp, NH warps per group: NQW * NH queries
__global__ void __launch_bounds__((FS_GROUPS * NH + 1) * 32, 1) flat_stream_kernel(FlatStreamParams p) {
    constexpr int NQ = NQW * NH, CONS = FS_GROUPS * NH, THREADS = (CONS + 1) * 32;
    extern __shared__ __align__(128) unsigned char fs_raw[];
    const uint32_t nf4 = p.ld >> 2;
    const uint32_t tile_bytes = p.tile_rows * p.ld * 4;
    unsigned char* tiles = fs_raw;
    float* qs = reinterpret_cast<float*>(fs_raw + (size_t)FS_STAGES * tile_bytes);
    float* list_d = qs + NQ * p.ld;
    uint32_t* list_p = reinterpret_cast<uint32_t*>(list_d + CONS * NQW * p.kpad);
    uint64_t* full = reinterpret_cast<uint64_t*>(list_p + CONS * NQW * p.kpad);
    uint64_t* empty = full + FS_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t ntiles = (p.n + p.tile_rows - 1) / p.tile_rows;
    const uint64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < FS_STAGES; ++s) {
            fs_mbar_init(&full[s], 1);
            fs_mbar_init(&empty[s], NH * 32);  // every consumer lane arrives itself (its own reads ordered by its own arrive)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = threadIdx.x; i < NQ * p.ld; i += THREADS) qs[i] = p.queries[i];
    for (uint32_t i = threadIdx.x; i < CONS * NQW * p.kpad; i += THREADS) {
        list_d[i] = __int_as_float(0x7f800000);
        list_p[i] = 0xffffffffu;
    }
    __syncthreads();

    if (warp == CONS) {
        // producer: one linear bulk copy per tile
        for (uint64_t i = 0; i < my_tiles; ++i) {
            const uint32_t stage = (uint32_t)(i % FS_STAGES), use = (uint32_t)(i / FS_STAGES);
            fs_mbar_wait(&empty[stage], (use & 1u) ^ 1u);
            if (fs_elect_one()) {
                const uint64_t r0 = (blockIdx.x + i * gridDim.x) * p.tile_rows;
                const uint32_t bytes = (uint32_t)(min((uint64_t)p.tile_rows, p.n - r0) * p.ld * 4);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fs_smem(&full[stage])), "r"(bytes)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 fs_smem(tiles + (size_t)stage * tile_bytes)),
                             "l"(p.rows + r0 * p.ld), "r"(bytes), "r"(fs_smem(&full[stage]))
                             : "memory");
            }
            __syncwarp();
        }
    } else {
        const int group = warp % FS_GROUPS, half = warp / FS_GROUPS;
        const uint32_t sg = (nf4 & 1u) ? 0u : (uint32_t)(lane & 7);  // stagger (float4s) when the row stride is even
        const float4* q4 = reinterpret_cast<const float4*>(qs) + (size_t)half * NQW * nf4;
        float* my_d = list_d + (size_t)warp * NQW * p.kpad;
        uint32_t* my_p = list_p + (size_t)warp * NQW * p.kpad;
        for (uint64_t i = group; i < my_tiles; i += FS_GROUPS) {  // tile i lives in stage i % 6: stages g and g + 3
            const uint32_t stage = (uint32_t)(i % FS_STAGES), use = (uint32_t)(i / FS_STAGES);
            const uint64_t r0 = (blockIdx.x + i * gridDim.x) * p.tile_rows;
            fs_mbar_wait(&full[stage], use & 1u);
            const bool has_row = (uint32_t)lane < p.tile_rows;
            const float4* x4 = reinterpret_cast<const float4*>(tiles + (size_t)stage * tile_bytes) +
                               (size_t)(has_row ? lane : 0) * nf4;
            float acc[NQW];
#pragma unroll
            for (int q = 0; q < NQW; ++q) acc[q] = 0.0f;
            if (nf4 & 1u) {  // odd stride, no stagger: every lane walks j = 0 .. nf4-1 together
#pragma unroll 4
                for (uint32_t j = 0; j < nf4; ++j) fs_step<NQW, OP>(acc, x4, q4, nf4, j);
            } else {
                // staggered: lane runs sg float4s behind; head and tail iterations are predicated, the body is not
                const uint32_t head = min(7u, nf4);
                for (uint32_t t = 0; t < head; ++t)
                    if (t >= sg) fs_step<NQW, OP>(acc, x4, q4, nf4, t - sg);
#pragma unroll 4
                for (uint32_t t = head; t < nf4; ++t) fs_step<NQW, OP>(acc, x4, q4, nf4, t - sg);  // t >= 7 >= sg
                for (uint32_t t = nf4; t < nf4 + 7; ++t)
                    if (t >= sg && t - sg < nf4) fs_step<NQW, OP>(acc, x4, q4, nf4, t - sg);
            }
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fs_smem(&empty[stage])) : "memory");
            const uint64_t row = r0 + (uint64_t)lane;
#pragma unroll
            for (int q = 0; q < NQW; ++q) {
                float v = acc[q];
                if (OP == OP_DOT) v = __fsub_rn(1.0f, v);  // cosine distance 1 - dot (base.rs:155)
                if (p.dense_out) {
                    if (has_row && row < p.n) p.dense_out[(uint64_t)(half * NQW + q) * p.n + row] = v;
                } else {
                    fs_fold(my_d + (size_t)q * p.kpad, my_p + (size_t)q * p.kpad, p.k, v, (uint32_t)row, has_row && row < p.n, lane);
                }
            }
        }
    }
    __syncthreads();
    // fold the groups' lists of every query into group 0's, write the CTA's partial result
    if (!p.dense_out && warp < CONS && warp % FS_GROUPS == 0) {
        const int half = warp / FS_GROUPS;
        for (int ql = 0; ql < NQW; ++ql) {
            float* sd = list_d + ((size_t)warp * NQW + ql) * p.kpad;
            uint32_t* sp = list_p + ((size_t)warp * NQW + ql) * p.kpad;
            for (int g = 1; g < FS_GROUPS; ++g) {
                const float* od = list_d + ((size_t)(warp + g) * NQW + ql) * p.kpad;
                const uint32_t* op = list_p + ((size_t)(warp + g) * NQW + ql) * p.kpad;
                for (uint32_t e0 = 0; e0 < p.k; e0 += 32) {
                    const uint32_t e = e0 + lane;
                    const float v = e < p.k ? od[e] : 0.f;
                    const uint32_t pp = e < p.k ? op[e] : 0xffffffffu;
                    fs_fold(sd, sp, p.k, v, pp, pp != 0xffffffffu, lane);
                }
            }
            __syncwarp();
            const uint64_t base = ((uint64_t)(half * NQW + ql) * gridDim.x + blockIdx.x) * p.k;
            for (uint32_t e = lane; e < p.k; e += 32) {
                p.part_d[base + e] = sd[e];
                p.part_p[base + e] = sp[e];
            }
        }
    }
}

template <int NQW, int NH>
static int32_t launch_flat_stream(vers_ctx* ctx, const FlatStreamParams& p, uint32_t metric, unsigned grid, size_t smem) {
    constexpr int THREADS = (FS_GROUPS * NH + 1) * 32;
    if (metric == VERS_METRIC_L2SQ) {
        auto kern = flat_stream_kernel<NQW, NH, OP_L2SQ>;
        VERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, THREADS, smem, ctx->stream>>>(p);
    } else {
        auto kern = flat_stream_kernel<NQW, NH, OP_DOT>;
        VERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, THREADS, smem, ctx->stream>>>(p);
    }
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

// small-batch exhaustive scan of a contiguous row table; returns VERS_ERR_UNSUPPORTED when the shape does not fit
// (the caller then runs the tile engine).  Caller holds ctx->mu.
int32_t flat_stream_search(vers_ctx* ctx, const float* rows, uint64_t n, uint32_t ld, const float* d_queries, uint32_t nq,
                           uint32_t k, uint32_t metric, uint64_t id_base, uint64_t* d_ids, float* d_d, uint32_t* d_cnt,
                           int family) {
    if (nq == 0 || nq > 8 || k == 0 || k > VERS_MAX_TOPK || n == 0 || n >= 0xffffffffull) return VERS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(rows) & 15) != 0) return VERS_ERR_UNSUPPORTED;  // bulk copies need 16-byte alignment
    const uint32_t NQ = nq == 1 ? 1 : nq == 2 ? 2 : nq <= 4 ? 4 : 8;
    // queries per consumer warp x warps per group.  More warps rather than more queries per warp: the exact-order chains
    // are latency-bound below ~3 warps per scheduler (measured on 1M x 300, ms per scan: 2 queries 0.222 -> 0.198 with
    // <1, 2> instead of <2, 1>; 4 queries 0.303 -> 0.279 with <1, 4>; 8 queries 0.531 -> 0.443 with <2, 4>, 0.524 with <1, 8>)
    const uint32_t NH = NQ >= 4 ? 4 : NQ, NQW = NQ / NH, CONS = FS_GROUPS * NH;
    const uint32_t kpad = round_up(k, 32);
    const size_t fixed = (size_t)NQ * ld * 4 + (size_t)CONS * NQW * kpad * 8 + 2 * FS_STAGES * 8 + 128;
    const size_t budget = 227 * 1024;
    if (fixed + (size_t)FS_STAGES * 4 * ld * 4 > budget) return VERS_ERR_UNSUPPORTED;  // not even 4-row tiles fit
    const uint32_t tile_rows = (uint32_t)std::min<size_t>(FS_MAX_ROWS, (budget - fixed) / ((size_t)FS_STAGES * ld * 4));
    const size_t smem = (size_t)FS_STAGES * tile_rows * ld * 4 + fixed;
    const uint64_t ntiles = ceil_div(n, tile_rows);
    const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)ctx->sm_count);
    // queries padded to NQ rows (zero rows: their results are not merged), partial lists [NQ][grid][k]
    ScratchCarver plan(nullptr);
    plan.plan<float>((size_t)NQ * ld);
    plan.plan<float>((size_t)NQ * grid * k);
    plan.plan<uint32_t>((size_t)NQ * grid * k);
    VERS_TRY(scratch_reserve(ctx, plan.off + 256));
    ScratchCarver sc(ctx->scratch);
    float* qpad = sc.take<float>((size_t)NQ * ld);
    FlatStreamParams p;
    p.rows = rows;
    p.n = n;
    p.ld = ld;
    p.k = k;
    p.kpad = kpad;
    p.tile_rows = tile_rows;
    p.part_d = sc.take<float>((size_t)NQ * grid * k);
    p.part_p = sc.take<uint32_t>((size_t)NQ * grid * k);
    p.dense_out = nullptr;
    p.queries = d_queries;
    if (NQ != nq) {
        VERS_CUDA(cudaMemsetAsync(qpad, 0, (size_t)NQ * ld * 4, ctx->stream));
        VERS_CUDA(cudaMemcpyAsync(qpad, d_queries, (size_t)nq * ld * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        p.queries = qpad;
    }
    {
        FamilyTimer ft(ctx, family);
        switch (NQ) {
            case 1: VERS_TRY((launch_flat_stream<1, 1>(ctx, p, metric, grid, smem))); break;
            case 2: VERS_TRY((launch_flat_stream<1, 2>(ctx, p, metric, grid, smem))); break;
            case 4: VERS_TRY((launch_flat_stream<1, 4>(ctx, p, metric, grid, smem))); break;
            default: VERS_TRY((launch_flat_stream<2, 4>(ctx, p, metric, grid, smem))); break;
        }
    }
    MergeParams mp;
    mp.part_d = p.part_d;
    mp.part_p = p.part_p;
    mp.seg = nullptr;
    mp.seg_scale = 1;
    mp.seg_stride = 1;
    mp.per_query = (uint64_t)grid * k;
    mp.map = nullptr;
    mp.id_base = id_base;
    mp.nq = nq;
    mp.k = k;
    mp.out_ids = d_ids;
    mp.out_d = d_d;
    mp.out_cnt = d_cnt;
    mp.qmask = nullptr;
    return launch_merge(ctx, mp);
}

// Every exact distance of nq <= 8 queries to the n rows of a table, dense_out [NQ][n] (NQ = nq rounded up to 1, 2, 4, 8;
// rows of the padding queries are garbage), with the streaming kernel above.  qpad: [8][ld] floats of scratch.  Returns
// VERS_ERR_UNSUPPORTED when the shape does not fit the kernel (the caller then takes the tile engine).  Caller holds
// ctx->mu and owns the scratch.
bool flat_stream_fits(uint32_t ld, uint32_t nq) {
    if (nq == 0 || nq > 8) return false;
    const uint32_t NQ = nq == 1 ? 1 : nq == 2 ? 2 : nq <= 4 ? 4 : 8;
    const uint32_t NH = NQ >= 4 ? 4 : NQ, NQW = NQ / NH, CONS = FS_GROUPS * NH;
    const size_t fixed = (size_t)NQ * ld * 4 + (size_t)CONS * NQW * 32 * 8 + 2 * FS_STAGES * 8 + 128;
    return fixed + (size_t)FS_STAGES * 4 * ld * 4 <= (size_t)227 * 1024;
}
int32_t flat_stream_dense(vers_ctx* ctx, const float* rows, uint64_t n, uint32_t ld, const float* d_queries, uint32_t nq,
                          float* qpad, float* dense_out, int family) {
    if (!flat_stream_fits(ld, nq) || n == 0 || n >= 0xffffffffull || (reinterpret_cast<uintptr_t>(rows) & 15) != 0)
        return VERS_ERR_UNSUPPORTED;
    const uint32_t NQ = nq == 1 ? 1 : nq == 2 ? 2 : nq <= 4 ? 4 : 8;
    const uint32_t NH = NQ >= 4 ? 4 : NQ, NQW = NQ / NH, CONS = FS_GROUPS * NH;
    const uint32_t kpad = 32;
    const size_t fixed = (size_t)NQ * ld * 4 + (size_t)CONS * NQW * kpad * 8 + 2 * FS_STAGES * 8 + 128;
    const size_t budget = 227 * 1024;
    const uint32_t tile_rows = (uint32_t)std::min<size_t>(FS_MAX_ROWS, (budget - fixed) / ((size_t)FS_STAGES * ld * 4));
    const size_t smem = (size_t)FS_STAGES * tile_rows * ld * 4 + fixed;
    const unsigned grid = (unsigned)std::min<uint64_t>(ceil_div(n, tile_rows), (uint64_t)ctx->sm_count);
    FlatStreamParams p;
    p.rows = rows;
    p.n = n;
    p.ld = ld;
    p.k = 1;
    p.kpad = kpad;
    p.tile_rows = tile_rows;
    p.part_d = nullptr;
    p.part_p = nullptr;
    p.dense_out = dense_out;
    p.queries = d_queries;
    if (NQ != nq) {
        VERS_CUDA(cudaMemsetAsync(qpad, 0, (size_t)NQ * ld * 4, ctx->stream));
        VERS_CUDA(cudaMemcpyAsync(qpad, d_queries, (size_t)nq * ld * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        p.queries = qpad;
    }
    FamilyTimer ft(ctx, family);
    switch (NQ) {
        case 1: VERS_TRY((launch_flat_stream<1, 1>(ctx, p, VERS_METRIC_L2SQ, grid, smem))); break;
        case 2: VERS_TRY((launch_flat_stream<1, 2>(ctx, p, VERS_METRIC_L2SQ, grid, smem))); break;
        case 4: VERS_TRY((launch_flat_stream<1, 4>(ctx, p, VERS_METRIC_L2SQ, grid, smem))); break;
        default: VERS_TRY((launch_flat_stream<2, 4>(ctx, p, VERS_METRIC_L2SQ, grid, smem))); break;
    }
    return VERS_OK;
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
    // the next 32 entries are requested before the current ones are folded (a warp walking thousands of entries with a
    // trip to memory per round was 0.5 ms of the one-query call); the id behind an entry is only looked up when its
    // distance can still enter the list
    uint32_t pp_n = 0xffffffffu;
    float v_n = 0.f;
    if (beg + lane < end) {
        pp_n = p.part_p[beg + lane];
        v_n = p.part_d[beg + lane];
    }
    for (uint64_t e0 = beg; e0 < end; e0 += 32) {
        const uint32_t pp = pp_n;
        const float v = v_n;
        pp_n = 0xffffffffu;
        if (e0 + 32 + lane < end) {
            pp_n = p.part_p[e0 + 32 + lane];
            v_n = p.part_d[e0 + 32 + lane];
        }
        uint64_t id = 0;
        bool live = pp != 0xffffffffu && !(v > sd[k - 1]);
        if (__any_sync(FULL_MASK, live)) {
            if (live) id = p.map ? p.map[pp] : p.id_base + pp;
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
    if (metric == VERS_METRIC_L2SQ && ds->flat_mode != 1 && nq >= tc_min_batch() && top_k >= 1 && top_k <= 64) {
        // large batches: tensor-core candidate keys -> exact-order rerank -> certificate -> exact redo (ivf.cu)
        if (!ds->d_norm) {
            // all or nothing; ||row||^2 padded with +inf up to a multiple of 64 rows (tc_flat_kernel reads whole tiles)
            float* norm = nullptr;
            uint32_t* nmax = nullptr;
            unsigned long long* stats = nullptr;
            cudaError_t e = cudaMalloc(&norm, ((size_t)ds->n + 64) * 4);
            if (e == cudaSuccess) e = cudaMalloc(&nmax, 4);
            if (e == cudaSuccess) e = cudaMalloc(&stats, 128);
            if (e == cudaSuccess) {
                std::vector<float> inf(64, __builtin_inff());
                e = cudaMemcpyAsync(norm + ds->n, inf.data(), 64 * 4, cudaMemcpyHostToDevice, ctx->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            }
            if (e != cudaSuccess) {
                cudaFree(norm), cudaFree(nmax), cudaFree(stats);
                return fail(e == cudaErrorMemoryAllocation ? VERS_ERR_NOMEM : VERS_ERR_CUDA, "flat search buffers: %s",
                            cudaGetErrorString(e));
            }
            ds->d_norm = norm, ds->d_nmax = nmax, ds->d_stats = stats;
            VERS_TRY(launch_rownorm(ctx, ds->d_rows, ds->ld, ds->n, ds->d_norm, ds->d_nmax));
        }
        bool used = false;
        int32_t rc = flat_search_tc_plan_and_run(ctx, ds->d_rows, ds->n, ds->ld, ds->d_norm, ds->d_nmax, ds->id_base,
                                                 ds->d_stats, d_queries, nq, top_k, d_ids, d_dists, d_counts, &used, ds->flat_mode,
                                                 &ds->d_tiles);
        if (used || rc != VERS_ERR_UNSUPPORTED) return rc;
    }
    if (ds->d_stats) VERS_CUDA(cudaMemsetAsync(ds->d_stats, 0, 64, ctx->stream));  // the exact-order engine ran
    if (nq <= 8 && ds->flat_mode != 1) {  // single query / small batch: the streaming kernel (exact order as well)
        int32_t rc = flat_stream_search(ctx, ds->d_rows, ds->n, ds->ld, d_queries, nq, top_k, metric, ds->id_base, d_ids,
                                        d_dists, d_counts, KF_FLAT_SCAN);
        if (rc != VERS_ERR_UNSUPPORTED) return rc;
    }
    RowSrc A{ds->d_rows, nullptr, ds->ld, ds->n};
    RowSrc B{d_queries, nullptr, ds->ld, nq};
    return scan_topk_dev(ctx, A, B, nq, ds->ld, top_k, metric, nullptr, ds->id_base, d_ids, d_dists, d_counts,
                         KF_FLAT_SCAN);
}

extern "C" int32_t vers_flat_search_dev(vers_dataset* ds, const float* d_queries, uint32_t nq, uint32_t top_k,
                                        uint32_t metric, uint64_t* d_ids, float* d_dists, uint32_t* d_counts) {
    if (!ds || (!d_queries && nq) || !d_ids || !d_dists) return fail(VERS_ERR_ARG, "flat_search_dev: null argument");
    if (metric > VERS_METRIC_COSINE) return fail(VERS_ERR_ARG, "flat_search_dev: unknown metric %u", metric);
    std::lock_guard<std::recursive_mutex> lk(ds->ctx->mu);
    VERS_CUDA(cudaSetDevice(ds->ctx->device));
    return flat_search_dev_locked(ds, d_queries, nq, top_k, metric, d_ids, d_dists, d_counts);
}

extern "C" int32_t vers_flat_set_mode(vers_dataset* ds, int32_t mode) {
    if (!ds) return fail(VERS_ERR_ARG, "flat_set_mode: null");
    if (mode < 0 || mode > 2) return fail(VERS_ERR_ARG, "flat_set_mode: mode %d", mode);
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
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
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
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    merge_ids_kernel<<<(unsigned)ceil_div(nq, MERGE_WARPS), MERGE_WARPS * 32, (size_t)MERGE_WARPS * top_k * 12,
                       ctx->stream>>>(d_ids_all, d_dists_all, parts, part_stride_ids ? part_stride_ids : (uint64_t)nq * top_k,
                                      part_stride_dists ? part_stride_dists : (uint64_t)nq * top_k, nq, top_k, d_ids,
                                      d_dists, d_counts);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}
