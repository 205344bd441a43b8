// ivf.cu — IVFFlatIndex (indexes/ivfflat.rs) on one GPU.
//
// Device layout: LIST-MAJOR.  The rows of inverted list c are contiguous in d_lm (ascending id inside a list,
// exactly the order of `ids[c]` built at ivfflat.rs:123-127), so scanning a list is one dense HBM stream.
// d_lm_ids maps a list-major position back to the global id.  Lists carry slack capacity so Index::add
// (ivfflat.rs:200-213) appends in place.
//
// search_approximate for a batch:
//   probe      : exact-order l2sq of every query against every centroid + top-nprobe by (d, centroid)  [scan_topk_dev]
//   group      : (query, probe slot) pairs are bucketed by list on the device (count -> scan -> fill)
//   list scan  : a persistent kernel pulls work items (list, group of <= 8 queries, chunk of rows); each streams the
//                list rows once through shared memory for all the queries of the group (exact order), keeping a
//                private top-k per query
//   merge      : nprobe >= 1: per query global top-k by (d, id) over its partial lists
//                nprobe == 0: the reference's spill semantics (ivfflat.rs:163-197): lists are consumed in probe
//                order; every list but the last contributes all its rows (sorted), the last one the remainder.
#include <algorithm>

#include "kmeans.cuh"
#include "scan.cuh"

struct vers_ivf {
    vers_ctx* ctx = nullptr;
    uint32_t dim = 0, ld = 0, C = 0;
    uint64_t n = 0;        // rows in the index == assignments.len()
    uint64_t id_base = 0;  // global id of local row 0
    float* d_cents = nullptr;      // [C][ld]
    float* d_lm = nullptr;         // [cap_total][ld]
    uint64_t* d_lm_ids = nullptr;  // [cap_total]
    uint64_t cap_total = 0;
    std::vector<uint64_t> seg_off;  // [C] start of list c in d_lm
    std::vector<uint32_t> seg_len;  // [C]
    std::vector<uint32_t> seg_cap;  // [C]
    uint64_t* d_seg_off = nullptr;  // [C]
    uint32_t* d_seg_len = nullptr;  // [C]
    uint32_t* d_assign = nullptr;   // [n_built] assignments of the rows present at build time
    uint64_t n_built = 0;
    std::vector<uint32_t> assign_tail;  // assignments of rows added later
    float best_cost = 0.f;
    uint32_t best_attempt = 0;
    unsigned long long* d_stats = nullptr;  // [4] counters of the most recent search (vers_ivf_last_search_stats)
};

namespace vers {

constexpr uint32_t LIST_CHUNK_ROWS = 4096;  // rows per work item; multiple of NarrowCfg::TA
using ScanCfg = NarrowCfg;

// ---------------------------------------------------------------- layout
__global__ void gather_list_major_kernel(const float* __restrict__ rows, uint32_t ld,
                                         const uint32_t* __restrict__ sorted_rows, uint64_t n, uint64_t id_base,
                                         float* __restrict__ lm, uint64_t* __restrict__ lm_ids) {
    const uint32_t ld4 = ld >> 2;
    uint64_t total = n * ld4;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t j = i / ld4;
        uint32_t c4 = (uint32_t)(i - j * ld4);
        uint32_t src = sorted_rows[j];
        reinterpret_cast<float4*>(lm)[i] = reinterpret_cast<const float4*>(rows)[(uint64_t)src * ld4 + c4];
        if (c4 == 0) lm_ids[j] = id_base + src;
    }
}

// ---------------------------------------------------------------- grouping of (query, probe) pairs by list
struct GroupParams {
    const uint64_t* probe_ids;  // [nq][np]
    const uint32_t* seg_len;    // [C]
    const uint32_t* used;       // optional [nq]: only slots s < used[q] are active (reference spill mode)
    uint32_t nq, np, C;
    uint32_t* lq_cnt;     // [C]   queries per list
    uint32_t* pair_nch;   // [nq*np] chunks of the pair's list (0 for inactive pairs / empty lists)
    uint32_t* item_cnt;   // [C]   work items per list
    const uint64_t* lq_off;  // [C+1]
    uint32_t* cursor;     // [C]
    unsigned long long* stats;  // [4]
    uint32_t* lq_query;   // [npairs] query of each grouped pair
    uint32_t* lq_pair;    // [npairs] pair index q*np+s of each grouped pair
};

__global__ void group_count_kernel(GroupParams g) {
    uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= g.nq * g.np) return;
    uint32_t q = pi / g.np, s = pi % g.np;
    uint32_t nch = 0;
    if (!g.used || s < g.used[q]) {
        uint32_t l = (uint32_t)g.probe_ids[pi];
        uint32_t len = g.seg_len[l];
        nch = (len + LIST_CHUNK_ROWS - 1) / LIST_CHUNK_ROWS;
        if (nch) atomicAdd(&g.lq_cnt[l], 1u);
    }
    g.pair_nch[pi] = nch;
}

__global__ void group_items_kernel(GroupParams g) {
    uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= g.C) return;
    uint32_t m = g.lq_cnt[l];
    uint32_t nch = (g.seg_len[l] + LIST_CHUNK_ROWS - 1) / LIST_CHUNK_ROWS;
    uint32_t items = ((m + ScanCfg::TB - 1) / ScanCfg::TB) * nch;
    g.item_cnt[l] = items;
    if (m) {
        atomicAdd(&g.stats[0], (unsigned long long)g.seg_len[l]);
        atomicAdd(&g.stats[1], (unsigned long long)g.seg_len[l] * m);
        atomicAdd(&g.stats[2], (unsigned long long)items);
        atomicAdd(&g.stats[3], 1ull);
    }
}

__global__ void group_fill_kernel(GroupParams g) {
    uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= g.nq * g.np) return;
    if (g.pair_nch[pi] == 0) return;
    uint32_t l = (uint32_t)g.probe_ids[pi];
    uint32_t slot = atomicAdd(&g.cursor[l], 1u);
    uint64_t at = g.lq_off[l] + slot;
    g.lq_query[at] = pi / g.np;
    g.lq_pair[at] = pi;
}

// ---------------------------------------------------------------- list scan (the dominant kernel)
struct ListScanParams {
    const float* lm;
    const float* queries;
    uint32_t ld, C, k, kpad;
    const uint64_t* seg_off;
    const uint32_t* seg_len;
    const uint32_t* lq_query;
    const uint32_t* lq_pair;
    const uint64_t* lq_off;          // [C+1]
    const uint64_t* item_off;        // [C+1]
    const uint64_t* pair_chunk_off;  // [npairs+1]
    uint64_t nq;
    float* part_d;
    uint32_t* part_p;
    unsigned long long* counter;
};

__global__ void __launch_bounds__(ScanCfg::NT, 2) list_scan_kernel(ListScanParams p) {
    using Cfg = ScanCfg;
    extern __shared__ __align__(16) float smem[];
    float* list_d = smem + Cfg::TILE_FLOATS;
    uint32_t* list_p = reinterpret_cast<uint32_t*>(list_d + Cfg::NLISTS * p.kpad);
    __shared__ long long s_item;
    __shared__ uint32_t s_list;
    const uint64_t total_items = p.item_off[p.C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (true) {
        if (threadIdx.x == 0) {
            unsigned long long it = atomicAdd(p.counter, 1ull);
            if (it >= total_items) {
                s_item = -1;
            } else {
                // last l with item_off[l] <= it
                uint32_t lo = 0, hi = p.C;
                while (hi - lo > 1) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (p.item_off[mid] <= it) lo = mid; else hi = mid;
                }
                s_list = lo;
                s_item = (long long)(it - p.item_off[lo]);
            }
        }
        __syncthreads();
        const long long local = s_item;
        const uint32_t l = s_list;
        if (local < 0) break;
        const uint32_t len = p.seg_len[l];
        const uint32_t nch = (len + LIST_CHUNK_ROWS - 1) / LIST_CHUNK_ROWS;
        const uint32_t group = (uint32_t)(local / nch), chunk = (uint32_t)(local % nch);
        const uint64_t q0 = p.lq_off[l] + (uint64_t)group * Cfg::TB;
        const uint64_t m_l = p.lq_off[l + 1] - p.lq_off[l];
        const uint64_t nB = min((uint64_t)Cfg::TB, m_l - (uint64_t)group * Cfg::TB);
        const uint64_t base_pos = p.seg_off[l];
        RowSrc A{p.lm + base_pos * p.ld, nullptr, p.ld, len};
        RowSrc B{p.queries, p.lq_query + q0, p.ld, nB};
        const uint64_t r0 = (uint64_t)chunk * LIST_CHUNK_ROWS;
        const uint64_t r1 = min((uint64_t)len, r0 + LIST_CHUNK_ROWS);
        lists_init<Cfg>(list_d, list_p, p.kpad);
        for (uint64_t a0 = r0; a0 < r1; a0 += Cfg::TA) {
            float acc[Cfg::MA][Cfg::MB];
            tile_compute<Cfg, OP_L2SQ>(acc, A, a0, B, 0, p.ld, smem);
            tile_select_topk<Cfg, 0>(acc, a0, r1, 0, nB, p.k, p.kpad, list_d, list_p, base_pos);
        }
        __syncwarp();
        constexpr int SLOTS_PER_WARP = Cfg::TBS_PER_WARP * Cfg::MB;
        for (int s = 0; s < SLOTS_PER_WARP; ++s) {
            int slot = warp * SLOTS_PER_WARP + s, col, split;
            slot_to_col<Cfg>(slot, col, split);
            if ((uint64_t)col >= nB) continue;
            uint32_t pair = p.lq_pair[q0 + col];
            uint64_t base = ((p.pair_chunk_off[pair] + chunk) * Cfg::NSPLIT + split) * p.k;
            for (uint32_t e = lane; e < p.k; e += 32) {
                p.part_d[base + e] = list_d[slot * p.kpad + e];
                p.part_p[base + e] = list_p[slot * p.kpad + e];
            }
        }
        __syncthreads();  // s_item / lists are rewritten by the next iteration
    }
}

// ---------------------------------------------------------------- reference spill semantics (nprobe == 0)
// how many lists the reference opens for each query: lists are taken in probe order while the rows found so far
// are fewer than top_k (ivfflat.rs:168-195); depends only on the list lengths
__global__ void ref_plan_kernel(const uint64_t* __restrict__ probe_ids, const uint32_t* __restrict__ seg_len,
                                uint32_t nq, uint32_t np, uint32_t k, uint32_t* used, uint32_t* short_flag) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t remainder = k, m = 0;
    bool done = false;
    for (uint32_t s = 0; s < np && !done; ++s) {
        uint32_t len = seg_len[(uint32_t)probe_ids[(uint64_t)q * np + s]];
        m = s + 1;
        if (len < remainder) remainder -= len; else done = true;
    }
    used[q] = m;
    if (!done) atomicOr(short_flag, 1u);
}

// one warp per query: walk its pairs in probe order; merge the chunks of a pair into the pair's sorted top-k and
// emit the prefix the reference would take
__global__ void __launch_bounds__(128) ref_assemble_kernel(const float* __restrict__ part_d,
                                                          const uint32_t* __restrict__ part_p,
                                                          const uint64_t* __restrict__ pair_chunk_off,
                                                          const uint32_t* __restrict__ used,
                                                          const uint64_t* __restrict__ lm_ids, uint32_t nq,
                                                          uint32_t np, uint32_t k, uint32_t nsplit, uint64_t* out_ids,
                                                          float* out_d, uint32_t* out_cnt) {
    extern __shared__ __align__(16) unsigned char rsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * 4 + warp;
    if (q >= nq) return;
    uint32_t* sp = reinterpret_cast<uint32_t*>(rsm) + (size_t)warp * k;
    float* sd = reinterpret_cast<float*>(rsm + (size_t)4 * k * 4) + (size_t)warp * k;
    uint32_t written = 0, remainder = k;
    const uint32_t m = used[q];
    for (uint32_t s = 0; s < m && remainder > 0; ++s) {
        for (uint32_t e = lane; e < k; e += 32) {
            sd[e] = __int_as_float(0x7f800000);
            sp[e] = 0xffffffffu;
        }
        __syncwarp();
        uint64_t pi = (uint64_t)q * np + s;
        uint64_t beg = pair_chunk_off[pi] * nsplit * k, end = pair_chunk_off[pi + 1] * nsplit * k;
        for (uint64_t e0 = beg; e0 < end; e0 += 32) {
            uint64_t e = e0 + lane;
            float v = 0.f;
            uint32_t pp = 0xffffffffu;
            bool live = false;
            if (e < end) {
                pp = part_p[e];
                if (pp != 0xffffffffu) {
                    live = true;
                    v = part_d[e];
                }
            }
            while (true) {
                bool pass = live && entry_less<uint32_t>(v, pp, sd[k - 1], sp[k - 1]);
                unsigned mm = __ballot_sync(FULL_MASK, pass);
                if (!mm) break;
                int src = __ffs(mm) - 1;
                float bv = __shfl_sync(FULL_MASK, v, src);
                uint32_t bp = __shfl_sync(FULL_MASK, pp, src);
                warp_topk_insert<uint32_t>(sd, sp, (int)k, bv, bp, lane);
                if (lane == src) live = false;
            }
        }
        // number of valid entries in this pair's list = min(len, k)
        uint32_t have = 0;
        for (uint32_t e0 = 0; e0 < k; e0 += 32) {
            uint32_t e = e0 + lane;
            have += __popc(__ballot_sync(FULL_MASK, e < k && sp[e] != 0xffffffffu));
        }
        uint32_t take = have < remainder ? have : remainder;
        for (uint32_t e = lane; e < take; e += 32) {
            out_ids[(uint64_t)q * k + written + e] = lm_ids[sp[e]];
            out_d[(uint64_t)q * k + written + e] = sd[e];
        }
        written += take;
        remainder -= take;
        __syncwarp();
    }
    for (uint32_t e = written + lane; e < k; e += 32) {
        out_ids[(uint64_t)q * k + e] = 0xffffffffffffffffull;
        out_d[(uint64_t)q * k + e] = __int_as_float(0x7f800000);
    }
    if (out_cnt && lane == 0) out_cnt[q] = written;
}

// ---------------------------------------------------------------- host side
static int32_t ivf_upload_segments(vers_ivf* ivf) {
    VERS_CUDA(cudaMemcpyAsync(ivf->d_seg_off, ivf->seg_off.data(), (size_t)ivf->C * 8, cudaMemcpyHostToDevice,
                              ivf->ctx->stream));
    VERS_CUDA(cudaMemcpyAsync(ivf->d_seg_len, ivf->seg_len.data(), (size_t)ivf->C * 4, cudaMemcpyHostToDevice,
                              ivf->ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ivf->ctx->stream));
    return VERS_OK;
}

// builds the list-major mirror from a k-means state whose d_assign / d_cents are final
static int32_t ivf_from_state(vers_kmeans* km, float cost, uint32_t attempt, vers_ivf** out) {
    vers_dataset* ds = km->ds;
    vers_ctx* ctx = ds->ctx;
    vers_ivf* ivf = new vers_ivf();
    ivf->ctx = ctx;
    ivf->dim = ds->dim;
    ivf->ld = ds->ld;
    ivf->C = km->C;
    ivf->n = ds->n;
    ivf->n_built = ds->n;
    ivf->id_base = ds->id_base;
    ivf->best_cost = cost;
    ivf->best_attempt = attempt;
    *out = ivf;
    const size_t n1 = ds->n ? ds->n : 1;
    VERS_CUDA(cudaMalloc(&ivf->d_cents, (size_t)km->C * ds->ld * 4));
    VERS_CUDA(cudaMalloc(&ivf->d_lm, n1 * ds->ld * 4));
    VERS_CUDA(cudaMalloc(&ivf->d_lm_ids, n1 * 8));
    VERS_CUDA(cudaMalloc(&ivf->d_seg_off, (size_t)km->C * 8));
    VERS_CUDA(cudaMalloc(&ivf->d_seg_len, (size_t)km->C * 4));
    VERS_CUDA(cudaMalloc(&ivf->d_assign, n1 * 4));
    VERS_CUDA(cudaMalloc(&ivf->d_stats, 32));
    ivf->cap_total = ds->n;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        VERS_TRY(kmeans_build_csr(km));
        VERS_CUDA(cudaMemcpyAsync(ivf->d_cents, km->d_cents, (size_t)km->C * ds->ld * 4, cudaMemcpyDeviceToDevice,
                                  ctx->stream));
        if (ds->n) {
            VERS_CUDA(cudaMemcpyAsync(ivf->d_assign, km->d_assign, ds->n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            gather_list_major_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(
                ds->d_rows, ds->ld, km->d_sorted_rows, ds->n, ds->id_base, ivf->d_lm, ivf->d_lm_ids);
            VERS_LAUNCH_CHECK(ctx);
        }
        std::vector<uint64_t> off((size_t)km->C + 1);
        VERS_CUDA(cudaMemcpyAsync(off.data(), km->d_off, off.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        VERS_CUDA(cudaStreamSynchronize(ctx->stream));
        ivf->seg_off.resize(km->C);
        ivf->seg_len.resize(km->C);
        ivf->seg_cap.resize(km->C);
        for (uint32_t c = 0; c < km->C; ++c) {
            ivf->seg_off[c] = off[c];
            ivf->seg_len[c] = (uint32_t)(off[c + 1] - off[c]);
            ivf->seg_cap[c] = ivf->seg_len[c];
        }
    }
    return ivf_upload_segments(ivf);
}

// give every list slack capacity and move the rows (rare: only when an add hits a full list)
static int32_t ivf_relayout(vers_ivf* ivf) {
    vers_ctx* ctx = ivf->ctx;
    std::vector<uint64_t> noff(ivf->C);
    std::vector<uint32_t> ncap(ivf->C);
    uint64_t total = 0;
    for (uint32_t c = 0; c < ivf->C; ++c) {
        uint32_t len = ivf->seg_len[c];
        uint32_t cap = len + std::max<uint32_t>(32u, len / 8);
        noff[c] = total;
        ncap[c] = cap;
        total += cap;
    }
    if (total >= 0xffffffffull) return fail(VERS_ERR_UNSUPPORTED, "ivf: more than 2^32-2 list slots per GPU shard");
    float* nlm = nullptr;
    uint64_t* nids = nullptr;
    VERS_CUDA(cudaMalloc(&nlm, (size_t)total * ivf->ld * 4));
    cudaError_t e = cudaMalloc(&nids, (size_t)total * 8);
    if (e != cudaSuccess) {
        cudaFree(nlm);
        return fail(VERS_ERR_NOMEM, "ivf relayout: %s", cudaGetErrorString(e));
    }
    for (uint32_t c = 0; c < ivf->C && e == cudaSuccess; ++c) {
        uint32_t len = ivf->seg_len[c];
        if (!len) continue;
        e = cudaMemcpyAsync(nlm + noff[c] * ivf->ld, ivf->d_lm + ivf->seg_off[c] * ivf->ld, (size_t)len * ivf->ld * 4,
                            cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(nids + noff[c], ivf->d_lm_ids + ivf->seg_off[c], (size_t)len * 8,
                                cudaMemcpyDeviceToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(nlm);
        cudaFree(nids);
        return fail(VERS_ERR_CUDA, "ivf relayout: %s", cudaGetErrorString(e));
    }
    cudaFree(ivf->d_lm);
    cudaFree(ivf->d_lm_ids);
    ivf->d_lm = nlm;
    ivf->d_lm_ids = nids;
    ivf->cap_total = total;
    ivf->seg_off = noff;
    ivf->seg_cap = ncap;
    return ivf_upload_segments(ivf);
}

// upper bound on the number of (pair, chunk) partial lists one query can produce when it opens np lists
static uint64_t ivf_max_chunks_per_query(const vers_ivf* ivf, uint32_t np) {
    std::vector<uint32_t> nch(ivf->C);
    for (uint32_t c = 0; c < ivf->C; ++c) nch[c] = (ivf->seg_len[c] + LIST_CHUNK_ROWS - 1) / LIST_CHUNK_ROWS;
    np = std::min(np, ivf->C);
    std::partial_sort(nch.begin(), nch.begin() + np, nch.end(), std::greater<uint32_t>());
    uint64_t s = 0;
    for (uint32_t i = 0; i < np; ++i) s += nch[i];
    return s;
}

static int32_t ivf_search_dev_locked(vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t k, uint32_t nprobe,
                                     uint64_t* d_ids, float* d_d, uint32_t* d_cnt) {
    vers_ctx* ctx = ivf->ctx;
    const bool ref_mode = nprobe == 0;
    const uint32_t np = ref_mode ? std::min<uint32_t>(ivf->C, VERS_MAX_TOPK) : std::min<uint32_t>(nprobe, ivf->C);
    if (np > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "nprobe %u > %u", np, VERS_MAX_TOPK);
    const uint64_t npairs = (uint64_t)nq * np;
    const uint64_t max_chunks = (uint64_t)nq * ivf_max_chunks_per_query(ivf, np);
    const size_t entries = (size_t)std::max<uint64_t>(max_chunks, 1) * ScanCfg::NSPLIT * k;

    // the probe carves its partial buffers from the front of the arena, ours come after it
    const ScanPlan probe_plan = scan_topk_plan(ctx, ivf->C, nq, np);
    ScratchCarver plan(nullptr);
    const size_t probe_reserve = (probe_plan.bytes + 255) & ~size_t(255);
    plan.off = probe_reserve;
    plan.plan<uint64_t>(npairs);                   // probe ids
    plan.plan<float>(npairs);                      // probe dists
    plan.plan<uint32_t>(ivf->C);                   // lq_cnt
    plan.plan<uint32_t>(ivf->C);                   // cursor
    plan.plan<uint32_t>(ivf->C);                   // item_cnt
    plan.plan<uint32_t>(npairs);                   // pair_nch
    plan.plan<uint64_t>((size_t)ivf->C + 1);       // lq_off
    plan.plan<uint64_t>((size_t)ivf->C + 1);       // item_off
    plan.plan<uint64_t>(npairs + 1);               // pair_chunk_off
    plan.plan<uint32_t>(npairs);                   // lq_query
    plan.plan<uint32_t>(npairs);                   // lq_pair
    plan.plan<uint32_t>(nq);                       // used
    plan.plan<unsigned long long>(2);              // counter, flag
    plan.plan<float>(entries);
    plan.plan<uint32_t>(entries);
    VERS_TRY(scratch_reserve(ctx, plan.off + 4096));

    // 1. probe
    ScratchCarver sc(ctx->scratch);
    sc.off = probe_reserve;
    uint64_t* probe_ids = sc.take<uint64_t>(npairs);
    float* probe_d = sc.take<float>(npairs);
    uint32_t* lq_cnt = sc.take<uint32_t>(ivf->C);
    uint32_t* cursor = sc.take<uint32_t>(ivf->C);
    uint32_t* item_cnt = sc.take<uint32_t>(ivf->C);
    uint32_t* pair_nch = sc.take<uint32_t>(npairs);
    uint64_t* lq_off = sc.take<uint64_t>((size_t)ivf->C + 1);
    uint64_t* item_off = sc.take<uint64_t>((size_t)ivf->C + 1);
    uint64_t* pair_chunk_off = sc.take<uint64_t>(npairs + 1);
    uint32_t* lq_query = sc.take<uint32_t>(npairs);
    uint32_t* lq_pair = sc.take<uint32_t>(npairs);
    uint32_t* used = sc.take<uint32_t>(nq);
    unsigned long long* counter = sc.take<unsigned long long>(2);
    float* part_d = sc.take<float>(entries);
    uint32_t* part_p = sc.take<uint32_t>(entries);

    RowSrc CA{ivf->d_cents, nullptr, ivf->ld, ivf->C};
    RowSrc QB{d_queries, nullptr, ivf->ld, nq};
    VERS_TRY(scan_topk_run(ctx, probe_plan, ctx->scratch, CA, QB, nq, ivf->ld, np, VERS_METRIC_L2SQ, nullptr, 0,
                           probe_ids, probe_d, nullptr, KF_PROBE));

    // 2. group pairs by list
    VERS_CUDA(cudaMemsetAsync(lq_cnt, 0, (size_t)ivf->C * 4, ctx->stream));
    VERS_CUDA(cudaMemsetAsync(cursor, 0, (size_t)ivf->C * 4, ctx->stream));
    VERS_CUDA(cudaMemsetAsync(counter, 0, 16, ctx->stream));
    VERS_CUDA(cudaMemsetAsync(ivf->d_stats, 0, 32, ctx->stream));
    uint32_t* short_flag = reinterpret_cast<uint32_t*>(counter + 1);
    if (ref_mode) {
        ref_plan_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, ctx->stream>>>(probe_ids, ivf->d_seg_len, nq, np, k, used,
                                                                             short_flag);
        VERS_LAUNCH_CHECK(ctx);
    }
    GroupParams g;
    g.probe_ids = probe_ids;
    g.seg_len = ivf->d_seg_len;
    g.used = ref_mode ? used : nullptr;
    g.nq = nq;
    g.np = np;
    g.C = ivf->C;
    g.lq_cnt = lq_cnt;
    g.pair_nch = pair_nch;
    g.item_cnt = item_cnt;
    g.lq_off = lq_off;
    g.cursor = cursor;
    g.stats = ivf->d_stats;
    g.lq_query = lq_query;
    g.lq_pair = lq_pair;
    group_count_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, ctx->stream>>>(g);
    VERS_LAUNCH_CHECK(ctx);
    group_items_kernel<<<(unsigned)ceil_div(ivf->C, 256), 256, 0, ctx->stream>>>(g);
    VERS_LAUNCH_CHECK(ctx);
    VERS_TRY(launch_exclusive_scan(ctx, lq_cnt, ivf->C, lq_off));
    VERS_TRY(launch_exclusive_scan(ctx, item_cnt, ivf->C, item_off));
    VERS_TRY(launch_exclusive_scan(ctx, pair_nch, npairs, pair_chunk_off));
    group_fill_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, ctx->stream>>>(g);
    VERS_LAUNCH_CHECK(ctx);

    // 3. list scan
    ListScanParams lp;
    lp.lm = ivf->d_lm;
    lp.queries = d_queries;
    lp.ld = ivf->ld;
    lp.C = ivf->C;
    lp.k = k;
    lp.kpad = round_up(k, 32);
    lp.seg_off = ivf->d_seg_off;
    lp.seg_len = ivf->d_seg_len;
    lp.lq_query = lq_query;
    lp.lq_pair = lq_pair;
    lp.lq_off = lq_off;
    lp.item_off = item_off;
    lp.pair_chunk_off = pair_chunk_off;
    lp.nq = nq;
    lp.part_d = part_d;
    lp.part_p = part_p;
    lp.counter = counter;
    {
        size_t smem = scan_smem_bytes(ScanCfg::TILE_FLOATS, ScanCfg::NLISTS, lp.kpad);
        VERS_CUDA(cudaFuncSetAttribute(list_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FamilyTimer ft(ctx, KF_LIST_SCAN);
        list_scan_kernel<<<ctx->sm_count * 2, ScanCfg::NT, smem, ctx->stream>>>(lp);
        VERS_LAUNCH_CHECK(ctx);
    }

    // 4. merge
    if (!ref_mode) {
        MergeParams mp;
        mp.part_d = part_d;
        mp.part_p = part_p;
        mp.seg = pair_chunk_off;
        mp.seg_scale = (uint64_t)ScanCfg::NSPLIT * k;
        mp.seg_stride = np;
        mp.per_query = 0;
        mp.map = ivf->d_lm_ids;
        mp.id_base = 0;
        mp.nq = nq;
        mp.k = k;
        mp.out_ids = d_ids;
        mp.out_d = d_d;
        mp.out_cnt = d_cnt;
        return launch_merge(ctx, mp);
    }
    ref_assemble_kernel<<<(unsigned)ceil_div(nq, 4), 128, (size_t)4 * k * 8, ctx->stream>>>(
        part_d, part_p, pair_chunk_off, used, ivf->d_lm_ids, nq, np, k, ScanCfg::NSPLIT, d_ids, d_d, d_cnt);
    VERS_LAUNCH_CHECK(ctx);
    uint32_t flag = 0;
    VERS_CUDA(cudaMemcpyAsync(&flag, short_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        if (np == ivf->C)
            return fail(VERS_ERR_PANIC,
                        "ivf_search: fewer than top_k rows reachable (index out of bounds at ivfflat.rs:169)");
        return fail(VERS_ERR_UNSUPPORTED, "ivf_search: the %u nearest lists hold fewer than top_k rows", np);
    }
    return VERS_OK;
}

}  // namespace vers

using namespace vers;

extern "C" int32_t vers_ivf_free(vers_ivf* ivf) {
    if (!ivf) return VERS_OK;
    cudaSetDevice(ivf->ctx->device);
    cudaStreamSynchronize(ivf->ctx->stream);
    cudaFree(ivf->d_cents);
    cudaFree(ivf->d_lm);
    cudaFree(ivf->d_lm_ids);
    cudaFree(ivf->d_seg_off);
    cudaFree(ivf->d_seg_len);
    cudaFree(ivf->d_assign);
    cudaFree(ivf->d_stats);
    delete ivf;
    return VERS_OK;
}

extern "C" int32_t vers_ivf_from_kmeans(vers_kmeans* km, vers_ivf** out) {
    if (!km || !out) return fail(VERS_ERR_ARG, "ivf_from_kmeans: null argument");
    *out = nullptr;
    VERS_CUDA(cudaSetDevice(km->ds->ctx->device));
    int32_t rc = ivf_from_state(km, 0.f, 0, out);
    if (rc != VERS_OK) {
        vers_ivf_free(*out);
        *out = nullptr;
    }
    return rc;
}

extern "C" int32_t vers_ivf_build_index(vers_dataset* ds, uint32_t num_clusters, uint32_t num_attempts,
                                        uint32_t max_iterations, const uint64_t* init_rows, vers_ivf** out) {
    if (!ds || !out || (!init_rows && num_attempts)) return fail(VERS_ERR_ARG, "ivf_build_index: null argument");
    *out = nullptr;
    vers_ctx* ctx = ds->ctx;
    VERS_CUDA(cudaSetDevice(ctx->device));
    vers_kmeans* best = nullptr;
    float best_cost = __builtin_inff();
    uint32_t best_attempt = 0xffffffffu;
    int32_t rc = VERS_OK;
    for (uint32_t t = 0; t < num_attempts && rc == VERS_OK; ++t) {
        vers_kmeans* km = nullptr;
        rc = vers_kmeans_create(ds, num_clusters, &km);
        if (rc == VERS_OK) rc = vers_kmeans_init_from_rows(km, init_rows + (size_t)t * num_clusters);
        if (rc == VERS_OK) rc = vers_kmeans_fit(km, max_iterations, nullptr);
        float cost = 0.0f;
        if (rc == VERS_OK) rc = vers_kmeans_cost_step(km, &cost);
        if (rc == VERS_OK && cost < best_cost) {  // strict `<`, ivfflat.rs:116
            best_cost = cost;
            best_attempt = t;
            vers_kmeans_free(best);
            best = km;
        } else {
            vers_kmeans_free(km);
        }
    }
    if (rc == VERS_OK && !best)
        rc = fail(VERS_ERR_UNSUPPORTED,
                  "ivf_build_index: no attempt produced a finite cost (the reference would keep an empty index)");
    if (rc == VERS_OK) {
        rc = ivf_from_state(best, best_cost, best_attempt, out);
        if (rc != VERS_OK) {
            vers_ivf_free(*out);
            *out = nullptr;
        }
    }
    vers_kmeans_free(best);
    return rc;
}

extern "C" int32_t vers_ivf_from_parts(vers_dataset* ds, const float* centroids, uint32_t num_clusters,
                                       uint32_t stride_floats, const uint64_t* assignments, vers_ivf** out) {
    if (!ds || !centroids || !out) return fail(VERS_ERR_ARG, "ivf_from_parts: null argument");
    *out = nullptr;
    vers_ctx* ctx = ds->ctx;
    vers_kmeans* km = nullptr;
    VERS_TRY(vers_kmeans_create(ds, num_clusters, &km));
    int32_t rc = vers_kmeans_set_centroids(km, centroids, stride_floats);
    if (rc == VERS_OK) {
        if (assignments == nullptr) {
            rc = vers_kmeans_assign_step(km);
        } else if (ds->n) {
            // narrow u64 -> u32 on the host (load_index path, not hot)
            std::vector<uint32_t> a32(ds->n);
            for (uint64_t i = 0; i < ds->n && rc == VERS_OK; ++i) {
                if (assignments[i] >= num_clusters)
                    rc = fail(VERS_ERR_ARG, "ivf_from_parts: assignment %llu >= num_clusters",
                              (unsigned long long)assignments[i]);
                a32[i] = (uint32_t)assignments[i];
            }
            if (rc == VERS_OK) {
                cudaError_t e = cudaMemcpyAsync(km->d_assign, a32.data(), ds->n * 4, cudaMemcpyHostToDevice, ctx->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
                if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "ivf_from_parts: %s", cudaGetErrorString(e));
                km->csr_valid = false;
            }
        }
    }
    if (rc == VERS_OK) {
        rc = ivf_from_state(km, 0.f, 0, out);
        if (rc != VERS_OK) {
            vers_ivf_free(*out);
            *out = nullptr;
        }
    }
    vers_kmeans_free(km);
    return rc;
}

extern "C" int32_t vers_ivf_info(const vers_ivf* ivf, uint64_t* n, uint32_t* dim, uint32_t* num_clusters,
                                 float* best_cost, uint32_t* best_attempt) {
    if (!ivf) return fail(VERS_ERR_ARG, "ivf_info: null");
    if (n) *n = ivf->n;
    if (dim) *dim = ivf->dim;
    if (num_clusters) *num_clusters = ivf->C;
    if (best_cost) *best_cost = ivf->best_cost;
    if (best_attempt) *best_attempt = ivf->best_attempt;
    return VERS_OK;
}

extern "C" int32_t vers_ivf_get_centroids(const vers_ivf* ivf, float* centroids, uint32_t stride_floats) {
    if (!ivf || !centroids) return fail(VERS_ERR_ARG, "ivf_get_centroids: null argument");
    if (stride_floats < ivf->dim) return fail(VERS_ERR_ARG, "ivf_get_centroids: stride < dim");
    VERS_CUDA(cudaSetDevice(ivf->ctx->device));
    VERS_CUDA(cudaMemcpy2DAsync(centroids, (size_t)stride_floats * 4, ivf->d_cents, (size_t)ivf->ld * 4,
                                (size_t)ivf->dim * 4, ivf->C, cudaMemcpyDeviceToHost, ivf->ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ivf->ctx->stream));
    return VERS_OK;
}

extern "C" int32_t vers_ivf_get_assignments(const vers_ivf* ivf, uint64_t* assignments) {
    if (!ivf || (!assignments && ivf->n)) return fail(VERS_ERR_ARG, "ivf_get_assignments: null argument");
    VERS_CUDA(cudaSetDevice(ivf->ctx->device));
    if (ivf->n_built) {
        std::vector<uint32_t> a32(ivf->n_built);
        VERS_CUDA(cudaMemcpyAsync(a32.data(), ivf->d_assign, ivf->n_built * 4, cudaMemcpyDeviceToHost, ivf->ctx->stream));
        VERS_CUDA(cudaStreamSynchronize(ivf->ctx->stream));
        for (uint64_t i = 0; i < ivf->n_built; ++i) assignments[i] = a32[i];
    }
    for (size_t i = 0; i < ivf->assign_tail.size(); ++i) assignments[ivf->n_built + i] = ivf->assign_tail[i];
    return VERS_OK;
}

extern "C" int32_t vers_ivf_get_list_sizes(const vers_ivf* ivf, uint64_t* sizes) {
    if (!ivf || !sizes) return fail(VERS_ERR_ARG, "ivf_get_list_sizes: null argument");
    for (uint32_t c = 0; c < ivf->C; ++c) sizes[c] = ivf->seg_len[c];
    return VERS_OK;
}

extern "C" int32_t vers_ivf_search_dev(vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t top_k,
                                       uint32_t nprobe, uint64_t* d_ids, float* d_dists, uint32_t* d_counts) {
    if (!ivf || (!d_queries && nq) || !d_ids || !d_dists) return fail(VERS_ERR_ARG, "ivf_search_dev: null argument");
    if (top_k > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "top_k %u > %u", top_k, VERS_MAX_TOPK);
    if (nq == 0 || top_k == 0) return VERS_OK;
    std::lock_guard<std::mutex> lk(ivf->ctx->mu);
    VERS_CUDA(cudaSetDevice(ivf->ctx->device));
    return ivf_search_dev_locked(ivf, d_queries, nq, top_k, nprobe, d_ids, d_dists, d_counts);
}

namespace vers {
int32_t upload_queries(vers_ctx* ctx, const float* q, uint32_t nq, uint32_t stride, uint32_t dim, uint32_t ld,
                       float** d_q);
}

extern "C" int32_t vers_ivf_search(vers_ivf* ivf, const float* queries, uint32_t nq, uint32_t q_stride_floats,
                                   uint32_t top_k, uint32_t nprobe, uint64_t* ids, float* dists, uint32_t* counts) {
    if (!ivf || (!queries && nq) || (!ids && nq && top_k) || (!dists && nq && top_k))
        return fail(VERS_ERR_ARG, "ivf_search: null argument");
    if (q_stride_floats < ivf->dim) return fail(VERS_ERR_ARG, "ivf_search: query stride < dim");
    if (top_k > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "top_k %u > %u", top_k, VERS_MAX_TOPK);
    if (nq == 0) return VERS_OK;
    if (top_k == 0) {
        if (counts) memset(counts, 0, sizeof(uint32_t) * nq);
        return VERS_OK;
    }
    vers_ctx* ctx = ivf->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    const size_t nk = (size_t)nq * top_k;
    ScratchCarver plan(nullptr);
    plan.plan<float>((size_t)nq * ivf->ld);
    plan.plan<uint64_t>(nk);
    plan.plan<float>(nk);
    plan.plan<uint32_t>(nq);
    VERS_TRY(io_reserve(ctx, plan.off + 256));
    ScratchCarver io(ctx->io);
    float* d_q = io.take<float>((size_t)nq * ivf->ld);
    uint64_t* d_ids = io.take<uint64_t>(nk);
    float* d_d = io.take<float>(nk);
    uint32_t* d_c = io.take<uint32_t>(nq);
    if (ivf->ld != ivf->dim) VERS_CUDA(cudaMemsetAsync(d_q, 0, (size_t)nq * ivf->ld * 4, ctx->stream));
    VERS_CUDA(cudaMemcpy2DAsync(d_q, (size_t)ivf->ld * 4, queries, (size_t)q_stride_floats * 4, (size_t)ivf->dim * 4,
                                nq, cudaMemcpyHostToDevice, ctx->stream));
    VERS_TRY(ivf_search_dev_locked(ivf, d_q, nq, top_k, nprobe, d_ids, d_d, d_c));
    VERS_CUDA(cudaMemcpyAsync(ids, d_ids, nk * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaMemcpyAsync(dists, d_d, nk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) VERS_CUDA(cudaMemcpyAsync(counts, d_c, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    return VERS_OK;
}

extern "C" int32_t vers_ivf_get_list(const vers_ivf* ivf, uint32_t list, uint64_t* ids, float* rows,
                                     uint32_t stride_floats) {
    if (!ivf) return fail(VERS_ERR_ARG, "ivf_get_list: null");
    if (list >= ivf->C) return fail(VERS_ERR_ARG, "ivf_get_list: list %u >= %u", list, ivf->C);
    if (rows && stride_floats < ivf->dim) return fail(VERS_ERR_ARG, "ivf_get_list: stride < dim");
    const uint32_t len = ivf->seg_len[list];
    if (len == 0) return VERS_OK;
    vers_ctx* ctx = ivf->ctx;
    VERS_CUDA(cudaSetDevice(ctx->device));
    if (ids) VERS_CUDA(cudaMemcpyAsync(ids, ivf->d_lm_ids + ivf->seg_off[list], (size_t)len * 8, cudaMemcpyDeviceToHost,
                                       ctx->stream));
    if (rows)
        VERS_CUDA(cudaMemcpy2DAsync(rows, (size_t)stride_floats * 4, ivf->d_lm + ivf->seg_off[list] * ivf->ld,
                                    (size_t)ivf->ld * 4, (size_t)ivf->dim * 4, len, cudaMemcpyDeviceToHost,
                                    ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    return VERS_OK;
}

extern "C" int32_t vers_ivf_last_search_stats(const vers_ivf* ivf, uint64_t out[4]) {
    if (!ivf || !out) return fail(VERS_ERR_ARG, "ivf_last_search_stats: null");
    VERS_CUDA(cudaSetDevice(ivf->ctx->device));
    VERS_CUDA(cudaMemcpyAsync(out, ivf->d_stats, 32, cudaMemcpyDeviceToHost, ivf->ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ivf->ctx->stream));
    return VERS_OK;
}

extern "C" int32_t vers_ivf_add(vers_ivf* ivf, const float* embedding, uint64_t vec_id, uint64_t* assigned_id,
                                uint32_t* cluster) {
    (void)vec_id;  // the reference ignores the caller's id: `let vec_id = self.assignments.len()` (ivfflat.rs:209)
    if (!ivf || !embedding) return fail(VERS_ERR_ARG, "ivf_add: null argument");
    vers_ctx* ctx = ivf->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    float* d_row = nullptr;
    VERS_TRY(upload_queries(ctx, embedding, 1, ivf->dim, ivf->dim, ivf->ld, &d_row));
    uint64_t* d_best = nullptr;
    float* d_bd = nullptr;
    int32_t rc = VERS_OK;
    if (cudaMalloc(&d_best, 8) != cudaSuccess || cudaMalloc(&d_bd, 4) != cudaSuccess)
        rc = fail(VERS_ERR_NOMEM, "ivf_add: cudaMalloc");
    uint64_t best = 0;
    if (rc == VERS_OK) {
        // nearest centroid, first minimum on ties (min_by, ivfflat.rs:201-207) == top-1 by (d, centroid index)
        RowSrc CA{ivf->d_cents, nullptr, ivf->ld, ivf->C};
        RowSrc QB{d_row, nullptr, ivf->ld, 1};
        rc = scan_topk_dev(ctx, CA, QB, 1, ivf->ld, 1, VERS_METRIC_L2SQ, nullptr, 0, d_best, d_bd, nullptr, KF_PROBE);
    }
    if (rc == VERS_OK) {
        cudaError_t e = cudaMemcpyAsync(&best, d_best, 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "ivf_add: %s", cudaGetErrorString(e));
    }
    if (rc == VERS_OK) {
        uint32_t c = (uint32_t)best;
        if (ivf->seg_len[c] == ivf->seg_cap[c]) rc = ivf_relayout(ivf);
        if (rc == VERS_OK) {
            uint64_t pos = ivf->seg_off[c] + ivf->seg_len[c];
            uint64_t new_id = ivf->id_base + ivf->n;
            cudaError_t e = cudaMemcpyAsync(ivf->d_lm + pos * ivf->ld, d_row, (size_t)ivf->ld * 4,
                                            cudaMemcpyDeviceToDevice, ctx->stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(ivf->d_lm_ids + pos, &new_id, 8, cudaMemcpyHostToDevice, ctx->stream);
            ivf->seg_len[c] += 1;
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(ivf->d_seg_len + c, &ivf->seg_len[c], 4, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "ivf_add: %s", cudaGetErrorString(e));
            if (rc == VERS_OK) {
                ivf->assign_tail.push_back(c);
                if (assigned_id) *assigned_id = new_id;
                if (cluster) *cluster = c;
                ivf->n += 1;
            }
        }
    }
    cudaFree(d_row);
    cudaFree(d_best);
    cudaFree(d_bd);
    return rc;
}
