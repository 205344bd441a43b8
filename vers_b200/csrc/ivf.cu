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
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>

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
    unsigned long long* d_stats = nullptr;  // [16] counters of the most recent search (vers_ivf_last_search_stats)
    float* d_lm_norm = nullptr;             // [cap_total] ||row||^2 (any summation order; candidate pass only)
    uint32_t* d_nxmax = nullptr;            // [1] bit pattern of max ||row||^2 over the index (non-negative floats order as uints)
    float* d_cent_norm = nullptr;           // [C] ||centroid||^2 (any order; tensor-core probe only)
    uint32_t* d_ncmax = nullptr;            // [1] bit pattern of max ||centroid||^2
    int mode = 0;                           // 0 = candidate pass + exact rerank + certificate, 1 = exact-order everywhere
    // 16-bit candidate copy (mode 4): fp16(row * h16_scale), list-major like d_lm, ld16 = round_up(dim, 8) halfs per row;
    // rebuilt lazily by ivf_ensure_h16 when a bulk change invalidated it, updated in place by a single Index::add
    __half* d_lm16 = nullptr;
    uint64_t h16_cap = 0;                   // rows d_lm16 was allocated for
    uint32_t ld16 = 0;
    bool h16_valid = false;
    float h16_scale = 1.0f;                 // power of two: no element of a row with ||row||^2 <= h16_norm2_limit overflows
    double h16_norm2_limit = 0.0;
    uint32_t* d_h16_stat = nullptr;         // [0] bits of max ||x - x~||^2 over the live rows, [1] elements that did not fit
    vers::GraphCache call_graph;            // vers_ivf_search: the device work of a repeated call shape, as a CUDA graph
    // cache of ivf_max_chunks_per_query (host loop over the lists): valid while seg_epoch == mc_epoch
    uint64_t seg_epoch = 1, mc_epoch = 0;
    uint32_t mc_np = 0;
    uint64_t mc_val[3] = {0, 0, 0};
};

namespace vers {

constexpr uint32_t LIST_CHUNK_ROWS = 4096;  // rows per work item; multiple of NarrowCfg::TA
constexpr uint32_t LIST_CHUNK_ROWS_SMALL = 256;  // ... for batches of <= 8 queries (the reference's one-query call): a
                                                 // 2441-row list becomes 10 items instead of one CTA's 0.4 ms walk
constexpr uint32_t SMALL_BATCH = 8;
// tensor-core scan: the last sixth of the lists (the work items handed out last) is cut into small items so that the
// persistent CTAs drain together; everything before keeps whole-list items (one partial list per (query, list))
// (measured on the bench workload: 256..2048 rows and 1/10..1/3 of the lists are all within 1 % of each other)
constexpr uint32_t TC_TAIL_CHUNK_ROWS = 512;
inline uint32_t tc_tail_list0(uint32_t C) { return C - C / 6; }
// ... and when this GPU's share of the batch is small (many GPUs, few queries) every item shrinks so that each SM
// still gets >= ~4 of them (measured on the per-rank share of an 8-GPU step: 2048-row items 0.329 ms, 1024 0.343,
// 512 0.402, whole lists 0.346): rows the batch will stream ~ min(rows held, nq * nprobe * mean list length)
inline uint32_t tc_chunk_rows(const vers_ivf* ivf, uint32_t nq, uint32_t np) {
    static const int forced = getenv("VERS_TC_CHUNK_ROWS") ? atoi(getenv("VERS_TC_CHUNK_ROWS")) : 0;  // tuning knob
    if (forced >= (int)TC_TAIL_CHUNK_ROWS) return (uint32_t)forced / 128u * 128u;
    const double mean_len = ivf->C ? (double)ivf->n / ivf->C : 0.0;
    const double est_rows = std::min((double)ivf->n, (double)nq * np * mean_len);
    const double per_sm = est_rows / std::max(ivf->ctx->sm_count, 1);
    uint32_t cr = (uint32_t)(per_sm / 4.0) / 128u * 128u;  // >= ~4 full items per SM + the small tail items
    return std::min<uint32_t>(LIST_CHUNK_ROWS, std::max<uint32_t>(TC_TAIL_CHUNK_ROWS, cr));
}
using ScanCfg = NarrowCfg;

// everything about an index that a captured search depends on, folded into two words (graph cache keys)
void ivf_state_stamp(const vers_ivf* ivf, uint64_t out[4]) {
    out[0] = ivf->seg_epoch;
    out[1] = (uint64_t)ivf->mode | ((uint64_t)ivf->h16_valid << 8) | (ivf->cap_total << 16);
    out[2] = reinterpret_cast<uint64_t>(ivf->d_lm);
    out[3] = reinterpret_cast<uint64_t>(ivf->d_lm16);
}

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

// ||row||^2 for the candidate pass (one warp per row, lane-parallel FMA + shuffle tree: any order is fine here, the
// certificate's error bound covers every summation order) and the running maximum over the index
__global__ void rownorm_kernel(const float* __restrict__ lm, uint32_t ld, uint64_t pos0, uint64_t count,
                               float* __restrict__ norm, uint32_t* nxmax) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    float mx = 0.0f;
    for (uint64_t j = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); j < count; j += warps) {
        const float4* r = reinterpret_cast<const float4*>(lm + (pos0 + j) * ld);
        float s = 0.0f;
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            float4 v = r[c];
            s = __fmaf_rn(v.x, v.x, s);
            s = __fmaf_rn(v.y, v.y, s);
            s = __fmaf_rn(v.z, v.z, s);
            s = __fmaf_rn(v.w, v.w, s);
        }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
        if (lane == 0) norm[pos0 + j] = s;
        mx = fmaxf(mx, s);
    }
    if (lane == 0 && mx > 0.0f) atomicMax(nxmax, __float_as_uint(mx));
}

// ---------------------------------------------------------------- grouping of (query, probe) pairs by list
struct GroupParams {
    const uint64_t* probe_ids;  // [nq][np]
    const uint32_t* seg_len;    // [C]
    const uint32_t* used;       // optional [nq]: only slots s < used[q] are active (reference spill mode)
    const uint32_t* qmask;      // optional [nq]: only queries with a non-zero mask are active (exact fallback pass)
    uint32_t nq, np, C;
    uint32_t tb;          // queries per work-item group (8 for the SIMT scans, 32 for the tensor-core scan)
    uint32_t chunk_rows, chunk_rows_tail, tail_list0;  // rows per work item: lists >= tail_list0 use chunk_rows_tail
    uint32_t* qtau;       // optional [nq][4]: per-query shared bound (4 slots) of the tensor-core scan, reset here
    const uint32_t* skip_if_zero;  // optional: the whole pass is a no-op when this device word is 0 (no query to redo)
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
    if (g.skip_if_zero && *g.skip_if_zero == 0) return;
    uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= g.nq * g.np) return;
    uint32_t q = pi / g.np, s = pi % g.np;
    if (g.qtau && pi < g.nq)  // TAU_INF (ivf_tc.cuh): +inf in the ordered encoding
        reinterpret_cast<uint4*>(g.qtau)[pi] = make_uint4(0xff800000u, 0xff800000u, 0xff800000u, 0xff800000u);
    uint32_t nch = 0;
    if ((!g.used || s < g.used[q]) && (!g.qmask || g.qmask[q] != 0)) {
        uint32_t l = (uint32_t)g.probe_ids[pi];
        uint32_t len = g.seg_len[l];
        uint32_t cr = l >= g.tail_list0 ? g.chunk_rows_tail : g.chunk_rows;
        nch = (len + cr - 1) / cr;
        if (nch) atomicAdd(&g.lq_cnt[l], 1u);
    }
    g.pair_nch[pi] = nch;
}

__global__ void group_items_kernel(GroupParams g) {
    if (g.skip_if_zero && *g.skip_if_zero == 0) return;
    uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = l < g.C;
    uint32_t m = in ? g.lq_cnt[l] : 0u;
    uint32_t cr = l >= g.tail_list0 ? g.chunk_rows_tail : g.chunk_rows;
    uint32_t len = in ? g.seg_len[l] : 0u;
    uint32_t nch = (len + cr - 1) / cr;
    uint32_t items = ((m + g.tb - 1) / g.tb) * nch;
    if (in) g.item_cnt[l] = items;
    if (g.stats) {  // warp-aggregated: one atomic per counter per warp
        unsigned long long v0 = m ? len : 0ull, v1 = (unsigned long long)len * m, v2 = m ? items : 0ull, v3 = m ? 1ull : 0ull;
        for (int o = 16; o; o >>= 1) {
            v0 += __shfl_xor_sync(FULL_MASK, v0, o);
            v1 += __shfl_xor_sync(FULL_MASK, v1, o);
            v2 += __shfl_xor_sync(FULL_MASK, v2, o);
            v3 += __shfl_xor_sync(FULL_MASK, v3, o);
        }
        if ((threadIdx.x & 31) == 0 && v3) {
            atomicAdd(&g.stats[0], v0);
            atomicAdd(&g.stats[1], v1);
            atomicAdd(&g.stats[2], v2);
            atomicAdd(&g.stats[3], v3);
        }
    }
}

__global__ void group_fill_kernel(GroupParams g) {
    if (g.skip_if_zero && *g.skip_if_zero == 0) return;
    uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= g.nq * g.np) return;
    if (g.pair_nch[pi] == 0) return;
    uint32_t l = (uint32_t)g.probe_ids[pi];
    uint32_t slot = atomicAdd(&g.cursor[l], 1u);
    uint64_t at = g.lq_off[l] + slot;
    g.lq_query[at] = pi / g.np;
    g.lq_pair[at] = pi;
}

// The whole grouping (the memset and the six launches above) as ONE block when the per-list counters fit shared
// memory.  Used for the exact redo pass of the candidate path: that pass almost always has no query to redo, and one
// launch that returns at once replaces seven.  Same tables, same (arbitrary) order of the queries inside a list.
constexpr uint32_t GROUP_FUSED_MAX_C = 8192, GROUP_FUSED_MAX_PAIRS = 1u << 19;

// exclusive prefix of x over the block's 1024 threads (thread order) and the block total; two barriers
__device__ __forceinline__ uint64_t group_block_scan(uint64_t x, uint64_t* warp_tot, uint64_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = x;
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(FULL_MASK, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();  // warp_tot may still be read from the previous call
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint64_t w = warp_tot[lane];  // every warp scans the 32 warp totals itself
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(FULL_MASK, w, o);
        if (lane >= o) w += y;
    }
    total = __shfl_sync(FULL_MASK, w, 31);
    const uint64_t before = __shfl_sync(FULL_MASK, w, warp ? warp - 1 : 0);
    return (warp ? before : 0ull) + inc - x;
}

__global__ void __launch_bounds__(1024) group_fused_kernel(GroupParams g, uint64_t* __restrict__ lq_off,
                                                           uint64_t* __restrict__ item_off,
                                                           uint64_t* __restrict__ pair_chunk_off,
                                                           unsigned long long* __restrict__ work_counter) {
    if (g.skip_if_zero && *g.skip_if_zero == 0) return;
    extern __shared__ uint32_t gf_sm[];
    uint32_t* s_cnt = gf_sm;          // [C] queries per list, then the fill cursor (counts down)
    uint32_t* s_off = gf_sm + g.C;    // [C] start of the list's group in lq_query / lq_pair
    __shared__ uint64_t warp_tot[32];
    const uint32_t tid = threadIdx.x, npairs = g.nq * g.np;
    for (uint32_t l = tid; l < g.C; l += 1024) s_cnt[l] = 0;
    if (g.qtau)
        for (uint32_t q = tid; q < 4 * g.nq; q += 1024) g.qtau[q] = 0xff800000u;  // TAU_INF (ivf_tc.cuh)
    if (tid == 0) *work_counter = 0ull;
    __syncthreads();
    // 1. chunks of every active pair's list, queries per list (unrolled: the loads of 8 pairs are in flight together)
#pragma unroll 8
    for (uint32_t pi = tid; pi < npairs; pi += 1024) {
        const uint32_t q = pi / g.np, sl = pi % g.np;
        uint32_t nch = 0;
        if ((!g.used || sl < g.used[q]) && (!g.qmask || g.qmask[q] != 0)) {
            const uint32_t l = (uint32_t)g.probe_ids[pi];
            const uint32_t len = g.seg_len[l];
            const uint32_t cr = l >= g.tail_list0 ? g.chunk_rows_tail : g.chunk_rows;
            nch = (len + cr - 1) / cr;
            if (nch) atomicAdd(&s_cnt[l], 1u);
        }
        g.pair_nch[pi] = nch;
    }
    __syncthreads();
    // 2. per list: work items; exclusive scans of both over the lists (8 consecutive lists per thread)
    {
        constexpr int IT = GROUP_FUSED_MAX_C / 1024;
        uint32_t m[IT], items[IT];
        uint64_t msum = 0, isum = 0;
        unsigned long long v0 = 0, v1 = 0, v3 = 0;
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            const uint32_t l = tid * IT + k;
            m[k] = 0, items[k] = 0;
            if (l < g.C) {
                m[k] = s_cnt[l];
                const uint32_t len = g.seg_len[l];
                const uint32_t cr = l >= g.tail_list0 ? g.chunk_rows_tail : g.chunk_rows;
                items[k] = ((m[k] + g.tb - 1) / g.tb) * ((len + cr - 1) / cr);
                if (m[k]) v0 += len, v1 += (unsigned long long)len * m[k], v3 += 1;
            }
            msum += m[k];
            isum += items[k];
        }
        uint64_t mtot, itot;
        uint64_t mrun = group_block_scan(msum, warp_tot, mtot);
        uint64_t irun = group_block_scan(isum, warp_tot, itot);
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            const uint32_t l = tid * IT + k;
            if (l < g.C) {
                lq_off[l] = mrun;
                item_off[l] = irun;
                s_off[l] = (uint32_t)mrun;
            }
            mrun += m[k];
            irun += items[k];
        }
        if (tid == 0) {
            lq_off[g.C] = mtot;
            item_off[g.C] = itot;
        }
        if (g.stats) {  // rows of the distinct lists, (row, query) pairs, work items, lists touched
            for (int o = 16; o; o >>= 1) {
                v0 += __shfl_xor_sync(FULL_MASK, v0, o);
                v1 += __shfl_xor_sync(FULL_MASK, v1, o);
                v3 += __shfl_xor_sync(FULL_MASK, v3, o);
            }
            if ((tid & 31) == 0 && v3) {
                atomicAdd(&g.stats[0], v0);
                atomicAdd(&g.stats[1], v1);
                atomicAdd(&g.stats[3], v3);
            }
            if (tid == 0) atomicAdd(&g.stats[2], (unsigned long long)itot);
        }
    }
    // 3. exclusive scan of the pairs' chunk counts (tiles of 8192 with a running carry)
    {
        uint64_t carry = 0;
        for (uint32_t base = 0; base < npairs; base += 8192) {
            const uint32_t i0 = base + tid * 8;
            uint32_t v[8];
            uint64_t tsum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = i0 + k < npairs ? g.pair_nch[i0 + k] : 0u;
                tsum += v[k];
            }
            uint64_t tot;
            uint64_t run = carry + group_block_scan(tsum, warp_tot, tot);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (i0 + k < npairs) pair_chunk_off[i0 + k] = run;
                run += v[k];
            }
            carry += tot;
        }
        if (tid == 0) pair_chunk_off[npairs] = carry;
    }
    __syncthreads();  // s_off complete (and every read of s_cnt as a count is over)
    // 4. the pairs grouped by list
#pragma unroll 8
    for (uint32_t pi = tid; pi < npairs; pi += 1024) {
        if (g.pair_nch[pi] == 0) continue;
        const uint32_t l = (uint32_t)g.probe_ids[pi];
        const uint32_t slot = atomicSub(&s_cnt[l], 1u) - 1u;
        const uint64_t at = (uint64_t)s_off[l] + slot;
        g.lq_query[at] = pi / g.np;
        g.lq_pair[at] = pi;
    }
}

// ---------------------------------------------------------------- list scan (the dominant kernel)
struct ListScanParams {
    uint32_t chunk_rows;  // rows per work item (multiple of Cfg::TA): LIST_CHUNK_ROWS, or LIST_CHUNK_ROWS_SMALL when a
                          // handful of queries would otherwise leave one CTA to walk a whole list
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
    const float* lm_norm;  // MODE 1 only
    const uint32_t* skip_if_zero;  // optional: no-op launch when this device word is 0
};

// MODE 0: exact order (OP_L2SQ), private top-k by (distance, position)
// MODE 1: candidate pass (FMA dot, key = ||x||^2 - 2 x.q), private top-M by (key, position); p.k == p.kpad == M
template <class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::NT, (Cfg::TILE_FLOATS * 4 > 110 * 1024) ? 1 : 2) list_scan_kernel(ListScanParams p) {
    extern __shared__ __align__(16) float smem[];
    float* list_d = smem + Cfg::TILE_FLOATS;
    uint32_t* list_p = reinterpret_cast<uint32_t*>(list_d + Cfg::NLISTS * p.kpad);
    __shared__ long long s_item;
    __shared__ uint32_t s_list;
    const uint64_t total_items = p.item_off[p.C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (p.skip_if_zero && *p.skip_if_zero == 0) return;
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
        const uint32_t nch = (len + p.chunk_rows - 1) / p.chunk_rows;
        const uint32_t group = (uint32_t)(local / nch), chunk = (uint32_t)(local % nch);
        const uint64_t q0 = p.lq_off[l] + (uint64_t)group * Cfg::TB;
        const uint64_t m_l = p.lq_off[l + 1] - p.lq_off[l];
        const uint64_t nB = min((uint64_t)Cfg::TB, m_l - (uint64_t)group * Cfg::TB);
        const uint64_t base_pos = p.seg_off[l];
        RowSrc A{p.lm + base_pos * p.ld, nullptr, p.ld, len};
        RowSrc B{p.queries, p.lq_query + q0, p.ld, nB};
        const uint64_t r0 = (uint64_t)chunk * p.chunk_rows;
        const uint64_t r1 = min((uint64_t)len, r0 + p.chunk_rows);
        if (MODE == 0) {
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
        } else {
            // candidate pass: every warp keeps, per query column, a private top-32 by (key, position) IN REGISTERS
            // (lane l holds the l-th smallest); an insert is one ballot + one shuffle-up.  Requires the layout of
            // StreamCfg: all lanes of a warp share tb (= 0) and own all MB columns.
            static_assert(MODE == 0 || (Cfg::TBS_PER_WARP == 1 && Cfg::NTB == 1), "register lists need StreamCfg");
            float rl_d[Cfg::MB];
            uint32_t rl_p[Cfg::MB];
            float tau_d[Cfg::MB];
            uint32_t tau_p[Cfg::MB];
#pragma unroll
            for (int j = 0; j < Cfg::MB; ++j) {
                rl_d[j] = tau_d[j] = __int_as_float(0x7f800000);
                rl_p[j] = tau_p[j] = 0xffffffffu;
            }
            const int ta = threadIdx.x % Cfg::NTA;
            for (uint64_t a0 = r0; a0 < r1; a0 += Cfg::TA) {
                float acc[Cfg::MA][Cfg::MB];
                tile_compute<Cfg, OP_DOT_FMA>(acc, A, a0, B, 0, p.ld, smem);
#pragma unroll
                for (int i = 0; i < Cfg::MA; ++i) {
                    const uint64_t row = a0 + (uint64_t)(ta + i * Cfg::NTA);
                    const bool rowlive = row < r1;
                    const uint32_t pos = (uint32_t)(row + base_pos);
                    const float nx = rowlive ? __ldg(p.lm_norm + pos) : 0.0f;
#pragma unroll
                    for (int j = 0; j < Cfg::MB; ++j) {
                        const float v = __fmaf_rn(-2.0f, acc[i][j], nx);  // ||x||^2 - 2 x.q
                        bool live = rowlive && (uint64_t)j < nB;
                        while (true) {
                            bool pass = live && entry_less<uint32_t>(v, pos, tau_d[j], tau_p[j]);
                            unsigned m = __ballot_sync(FULL_MASK, pass);
                            if (!m) break;
                            int src = __ffs(m) - 1;
                            float cv = __shfl_sync(FULL_MASK, v, src);
                            uint32_t cp = __shfl_sync(FULL_MASK, pos, src);
                            int ins = __popc(__ballot_sync(FULL_MASK, entry_less<uint32_t>(rl_d[j], rl_p[j], cv, cp)));
                            float ud = __shfl_up_sync(FULL_MASK, rl_d[j], 1);
                            uint32_t up = __shfl_up_sync(FULL_MASK, rl_p[j], 1);
                            if (lane > ins) {
                                rl_d[j] = ud;
                                rl_p[j] = up;
                            } else if (lane == ins) {
                                rl_d[j] = cv;
                                rl_p[j] = cp;
                            }
                            tau_d[j] = __shfl_sync(FULL_MASK, rl_d[j], 31);
                            tau_p[j] = __shfl_sync(FULL_MASK, rl_p[j], 31);
                            if (lane == src) live = false;
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < Cfg::MB; ++j) {
                if ((uint64_t)j < nB) {
                    uint32_t pair = p.lq_pair[q0 + j];
                    uint64_t base = ((p.pair_chunk_off[pair] + chunk) * Cfg::NSPLIT + warp) * 32;
                    p.part_d[base + lane] = rl_d[j];
                    p.part_p[base + lane] = rl_p[j];
                }
            }
        }
        __syncthreads();  // s_item / lists are rewritten by the next iteration
    }
}

}  // namespace vers
#include "ivf_tc.cuh"
#include "flat_tc.cuh"
namespace vers {

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
    VERS_CUDA(cudaMalloc(&ivf->d_stats, 128));
    VERS_CUDA(cudaMalloc(&ivf->d_lm_norm, n1 * 4));
    VERS_CUDA(cudaMalloc(&ivf->d_nxmax, 4));
    VERS_CUDA(cudaMemsetAsync(ivf->d_nxmax, 0, 4, ctx->stream));
    VERS_CUDA(cudaMalloc(&ivf->d_cent_norm, (size_t)km->C * 4));
    VERS_CUDA(cudaMalloc(&ivf->d_ncmax, 4));
    VERS_CUDA(cudaMemsetAsync(ivf->d_ncmax, 0, 4, ctx->stream));
    ivf->cap_total = ds->n;
    {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        VERS_TRY(kmeans_build_csr(km));
        VERS_CUDA(cudaMemcpyAsync(ivf->d_cents, km->d_cents, (size_t)km->C * ds->ld * 4, cudaMemcpyDeviceToDevice,
                                  ctx->stream));
        rownorm_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(ivf->d_cents, ds->ld, 0, km->C, ivf->d_cent_norm,
                                                                  ivf->d_ncmax);
        VERS_LAUNCH_CHECK(ctx);
        if (ds->n) {
            VERS_CUDA(cudaMemcpyAsync(ivf->d_assign, km->d_assign, ds->n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            gather_list_major_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(
                ds->d_rows, ds->ld, km->d_sorted_rows, ds->n, ds->id_base, ivf->d_lm, ivf->d_lm_ids);
            VERS_LAUNCH_CHECK(ctx);
            rownorm_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ivf->d_lm, ds->ld, 0, ds->n, ivf->d_lm_norm,
                                                                      ivf->d_nxmax);
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
            ivf->seg_epoch += 1;
        }
    }
    return ivf_upload_segments(ivf);
}

// give every list slack capacity and move the rows (rare: only when an add hits a full list)
// (extra: optional per-list number of rows about to be appended; the slack is computed on top of it)
static int32_t ivf_relayout(vers_ivf* ivf, const uint32_t* extra = nullptr) {
    vers_ctx* ctx = ivf->ctx;
    std::vector<uint64_t> noff(ivf->C);
    std::vector<uint32_t> ncap(ivf->C);
    uint64_t total = 0;
    for (uint32_t c = 0; c < ivf->C; ++c) {
        uint32_t len = ivf->seg_len[c];
        const uint32_t want = len + (extra ? extra[c] : 0u);
        uint32_t cap = want + std::max<uint32_t>(32u, want / 8);
        noff[c] = total;
        ncap[c] = cap;
        total += cap;
    }
    if (total >= 0xffffffffull) return fail(VERS_ERR_UNSUPPORTED, "ivf: more than 2^32-2 list slots per GPU shard");
    float* nlm = nullptr;
    uint64_t* nids = nullptr;
    float* nnorm = nullptr;
    VERS_CUDA(cudaMalloc(&nlm, (size_t)total * ivf->ld * 4));
    cudaError_t e = cudaMalloc(&nids, (size_t)total * 8);
    if (e == cudaSuccess) e = cudaMalloc(&nnorm, (size_t)total * 4);
    if (e != cudaSuccess) {
        cudaFree(nlm);
        cudaFree(nids);
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
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(nnorm + noff[c], ivf->d_lm_norm + ivf->seg_off[c], (size_t)len * 4,
                                cudaMemcpyDeviceToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(nlm);
        cudaFree(nids);
        cudaFree(nnorm);
        return fail(VERS_ERR_CUDA, "ivf relayout: %s", cudaGetErrorString(e));
    }
    cudaFree(ivf->d_lm);
    cudaFree(ivf->d_lm_ids);
    cudaFree(ivf->d_lm_norm);
    ivf->d_lm = nlm;
    ivf->d_lm_ids = nids;
    ivf->d_lm_norm = nnorm;
    ivf->cap_total = total;
    ivf->seg_off = noff;
    ivf->seg_cap = ncap;
    ivf->h16_valid = false;  // positions moved: the 16-bit copy is rebuilt by the next search that wants it
    return ivf_upload_segments(ivf);
}

// (re)builds the fp16 candidate copy when it is missing or stale.  One pass over the live rows (10M x 768: ~8 ms).
// Caller holds ctx->mu; never called under stream capture with a stale copy (the first eager search builds it).
static int32_t ivf_ensure_h16(vers_ivf* ivf) {
    if (ivf->h16_valid) return VERS_OK;
    vers_ctx* ctx = ivf->ctx;
    ivf->ld16 = round_up(ivf->dim, 8);
    const uint64_t cap = std::max<uint64_t>(ivf->cap_total, 1);
    if (!ivf->d_h16_stat) VERS_CUDA(cudaMalloc(&ivf->d_h16_stat, 8));
    if (!ivf->d_lm16 || ivf->h16_cap < cap) {
        VERS_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ivf->d_lm16);
        ivf->d_lm16 = nullptr;
        ivf->h16_cap = 0;
        VERS_CUDA(cudaMalloc(&ivf->d_lm16, (size_t)cap * ivf->ld16 * 2));
        ivf->h16_cap = cap;
    }
    uint32_t nxbits = 0;
    VERS_CUDA(cudaMemcpyAsync(&nxbits, ivf->d_nxmax, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    float nxmax;
    memcpy(&nxmax, &nxbits, 4);
    // |x_i| <= ||x|| < 2^ex  =>  |x_i| * 2^(14 - ex) < 2^14: a row may grow to 3x the current largest norm (later adds)
    // before an element could pass fp16's 65504
    int ex = 0;
    if (nxmax > 0.0f && std::isfinite(nxmax)) (void)std::frexp(std::sqrt((double)nxmax) * 1.0001, &ex);
    int e2 = std::max(-100, std::min(100, 14 - ex));
    ivf->h16_scale = std::ldexp(1.0f, e2);
    const double lim = 65504.0 / (double)ivf->h16_scale * 0.999;
    ivf->h16_norm2_limit = lim * lim;
    VERS_CUDA(cudaMemsetAsync(ivf->d_h16_stat, 0, 8, ctx->stream));
    // pad columns of the copy are written by the kernel (zeros), gap rows between the lists are never read as results
    rows_to_h16_kernel<<<std::min<uint32_t>(ivf->C, (uint32_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
        ivf->d_lm, ivf->ld, ivf->ld16, ivf->d_seg_off, ivf->d_seg_len, 0, ivf->C, 0, ivf->h16_scale, 1.0f / ivf->h16_scale,
        ivf->d_lm16, ivf->d_h16_stat, ivf->d_h16_stat + 1);
    VERS_LAUNCH_CHECK(ctx);
    ivf->h16_valid = true;
    return VERS_OK;
}

// upper bound on the number of (pair, chunk) partial lists one query can produce when it opens np lists
static uint64_t ivf_max_chunks_per_query_uncached(const vers_ivf* ivf, uint32_t np, uint32_t chunk_rows) {
    std::vector<uint32_t> nch(ivf->C);
    for (uint32_t c = 0; c < ivf->C; ++c) nch[c] = (ivf->seg_len[c] + chunk_rows - 1) / chunk_rows;
    np = std::min(np, ivf->C);
    std::partial_sort(nch.begin(), nch.begin() + np, nch.end(), std::greater<uint32_t>());
    uint64_t s = 0;
    for (uint32_t i = 0; i < np; ++i) s += nch[i];
    return s;
}

// [0]: whole-list items (LIST_CHUNK_ROWS), [1]: every list cut into TC_TAIL_CHUNK_ROWS items (upper bound of the
// tensor-core scan's mixed chunking)
static void ivf_max_chunks(vers_ivf* ivf, uint32_t np, uint64_t out[3]) {
    if (ivf->mc_epoch != ivf->seg_epoch || ivf->mc_np != np) {
        ivf->mc_val[0] = ivf_max_chunks_per_query_uncached(ivf, np, LIST_CHUNK_ROWS);
        ivf->mc_val[1] = ivf_max_chunks_per_query_uncached(ivf, np, TC_TAIL_CHUNK_ROWS);
        ivf->mc_val[2] = ivf_max_chunks_per_query_uncached(ivf, np, LIST_CHUNK_ROWS_SMALL);
        ivf->mc_epoch = ivf->seg_epoch;
        ivf->mc_np = np;
    }
    out[0] = ivf->mc_val[0];
    out[1] = ivf->mc_val[1];
    out[2] = ivf->mc_val[2];
}

// ---------------------------------------------------------------- candidate pass: merge, exact rerank, certificate
// Error model.  u = 2^-24, n = ld.
//   candidate value  d~ = ||x||^2 + ||q||^2 - 2 x.q with FMA / tree sums:  |d~ - d_true| <= E(x,q),
//                    E = 1.01 * (2n + 8) * u * (||x||^2 + ||q||^2)
//   reference value  d_ref (left-to-right fp32, base.rs:119-126):          d_ref >= d_true * (1 - (n + 3) u)
// A row that is NOT exactly re-ranked has d~ >= bound, hence d_ref >= (bound - Emax) * (1 - (n+3)u) =: lower.
// If lower > (exact distance of the k-th re-ranked row) no such row can enter or tie the top-k: certified.
// Otherwise the query is flagged and redone by the exact-order scan.

// one warp per query: top-M by (key, position) over the query's partial lists; bound = the smallest key any
// non-selected row can have = min(last key of every FULL partial list, M-th merged key if the merged list is full).
// The partial lists are sorted runs of 32, so the fold is a chain of bitonic merge-splits across the lanes: the
// running top-M lives in R = M/32 sorted registers per lane; a run whose head does not beat the current M-th entry
// is skipped after one comparison (most runs: far lists, late chunks).
__device__ __forceinline__ bool cm_less(float d0, uint32_t p0, float d1, uint32_t p1) {
    return d0 < d1 || (d0 == d1 && p0 < p1);
}
__device__ __forceinline__ void cm_merge_asc(float& d, uint32_t& p, int lane) {  // bitonic sequence -> ascending
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const float od = __shfl_xor_sync(FULL_MASK, d, j);
        const uint32_t op = __shfl_xor_sync(FULL_MASK, p, j);
        const bool keep_min = (lane & j) == 0;
        const bool other_less = cm_less(od, op, d, p), self_less = cm_less(d, p, od, op);
        if (keep_min ? other_less : self_less) {
            d = od;
            p = op;
        }
    }
}

// one BLOCK (W warps) per query: each warp folds every W-th group of runs into its own top-M, then warp 0 folds the
// other warps' sorted results (handed over through shared memory) into the final list.  W = 4 for the inverted-list
// scan (~100 runs per query, 1000 queries), 16 when a few queries each own thousands of runs (a table scanned in
// hundreds of work items: the walk is a chain of dependent trips to the partial lists, so more warps, not more bytes)
template <int R, int W>
__global__ void __launch_bounds__(32 * W)
    cand_merge_kernel(const float* __restrict__ part_d, const uint32_t* __restrict__ part_p,
                      const uint64_t* __restrict__ pair_chunk_off, uint32_t nq, uint32_t np, uint32_t nsplit,
                      uint32_t* __restrict__ cand_pos, float* __restrict__ cand_key, float* __restrict__ cand_bound) {
    constexpr uint32_t M = 32 * R;
    constexpr int PF = 4;  // runs fetched together
    __shared__ float xd[W - 1][R][32];
    __shared__ uint32_t xp[W - 1][R][32];
    __shared__ float xfull[W - 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x;
    const float INF = __int_as_float(0x7f800000);
    float ad[R];
    uint32_t ap[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ad[r] = INF;
        ap[r] = 0xffffffffu;
    }
    // folds one sorted run (d, pp: lane i = i-th smallest, empties last) into the running top-M
    auto fold = [&](float d, uint32_t pp) {
        const float first_d = __shfl_sync(FULL_MASK, d, 0);
        const uint32_t first_p = __shfl_sync(FULL_MASK, pp, 0);
        const float tau_d = __shfl_sync(FULL_MASK, ad[R - 1], 31);
        const uint32_t tau_p = __shfl_sync(FULL_MASK, ap[R - 1], 31);
        if (first_p == 0xffffffffu || !cm_less(first_d, first_p, tau_d, tau_p)) return;  // warp-uniform
#pragma unroll
        for (int r = 0; r < R; ++r) {
            // incoming run ascending -> reversed; lane-wise min with the ascending register = lower half of the
            // union (bitonic), lane-wise max = upper half (bitonic), which cascades into the next register
            const float rd = __shfl_sync(FULL_MASK, d, 31 - lane);
            const uint32_t rp = __shfl_sync(FULL_MASK, pp, 31 - lane);
            const bool take = cm_less(rd, rp, ad[r], ap[r]);
            float lo_d = take ? rd : ad[r], hi_d = take ? ad[r] : rd;
            uint32_t lo_p = take ? rp : ap[r], hi_p = take ? ap[r] : rp;
            cm_merge_asc(lo_d, lo_p, lane);
            ad[r] = lo_d;
            ap[r] = lo_p;
            if (r + 1 < R) {
                cm_merge_asc(hi_d, hi_p, lane);
                d = hi_d;
                pp = hi_p;
            }
        }
    };
    const uint64_t beg = pair_chunk_off[(uint64_t)q * np] * nsplit * 32;  // partial lists hold 32 entries each
    const uint64_t end = pair_chunk_off[(uint64_t)(q + 1) * np] * nsplit * 32;
    float tfull = INF;
    // the next group of runs is requested before the current one is folded: the fold of mostly-skipped runs is far
    // shorter than a trip to the partial lists (46 % of this kernel's stall samples sat on the first use of a load)
    float fd[PF], nd[PF];
    uint32_t fp[PF], npp[PF];
    auto load = [&](uint64_t e0, float (&d)[PF], uint32_t (&pp)[PF]) {
#pragma unroll
        for (int f = 0; f < PF; ++f) {
            const uint64_t e = e0 + (uint64_t)f * 32 + lane;
            const bool in = e < end;
            d[f] = in ? __ldcg(part_d + e) : INF;
            pp[f] = in ? __ldcg(part_p + e) : 0xffffffffu;
        }
    };
    uint64_t e0 = beg + (uint64_t)warp * 32 * PF;
    if (e0 < end) load(e0, fd, fp);
    for (; e0 < end; e0 += W * 32 * PF) {
        const uint64_t e1 = e0 + W * 32 * PF;
        if (e1 < end) load(e1, nd, npp);
#pragma unroll
        for (int f = 0; f < PF; ++f) {
            const float last_d = __shfl_sync(FULL_MASK, fd[f], 31);
            const uint32_t last_p = __shfl_sync(FULL_MASK, fp[f], 31);
            if (last_p != 0xffffffffu) tfull = fminf(tfull, last_d);  // a full run: its dropped rows are >= last_d
            fold(fd[f], fp[f]);
        }
#pragma unroll
        for (int f = 0; f < PF; ++f) {
            fd[f] = nd[f];
            fp[f] = npp[f];
        }
    }
    if (warp > 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            xd[warp - 1][r][lane] = ad[r];
            xp[warp - 1][r][lane] = ap[r];
        }
        if (lane == 0) xfull[warp - 1] = tfull;
    }
    __syncthreads();
    if (warp > 0) return;
    for (int w = 0; w < W - 1; ++w) {
        tfull = fminf(tfull, xfull[w]);
#pragma unroll
        for (int r = 0; r < R; ++r) fold(xd[w][r][lane], xp[w][r][lane]);  // sorted runs of 32, ascending overall
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        cand_pos[(uint64_t)q * M + r * 32 + lane] = ap[r];
        cand_key[(uint64_t)q * M + r * 32 + lane] = ad[r];
    }
    const float mlast_d = __shfl_sync(FULL_MASK, ad[R - 1], 31);
    const uint32_t mlast_p = __shfl_sync(FULL_MASK, ap[R - 1], 31);
    if (lane == 0) {
        float b = tfull;
        if (mlast_p != 0xffffffffu) b = fminf(b, mlast_d);
        cand_bound[q] = b;
    }
}

static int32_t launch_cand_merge(vers_ctx* ctx, uint32_t M, const float* part_d, const uint32_t* part_p,
                                 const uint64_t* pair_chunk_off, uint32_t nq, uint32_t np, uint32_t nsplit,
                                 uint32_t* cand_pos, float* cand_key, float* cand_bound, uint64_t runs_per_query = 0) {
    const unsigned grid = nq;
    if (M == 32 && runs_per_query >= 256)
        cand_merge_kernel<1, 16><<<grid, 512, 0, ctx->stream>>>(part_d, part_p, pair_chunk_off, nq, np, nsplit, cand_pos,
                                                               cand_key, cand_bound);
    else if (M == 32)
        cand_merge_kernel<1, 4><<<grid, 128, 0, ctx->stream>>>(part_d, part_p, pair_chunk_off, nq, np, nsplit, cand_pos,
                                                              cand_key, cand_bound);
    else if (M == 64)
        cand_merge_kernel<2, 4><<<grid, 128, 0, ctx->stream>>>(part_d, part_p, pair_chunk_off, nq, np, nsplit, cand_pos,
                                                              cand_key, cand_bound);
    else if (M == 128)
        cand_merge_kernel<4, 4><<<grid, 128, 0, ctx->stream>>>(part_d, part_p, pair_chunk_off, nq, np, nsplit, cand_pos,
                                                              cand_key, cand_bound);
    else
        return fail(VERS_ERR_ARG, "cand_merge: unsupported candidate count %u", M);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

// Exact-order l2sq of the M = 32 R candidates of every query (lane = candidate, strictly sequential over the
// dimensions like base.rs:119-126), top-k by (distance, id), certificate.  R warps per query, 4 warps per block.
// The candidate rows are scattered over the index, so each warp stages them chunk by chunk: 16 coalesced 16-byte
// asynchronous copies per lane (cp.async; two candidates' 64-float chunks per instruction) into one of two padded
// shared-memory tiles, chunk c + 1 in flight while lane i walks row i of chunk c sequentially.
// Dimensions per staged chunk: 64 with one warp per query (72 KB of staging per block, the 250 blocks of a 1000-query
// batch are all resident); 32 when R warps share a query (37 KB: the 500+ blocks still fit one wave)
template <int R>
struct RrCfg {
    static constexpr int KCH = R == 1 ? 64 : 32;
    static constexpr int LDS = KCH + 4;  // padded tile row (floats): lanes walking their own rows stay conflict-free
    static constexpr size_t STAGE_BYTES = (size_t)(2 * 4 * 32 * LDS + 2 * 4 * KCH) * 4;
};
template <int R>
__global__ void __launch_bounds__(128)
    rerank_certify_kernel(const float* __restrict__ lm, const uint64_t* __restrict__ lm_ids, uint64_t id_base,
                          uint32_t ld, const float* __restrict__ queries, uint32_t nq, uint32_t k,
                          const uint32_t* __restrict__ cand_pos, const float* __restrict__ cand_bound,
                          const uint32_t* __restrict__ nxmax_bits, const float* __restrict__ cand_key, int tf32_pass,
                          uint64_t* out_ids, float* out_d, uint32_t* out_cnt, uint32_t* fail_flag,
                          unsigned long long* stats, uint32_t* fail_list, uint32_t* n_fail,
                          const uint32_t* __restrict__ h16_stat) {
    constexpr uint32_t M = 32 * R;
    constexpr int QPB = 4 / R;  // queries per block
    constexpr int RR_KCH = RrCfg<R>::KCH, RR_LDS = RrCfg<R>::LDS;
    constexpr int LPR = RR_KCH / 4, RPI = 32 / LPR;  // lanes per row chunk, rows per copy instruction
    __shared__ float sdist[4][32], skey[4][32];
    __shared__ uint64_t sid[4][32];
    // dynamic: two staging buffers (row tiles [2][4 warps][32][RR_LDS], query chunks [2][4][RR_KCH]), then per query of
    // the block k ids (u64) and k distances
    extern __shared__ __align__(16) unsigned char rr_dyn[];
    float* tile_base = reinterpret_cast<float*>(rr_dyn);
    float* qs_base = tile_base + 2 * 4 * 32 * RR_LDS;
    unsigned char* rsm2 = reinterpret_cast<unsigned char*>(qs_base + 2 * 4 * RR_KCH);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wq = warp / R, wr = warp % R;
    const uint32_t q = blockIdx.x * QPB + wq;
    const bool active = q < nq;
    const uint32_t pos = active ? cand_pos[(uint64_t)q * M + wr * 32 + lane] : 0xffffffffu;
    const bool live = pos != 0xffffffffu;
    const float* qrow = queries + (uint64_t)(active ? q : 0) * ld;
    const int half = lane / LPR, sub = lane % LPR;
    // chunk c of the warp's 32 candidate rows and of its query -> staging buffer c & 1, asynchronously (cp.async, 16 B
    // per lane and copy: RPI candidates' chunks per instruction): the copies of chunk c + 1 are in flight
    // while chunk c is consumed, so the scattered-row latency is paid once, not once per chunk
    auto issue = [&](uint32_t c) {
        const uint32_t k0 = c * RR_KCH, col = k0 + sub * 4;
        float* tb = tile_base + ((size_t)(c & 1u) * 4 + warp) * 32 * RR_LDS;
#pragma unroll
        for (int i = 0; i < 32 / RPI; ++i) {
            const uint32_t cpos = __shfl_sync(FULL_MASK, pos, RPI * i + half);
            const bool ok = cpos != 0xffffffffu && col < ld;
            cp_async16(tb + (RPI * i + half) * RR_LDS + sub * 4, ok ? (const void*)(lm + (uint64_t)cpos * ld + col) : (const void*)lm, ok);
        }
        if (lane < LPR) {
            const uint32_t qc = k0 + lane * 4;
            const bool ok = active && qc < ld;
            cp_async16(qs_base + ((size_t)(c & 1u) * 4 + warp) * RR_KCH + lane * 4, ok ? (const void*)(qrow + qc) : (const void*)queries, ok);
        }
        cp_async_commit();
    };
    float s = 0.0f, nq2 = 0.0f, qres2 = 0.0f;  // qres2: ||q - (q_hi + q_lo)||^2 of the fp16 pass's query split
    const uint32_t nch = (ld + RR_KCH - 1) / RR_KCH;
    issue(0);
    for (uint32_t c = 0; c < nch; ++c) {
        const uint32_t k0 = c * RR_KCH, kn = min((uint32_t)RR_KCH, ld - k0);  // multiple of 4
        if (c + 1 < nch) {
            issue(c + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();  // every lane's copies of chunk c have landed
        const float* tl = tile_base + (((size_t)(c & 1u) * 4 + warp) * 32 + lane) * RR_LDS;
        const float* qc = qs_base + ((size_t)(c & 1u) * 4 + warp) * RR_KCH;
#pragma unroll
        for (int e = 0; e < RR_KCH / 32; ++e) {
            const float v = qc[lane + 32 * e];  // zero past ld
            nq2 = __fmaf_rn(v, v, nq2);
            if (tf32_pass == 4) {
                __half h, l;
                h16_split(v, h, l);
                const float r = __fsub_rn(__fsub_rn(v, __half2float(h)), __fmul_rn(__half2float(l), H16_LO_INV));
                qres2 = __fmaf_rn(r, r, qres2);
            }
        }
        if (live) {
#pragma unroll 4
            for (uint32_t i = 0; i < kn; i += 4) {
                const float4 a = *reinterpret_cast<const float4*>(tl + i);
                const float4 b = *reinterpret_cast<const float4*>(qc + i);
                float t;
                t = __fsub_rn(a.x, b.x); s = __fadd_rn(s, __fmul_rn(t, t));
                t = __fsub_rn(a.y, b.y); s = __fadd_rn(s, __fmul_rn(t, t));
                t = __fsub_rn(a.z, b.z); s = __fadd_rn(s, __fmul_rn(t, t));
                t = __fsub_rn(a.w, b.w); s = __fadd_rn(s, __fmul_rn(t, t));
            }
        }
        __syncwarp();  // the buffer is rewritten by the copies of chunk c + 2
    }
    for (int o = 16; o; o >>= 1) {
        nq2 += __shfl_xor_sync(FULL_MASK, nq2, o);
        qres2 += __shfl_xor_sync(FULL_MASK, qres2, o);
    }
    sdist[warp][lane] = s;
    skey[warp][lane] = live ? cand_key[(uint64_t)q * M + wr * 32 + lane] : 0.f;
    sid[warp][lane] = live ? (lm_ids ? lm_ids[pos] : id_base + pos) : 0xffffffffffffffffull;
    __syncthreads();
    if (!active || wr != 0) return;

    // ---- one warp per query: top-k by (distance, id) over its M exact distances, certificate
    uint64_t* sp = reinterpret_cast<uint64_t*>(rsm2) + (size_t)wq * k;
    float* sd = reinterpret_cast<float*>(rsm2 + (size_t)QPB * k * 8) + (size_t)wq * k;
    for (uint32_t e = lane; e < k; e += 32) {
        sd[e] = __int_as_float(0x7f800000);
        sp[e] = 0xffffffffffffffffull;
    }
    __syncwarp();
    uint32_t reranked = 0;
    float err = 0.0f;
    for (int g = 0; g < R; ++g) {
        const float v = sdist[wq * R + g][lane];
        const uint64_t id = sid[wq * R + g][lane];
        const bool lv = id != 0xffffffffffffffffull;
        reranked += __popc(__ballot_sync(FULL_MASK, lv));
        // observed candidate-pass error |d~ - d_ref| (statistic only: validates the certificate's error allowance)
        if (lv) err = fmaxf(err, fabsf((skey[wq * R + g][lane] + nq2) - v));
        bool pend = lv;
        while (true) {
            bool pass = pend && entry_less<uint64_t>(v, id, sd[k - 1], sp[k - 1]);
            unsigned m = __ballot_sync(FULL_MASK, pass);
            if (!m) break;
            int src = __ffs(m) - 1;
            float bv = __shfl_sync(FULL_MASK, v, src);
            uint64_t bid = __shfl_sync(FULL_MASK, id, src);
            warp_topk_insert<uint64_t>(sd, sp, (int)k, bv, bid, lane);
            if (lane == src) pend = false;
        }
    }
    for (int o = 16; o; o >>= 1) err = fmaxf(err, __shfl_xor_sync(FULL_MASK, err, o));
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
    if (lane == 0) {
        if (out_cnt) out_cnt[q] = cnt;
        const float bound = cand_bound[q];  // key bound: d~ = key + ||q||^2
        bool certified;
        if (bound == __int_as_float(0x7f800000)) {
            certified = true;  // every row of the probed lists was re-ranked exactly
        } else if (cnt < k) {
            certified = false;
        } else {
            const double u = 5.9604644775390625e-08;  // 2^-24
            const double nxmax = (double)__uint_as_float(*nxmax_bits);
            // fp32 terms (norms, key arithmetic, and the FMA dot of the SIMT pass) + for the tensor-core pass the TF32
            // operand truncation (|x - tf32(x)| <= 2^-10 |x| per operand => 2^-9 (1 + 2^-11) per product, times
            // x.q <= (||x||^2 + ||q||^2)/2, times the factor 2 of the key) and an accumulation allowance of
            // (n + 8) 2^-21 per unit of sum |x_i q_i| (4x the bound of a correctly rounded fp32 chain)
            double E = 1.01 * (2.0 * ld + 8.0) * u * (nxmax + (double)nq2);
            if (tf32_pass == 1) E += (1.001 / 512.0 + (ld + 8.0) * 4.76837158203125e-07) * (nxmax + (double)nq2);
            // split precision (hi.hi + lo.hi + hi.lo): dropped lo.lo and the truncation of lo are <= 3 * 2^-20 per
            // product; accumulation allowance (n + 8) 2^-22 per unit of sum |x_i q_i|
            if (tf32_pass == 2) E += (3.003 / 1048576.0 + (ld + 8.0) * 2.384185791015625e-07) * (nxmax + (double)nq2);
            // both operands ROUNDED to nearest tf32 (tc_flat_kernel): 2^-11 per operand => 2^-10 (1 + 2^-12) per product
            if (tf32_pass == 3) E += (1.001 / 1024.0 + (ld + 8.0) * 4.76837158203125e-07) * (nxmax + (double)nq2);
            // fp16 candidate copy: key error 2 |x.q - x~.q~| <= 2 (||x - x~|| ||q|| + ||x~|| ||q - q~||) (Cauchy-Schwarz;
            // max ||x - x~||^2 over the index measured when the copy was written, ||q - q~||^2 computed above,
            // ||x~|| <= ||x|| + ||x - x~||), fp16 products are exact in the fp32 accumulator: allowance (n + 8) 2^-22 per
            // unit of sum |x_i q_i|.  A copy with elements that did not fit certifies nothing.
            bool h16_ok = true;
            if (tf32_pass == 4) {
                const double xlo2 = (double)__uint_as_float(h16_stat[0]);
                h16_ok = h16_stat[1] == 0u;
                E += 2.002 * (sqrt(xlo2 * (double)nq2) + (sqrt(nxmax) + sqrt(xlo2)) * sqrt((double)qres2)) +
                     (ld + 8.0) * 2.384185791015625e-07 * (nxmax + (double)nq2);
            }
            const double lower = ((double)bound + (double)nq2 - E) * (1.0 - (ld + 3.0) * u);
            certified = h16_ok && lower > (double)sd[k - 1];
        }
        fail_flag[q] = certified ? 0u : 1u;
        if (!certified && fail_list) fail_list[atomicAdd(n_fail, 1u)] = q;
        if (!certified) atomicAdd(&stats[4], 1ull);
        atomicAdd(&stats[5], (unsigned long long)reranked);
        atomicMax(&stats[6], (unsigned long long)__float_as_uint(err));  // non-negative floats order as uints
    }
}

template <int R>
static int32_t launch_rerank_r(vers_ctx* ctx, const float* lm, const uint64_t* lm_ids, uint64_t id_base, uint32_t ld,
                               const float* queries, uint32_t nq, uint32_t k, const uint32_t* cand_pos,
                               const float* cand_bound, const uint32_t* nxmax_bits, const float* cand_key, int tf32_pass,
                               uint64_t* out_ids, float* out_d, uint32_t* out_cnt, uint32_t* fail_flag,
                               unsigned long long* stats, uint32_t* fail_list, uint32_t* n_fail, const uint32_t* h16_stat) {
    static bool attr_set = false;
    if (!attr_set) {
        VERS_CUDA(cudaFuncSetAttribute(rerank_certify_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(RrCfg<R>::STAGE_BYTES + 4 * VERS_MAX_TOPK * 12)));
        attr_set = true;
    }
    const size_t smem = RrCfg<R>::STAGE_BYTES + (size_t)(4 / R) * k * 12;
    rerank_certify_kernel<R><<<(unsigned)ceil_div(nq, 4 / R), 128, smem, ctx->stream>>>(
        lm, lm_ids, id_base, ld, queries, nq, k, cand_pos, cand_bound, nxmax_bits, cand_key, tf32_pass, out_ids, out_d,
        out_cnt, fail_flag, stats, fail_list, n_fail, h16_stat);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

static int32_t launch_rerank(vers_ctx* ctx, uint32_t M, const float* lm, const uint64_t* lm_ids, uint64_t id_base,
                             uint32_t ld, const float* queries, uint32_t nq, uint32_t k, const uint32_t* cand_pos,
                             const float* cand_bound, const uint32_t* nxmax_bits, const float* cand_key, int tf32_pass,
                             uint64_t* out_ids, float* out_d, uint32_t* out_cnt, uint32_t* fail_flag,
                             unsigned long long* stats, uint32_t* fail_list, uint32_t* n_fail,
                             const uint32_t* h16_stat = nullptr) {
    if (k > M) return fail(VERS_ERR_ARG, "rerank: k %u > %u candidates", k, M);
#define VERS_RR_ARGS                                                                                                   \
    ctx, lm, lm_ids, id_base, ld, queries, nq, k, cand_pos, cand_bound, nxmax_bits, cand_key, tf32_pass, out_ids, out_d, \
        out_cnt, fail_flag, stats, fail_list, n_fail, h16_stat
    if (M == 32) return launch_rerank_r<1>(VERS_RR_ARGS);
    if (M == 64) return launch_rerank_r<2>(VERS_RR_ARGS);
    if (M == 128) return launch_rerank_r<4>(VERS_RR_ARGS);
#undef VERS_RR_ARGS
    return fail(VERS_ERR_ARG, "rerank: unsupported candidate count %u", M);
}

// ---------------------------------------------------------------- host: one batched search
struct SearchBufs {
    uint64_t* probe_ids;
    float* probe_d;
    uint32_t *lq_cnt, *cursor, *item_cnt, *pair_nch, *lq_query, *lq_pair, *used, *fail_flag, *cand_pos;
    uint64_t *lq_off, *item_off, *pair_chunk_off;
    unsigned long long* counter;  // [0] work counter, [1] short flag
    float* cand_bound;
    float* part_d;
    uint32_t* part_p;
    float* gq;     // [npairs + 32][ld] queries regrouped by list (tensor-core scan only), tf32 hi part
    float* gq_lo;  // [npairs + 32][ld] their tf32 lo part (split-precision scan)
    float* cand_key;  // [nq][M] candidate keys (observed-error statistic)
    uint32_t* qtau;   // [nq][4] shared per-query bound of the tensor-core scan
};

static int32_t run_group(vers_ivf* ivf, const SearchBufs& b, uint32_t nq, uint32_t np, const uint32_t* used,
                         const uint32_t* qmask, bool record_stats, uint32_t tb = ScanCfg::TB,
                         uint32_t chunk_rows = LIST_CHUNK_ROWS, uint32_t chunk_rows_tail = LIST_CHUNK_ROWS,
                         uint32_t tail_list0 = 0xffffffffu,
                         uint32_t* qtau = nullptr, const uint32_t* skip_if_zero = nullptr) {
    vers_ctx* ctx = ivf->ctx;
    const uint64_t npairs = (uint64_t)nq * np;
    // the fused single-block kernel only for the exact redo pass, which almost always has nothing to do: ONE launch
    // that returns at once instead of seven (measured with work to do: 85 us for 32000 pairs in one block against 31 us
    // for the multi-block path, so the main grouping keeps the latter)
    // ... and for a handful of queries (the reference's one-query call): a few hundred pairs, where one small launch
    // beats seven
    const bool fused = (skip_if_zero != nullptr || npairs <= 2048) && ivf->C <= GROUP_FUSED_MAX_C &&
                       npairs <= GROUP_FUSED_MAX_PAIRS;
    // lq_cnt, cursor and the work counter are carved back to back: one memset (counter[1], the short flag, survives)
    if (!fused) VERS_CUDA(cudaMemsetAsync(b.lq_cnt, 0, (size_t)((char*)b.counter - (char*)b.lq_cnt) + 8, ctx->stream));
    GroupParams g;
    g.probe_ids = b.probe_ids;
    g.seg_len = ivf->d_seg_len;
    g.used = used;
    g.qmask = qmask;
    g.nq = nq;
    g.np = np;
    g.C = ivf->C;
    g.tb = tb;
    g.chunk_rows = chunk_rows;
    g.chunk_rows_tail = chunk_rows_tail;
    g.tail_list0 = tail_list0;
    g.qtau = qtau;
    g.skip_if_zero = skip_if_zero;
    g.lq_cnt = b.lq_cnt;
    g.pair_nch = b.pair_nch;
    g.item_cnt = b.item_cnt;
    g.lq_off = b.lq_off;
    g.cursor = b.cursor;
    g.stats = record_stats ? ivf->d_stats : nullptr;
    g.lq_query = b.lq_query;
    g.lq_pair = b.lq_pair;
    if (fused) {
        static bool attr_set = false;
        if (!attr_set) {
            VERS_CUDA(cudaFuncSetAttribute(group_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(2 * GROUP_FUSED_MAX_C * 4)));
            attr_set = true;
        }
        group_fused_kernel<<<1, 1024, (size_t)2 * ivf->C * 4, ctx->stream>>>(g, b.lq_off, b.item_off, b.pair_chunk_off,
                                                                            b.counter);
        VERS_LAUNCH_CHECK(ctx);
        return VERS_OK;
    }
    group_count_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, ctx->stream>>>(g);
    VERS_LAUNCH_CHECK(ctx);
    group_items_kernel<<<(unsigned)ceil_div(ivf->C, 256), 256, 0, ctx->stream>>>(g);
    VERS_LAUNCH_CHECK(ctx);
    VERS_TRY(launch_exclusive_scan(ctx, b.lq_cnt, ivf->C, b.lq_off, skip_if_zero));
    VERS_TRY(launch_exclusive_scan(ctx, b.item_cnt, ivf->C, b.item_off, skip_if_zero));
    VERS_TRY(launch_exclusive_scan(ctx, b.pair_nch, npairs, b.pair_chunk_off, skip_if_zero));
    group_fill_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, ctx->stream>>>(g);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

template <class Cfg, int MODE>
static int32_t run_list_scan(vers_ivf* ivf, const SearchBufs& b, const float* d_queries, uint32_t nq, uint32_t klist,
                             const uint32_t* skip_if_zero = nullptr, uint32_t chunk_rows = LIST_CHUNK_ROWS) {
    vers_ctx* ctx = ivf->ctx;
    ListScanParams lp;
    lp.chunk_rows = chunk_rows;
    lp.lm = ivf->d_lm;
    lp.queries = d_queries;
    lp.ld = ivf->ld;
    lp.C = ivf->C;
    lp.k = klist;
    lp.kpad = round_up(klist, 32);
    lp.seg_off = ivf->d_seg_off;
    lp.seg_len = ivf->d_seg_len;
    lp.lq_query = b.lq_query;
    lp.lq_pair = b.lq_pair;
    lp.lq_off = b.lq_off;
    lp.item_off = b.item_off;
    lp.pair_chunk_off = b.pair_chunk_off;
    lp.nq = nq;
    lp.part_d = b.part_d;
    lp.part_p = b.part_p;
    lp.counter = b.counter;
    lp.lm_norm = ivf->d_lm_norm;
    lp.skip_if_zero = skip_if_zero;
    auto kern = list_scan_kernel<Cfg, MODE>;
    size_t smem = scan_smem_bytes(Cfg::TILE_FLOATS, Cfg::NLISTS, lp.kpad);
    VERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FamilyTimer ft(ctx, MODE == 0 ? KF_LIST_SCAN : KF_CAND_SCAN);
    kern<<<ctx->sm_count * ((Cfg::TILE_FLOATS * 4 > 110 * 1024) ? 1 : 2), Cfg::NT, smem, ctx->stream>>>(lp);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

// launches the tensor-core scan over `rows` ([n_rows][ld], norms in tp.lm_norm) with the grouped queries gq / gq_lo
// ([gq_rows][ld]); tp carries the work-item tables.  PREC 2: rows / gq / gq_lo are fp16 tables of ld halfs per row.
template <int PREC>
static int32_t launch_tc_scan(vers_ctx* ctx, const void* rows, uint64_t n_rows, uint32_t ld, const void* gq,
                              const void* gq_lo, uint64_t gq_rows, const TcScanParams& tp, int family) {
    using Cfg = TcCfg<PREC>;
    CUtensorMap tm_rows, tm_rows32, tm_q16, tm_ql16, tm_q32, tm_ql32;
    const void* lo_src = (Cfg::SPLIT3 || Cfg::H16) ? gq_lo : gq;
    const uint64_t nr = n_rows ? n_rows : 1;
    if (Cfg::H16) {
        VERS_TRY(make_tmap_2d_f16(&tm_rows, rows, nr, ld, ld, TC_M, Cfg::KC_ELEMS));
        VERS_TRY(make_tmap_2d_f16(&tm_rows32, rows, nr, ld, ld, 32, Cfg::KC_ELEMS));
        VERS_TRY(make_tmap_2d_f16(&tm_q16, gq, gq_rows, ld, ld, 16, Cfg::KC_ELEMS));
        VERS_TRY(make_tmap_2d_f16(&tm_ql16, lo_src, gq_rows, ld, ld, 16, Cfg::KC_ELEMS));
        VERS_TRY(make_tmap_2d_f16(&tm_q32, gq, gq_rows, ld, ld, 32, Cfg::KC_ELEMS));
        VERS_TRY(make_tmap_2d_f16(&tm_ql32, lo_src, gq_rows, ld, ld, 32, Cfg::KC_ELEMS));
    } else {
        const float* frows = static_cast<const float*>(rows);
        const float* fq = static_cast<const float*>(gq);
        const float* fl = static_cast<const float*>(lo_src);
        VERS_TRY(make_tmap_2d_f32(&tm_rows, frows, nr, ld, ld, TC_M, TC_KC));
        VERS_TRY(make_tmap_2d_f32(&tm_rows32, frows, nr, ld, ld, 32, TC_KC));
        VERS_TRY(make_tmap_2d_f32(&tm_q16, fq, gq_rows, ld, ld, 16, TC_KC));
        VERS_TRY(make_tmap_2d_f32(&tm_ql16, fl, gq_rows, ld, ld, 16, TC_KC));
        VERS_TRY(make_tmap_2d_f32(&tm_q32, fq, gq_rows, ld, ld, 32, TC_KC));
        VERS_TRY(make_tmap_2d_f32(&tm_ql32, fl, gq_rows, ld, ld, 32, TC_KC));
    }
    auto kern = tc_list_scan_kernel<PREC>;
    VERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    FamilyTimer ft(ctx, family);
    kern<<<ctx->sm_count, Cfg::THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(tm_rows, tm_rows32, tm_q16, tm_ql16, tm_q32,
                                                                        tm_ql32, tp);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

template <int PREC>
static int32_t run_list_scan_tc(vers_ivf* ivf, const SearchBufs& b, const float* d_queries, uint32_t nq, uint32_t np,
                                uint32_t chunk_rows) {
    vers_ctx* ctx = ivf->ctx;
    constexpr bool H16 = PREC == 2;
    const uint64_t npairs = (uint64_t)nq * np;
    if (H16)
        gather_queries_h16_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(
            d_queries, b.lq_query, b.lq_off, ivf->C, ivf->ld, ivf->ld16, reinterpret_cast<__half*>(b.gq),
            reinterpret_cast<__half*>(b.gq_lo));
    else
        gather_queries_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_queries, b.lq_query, b.lq_off, ivf->C, ivf->ld,
                                                                         b.gq, b.gq_lo, PREC == 1 ? 1 : 0);
    VERS_LAUNCH_CHECK(ctx);
    TcScanParams tp;
    tp.ld = H16 ? ivf->ld16 : ivf->ld;
    tp.C = ivf->C;
    tp.chunk_rows = chunk_rows;
    tp.chunk_rows_tail = TC_TAIL_CHUNK_ROWS;
    tp.tail_list0 = tc_tail_list0(ivf->C);
    tp.seg_off = ivf->d_seg_off;
    tp.seg_len = ivf->d_seg_len;
    tp.lq_pair = b.lq_pair;
    tp.lq_off = b.lq_off;
    tp.item_off = b.item_off;
    tp.pair_chunk_off = b.pair_chunk_off;
    tp.lm_norm = ivf->d_lm_norm;
    tp.part_d = b.part_d;
    tp.part_p = b.part_p;
    tp.counter = b.counter;
    tp.qtau = b.qtau;
    tp.lq_query = b.lq_query;
    tp.dense_out = nullptr;
    tp.dense_ld = 0;
    tp.key_scale = H16 ? -2.0f / ivf->h16_scale : -2.0f;
    const void* rows = H16 ? static_cast<const void*>(ivf->d_lm16) : static_cast<const void*>(ivf->d_lm);
    return launch_tc_scan<PREC>(ctx, rows, ivf->cap_total, tp.ld, b.gq, b.gq_lo, npairs + TC_NQ, tp, KF_CAND_SCAN);
}

// ---------------------------------------------------------------- centroid probe
// nprobe nearest lists of every query by (distance, centroid index), exact order (ivfflat.rs:155-161).
//   exact   : the exact-order tile engine over all C centroids (scan_topk_run), fp32-pipe bound
//   tensor  : (mode 0, batches of >= 32 queries) the centroid table is scanned like one inverted list by the
//             tensor-core candidate kernel -> top-M keys per query (M >= 2 nprobe) -> exact-order distances of the M
//             candidates -> the same rounding-error certificate as the list scan proves that no other centroid can
//             enter or tie the top-nprobe; queries that fail it are appended to a list on the device and redone by
//             the exact engine in the same call (blocks of that launch exit at once when the list is empty).
constexpr uint32_t PROBE_DENSE_MAX_C = 8192;  // C keys of one query fit the select kernel's shared memory
// the table the queries are ranked against: the centroids (probe) or a whole dataset (exhaustive search)
struct RankTable {
    const float* rows;          // [n][ld]
    uint64_t n;
    uint32_t ld;
    const float* norm;          // [n] ||row||^2, any order
    const uint32_t* nmax_bits;  // max ||row||^2
    uint64_t id_base;           // reported id = id_base + row
    bool allow_tc;
    unsigned long long* stats;  // stats[4] uncertified queries, [5] re-ranked rows, [6] max observed candidate error
};
struct ProbePlan {
    ScanPlan exact;
    bool stream = false;  // <= 8 queries against a small table: streaming dense distances + radix select + sort
    bool tc = false;
    bool dense = false;  // tc && C small: all C keys per query are written out and selected by probe_select_kernel
    uint32_t nch = 0, chunk_rows = 0, M = 0;
    size_t bytes = 0;
};

struct ProbeBufs {
    uint64_t *seg_off, *lq_off, *item_off, *pair_chunk_off;
    uint32_t *seg_len, *lq_pair, *cand_pos, *fail, *fail_idx, *n_fail, *part_p, *qtau;
    unsigned long long *counter, *pstats;
    float *gq, *gq_lo, *part_d, *cand_key, *bound, *tmp_d, *dense;
    uint64_t* tmp_ids;
};

static void probe_carve(ScratchCarver& sc, const RankTable& tb, const ProbePlan& pp, uint32_t nq, uint32_t np,
                        ProbeBufs& b) {
    b.seg_off = sc.take<uint64_t>(1);
    b.seg_len = sc.take<uint32_t>(1);
    b.lq_off = sc.take<uint64_t>(2);
    b.item_off = sc.take<uint64_t>(2);
    b.pair_chunk_off = sc.take<uint64_t>((size_t)nq + 1);
    b.lq_pair = sc.take<uint32_t>(nq);
    b.counter = sc.take<unsigned long long>(2);
    b.pstats = sc.take<unsigned long long>(8);
    b.n_fail = sc.take<uint32_t>(1);
    b.gq = sc.take<float>((size_t)(nq + TC_NQ) * tb.ld);
    b.gq_lo = sc.take<float>((size_t)(nq + TC_NQ) * tb.ld);
    b.part_d = sc.take<float>(pp.dense ? 4 : (size_t)nq * pp.nch * TC_PARTS * 32);
    b.part_p = sc.take<uint32_t>(pp.dense ? 4 : (size_t)nq * pp.nch * TC_PARTS * 32);
    b.dense = sc.take<float>(pp.dense ? (size_t)nq * tb.n : 4);
    b.cand_pos = sc.take<uint32_t>((size_t)nq * pp.M);
    b.cand_key = sc.take<float>((size_t)nq * pp.M);
    b.bound = sc.take<float>(nq);
    b.fail = sc.take<uint32_t>(nq);
    b.fail_idx = sc.take<uint32_t>(nq);
    b.tmp_ids = sc.take<uint64_t>((size_t)nq * np);
    b.tmp_d = sc.take<float>((size_t)nq * np);
    b.qtau = sc.take<uint32_t>((size_t)4 * nq);
}

static ProbePlan probe_plan(const vers_ctx* ctx, const RankTable& tb, uint32_t nq, uint32_t np) {
    ProbePlan pp;
    pp.exact = scan_topk_plan(ctx, tb.n, nq, np);
    pp.bytes = (pp.exact.bytes + 255) & ~size_t(255);
    pp.tc = tb.allow_tc && nq >= tc_min_batch() && np <= 64 && tb.n >= 512 && tb.n >= 4ull * np && tb.ld >= TC_KC &&
            tb.n < 0x7fffffffull;
    // the reference's one-query call (and any batch of <= 8): the tile engine leaves one CTA per 256 centroids to walk
    // them and one warp to merge thousands of entries (0.6 ms of a 1.1 ms call); the streaming kernel writes all C exact
    // distances in ~10 us, a radix select + a rank sort pick the np nearest in (distance, index) order
    pp.stream = !pp.tc && nq <= SMALL_BATCH && np <= 128 && tb.n >= np && tb.n <= PROBE_DENSE_MAX_C &&
                flat_stream_fits(tb.ld, nq);
    if (pp.stream) {
        ScratchCarver sc(nullptr);
        sc.plan<float>((size_t)8 * tb.ld);
        sc.plan<float>((size_t)8 * tb.n);
        sc.plan<uint32_t>((size_t)nq * np);
        sc.plan<float>((size_t)nq * np);
        sc.plan<float>(nq);
        pp.bytes = std::max(pp.bytes, (sc.off + 255) & ~size_t(255));
    }
    if (pp.tc) {
        pp.dense = tb.n <= PROBE_DENSE_MAX_C;
        // large tables: 32 candidates certify up to 16 requested entries (the inverted-list scan's rule), and 32 is what
        // one epilogue list holds: only then can the work items of a query share one running bound.  The dense probe
        // selects from all keys anyway and keeps 64: k-means leaves tied (zero-vector) centroids behind
        // (ivfflat.rs:63-67), and 8 requested lists must not fail their certificate on 30 of those
        pp.M = (!pp.dense && np <= 16) ? 32 : np <= 32 ? 64 : 128;
        const uint32_t ngroups = (nq + TC_NQ - 1) / TC_NQ;
        // small tables (the centroids): one wave of work items; large ones (a dataset): ~2 items per SM (every item
        // costs a pipeline fill and, per query, 4 partial lists the merge has to walk: 1M x 300, 32 queries with 8 items
        // per SM: scan 0.51 ms + merge 0.20 ms), and a chunk must keep its row offsets in 16 bits
        const uint32_t target = tb.n <= 65536 ? (uint32_t)ctx->sm_count : (uint32_t)ctx->sm_count * 2;
        uint64_t nch = std::max<uint32_t>(1, target / ngroups);
        nch = std::min<uint64_t>(nch, (tb.n + TC_M - 1) / TC_M);
        uint64_t cr = round_up((uint32_t)((tb.n + nch - 1) / nch), (uint32_t)TC_M);
        cr = std::min<uint64_t>(cr, 32768);
        pp.chunk_rows = (uint32_t)cr;
        pp.nch = (uint32_t)((tb.n + cr - 1) / cr);
        ScratchCarver sc(nullptr);
        sc.off = pp.bytes;
        ProbeBufs b;
        probe_carve(sc, tb, pp, nq, np, b);
        pp.bytes = (sc.off + 255) & ~size_t(255);
    }
    return pp;
}

static RankTable centroid_table(const vers_ivf* ivf) {
    RankTable tb;
    tb.rows = ivf->d_cents;
    tb.n = ivf->C;
    tb.ld = ivf->ld;
    tb.norm = ivf->d_cent_norm;
    tb.nmax_bits = ivf->d_ncmax;
    tb.id_base = 0;
    tb.allow_tc = ivf->mode == 0 || ivf->mode == 4;  // the probe itself always runs split-precision tf32
    tb.stats = ivf->d_stats + 3;  // stats[7] = uncertified probe queries, [8] = re-ranked centroids
    return tb;
}

// ---- dense probe: exact selection of the M smallest (key, centroid) pairs of one query out of its C keys.
// One block per query, keys staged in shared memory in the order-preserving uint32 encoding; 4 radix passes (8 bits
// each, most significant first) find the M-th smallest value T; everything below T is a candidate, ties at T are
// taken in centroid order (the (key, position) order of the other paths).  bound = T: no other centroid has a
// smaller key.  Candidates come out unsorted: the exact rerank orders them.
__global__ void __launch_bounds__(256)
    probe_select_kernel(const float* __restrict__ dense, uint32_t C, uint32_t M, uint32_t* __restrict__ cand_pos,
                        float* __restrict__ cand_key, float* __restrict__ cand_bound) {
    extern __shared__ uint32_t ps_keys[];  // [C]
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_need, s_cnt;
    const uint32_t q = blockIdx.x, tid = threadIdx.x;
    const float* row = dense + (uint64_t)q * C;
    for (uint32_t i = tid; i < C; i += 256) ps_keys[i] = tau_encode(row[i]);
    if (tid == 0) {
        s_prefix = 0;
        s_need = M;
        s_cnt = 0;
    }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        for (uint32_t i = tid; i < C; i += 256) {
            const uint32_t u = ps_keys[i];
            if (pass == 0 || (u >> (shift + 8)) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {  // 256-bin exclusive scan by one warp (8 bins per lane), find the bin holding the need-th element
            uint32_t h[8], sum = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                h[e] = hist[tid * 8 + e];
                sum += h[e];
            }
            uint32_t incl = sum;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(FULL_MASK, incl, o);
                if ((int)tid >= o) incl += y;
            }
            uint32_t before = incl - sum;
            const uint32_t need = s_need;
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (need > before && need <= before + h[e]) {  // exactly one (lane, e) matches
                    s_prefix = (prefix << 8) | (uint32_t)(tid * 8 + e);
                    s_need = need - before;
                }
                before += h[e];
            }
        }
        __syncthreads();
    }
    const uint32_t T = s_prefix, need_eq = s_need;  // take every u < T and the first need_eq entries with u == T
    const uint32_t n_less = M - need_eq;
    for (uint32_t i = tid; i < C; i += 256) {
        const uint32_t u = ps_keys[i];
        if (u < T) {
            const uint32_t slot = atomicAdd(&s_cnt, 1u);
            cand_pos[(uint64_t)q * M + slot] = i;
            cand_key[(uint64_t)q * M + slot] = tau_decode(u);
        }
    }
    if (tid < 32) {  // ties at T in centroid order
        uint32_t taken = 0;
        for (uint32_t i0 = 0; i0 < C && taken < need_eq; i0 += 32) {
            const uint32_t i = i0 + tid;
            const bool eq = i < C && ps_keys[i] == T;
            const unsigned m = __ballot_sync(FULL_MASK, eq);
            const uint32_t rank = taken + __popc(m & ((1u << tid) - 1u));
            if (eq && rank < need_eq) {
                cand_pos[(uint64_t)q * M + n_less + rank] = i;
                cand_key[(uint64_t)q * M + n_less + rank] = tau_decode(T);
            }
            taken += __popc(m);
        }
        if (tid == 0) cand_bound[q] = tau_decode(T);
    }
}

// work-item tables of the probe: one "list" (the centroid table) probed by every query, nch chunks
__global__ void probe_tables_kernel(uint32_t C, uint32_t nq, uint32_t nch, uint64_t* seg_off, uint32_t* seg_len,
                                    uint64_t* lq_off, uint64_t* item_off, uint64_t* pair_chunk_off, uint32_t* lq_pair,
                                    uint32_t* qtau) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        seg_off[0] = 0;
        seg_len[0] = C;
        lq_off[0] = 0;
        lq_off[1] = nq;
        item_off[0] = 0;
        item_off[1] = (uint64_t)((nq + TC_NQ - 1) / TC_NQ) * nch;
    }
    if (i <= nq) pair_chunk_off[i] = (uint64_t)i * nch;
    if (i < nq) {
        lq_pair[i] = i;
        reinterpret_cast<uint4*>(qtau)[i] = make_uint4(TAU_INF, TAU_INF, TAU_INF, TAU_INF);
    }
}

__global__ void probe_scatter_kernel(const uint32_t* __restrict__ fail_idx, const uint32_t* __restrict__ n_fail,
                                     const uint64_t* __restrict__ tmp_ids, const float* __restrict__ tmp_d, uint32_t np,
                                     uint64_t* __restrict__ out_ids, float* __restrict__ out_d) {
    const uint32_t n = *n_fail;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n * np; i += gridDim.x * blockDim.x) {
        const uint32_t r = i / np, e = i - r * np;
        const uint64_t at = (uint64_t)fail_idx[r] * np + e;
        out_ids[at] = tmp_ids[i];
        out_d[at] = tmp_d[i];
    }
}

// scratch: [0, pp.bytes) of the context arena.  out_ids / out_d: [nq][np]
// the M selected (position, key) pairs of every query in (key, position) order -> ids and distances
__global__ void __launch_bounds__(128) probe_sort_kernel(const uint32_t* __restrict__ cand_pos,
                                                        const float* __restrict__ cand_key, uint32_t M, uint64_t id_base,
                                                        uint64_t* __restrict__ out_ids, float* __restrict__ out_d,
                                                        uint32_t* __restrict__ out_cnt) {
    __shared__ float sk[128];
    __shared__ uint32_t spos[128];
    const uint32_t q = blockIdx.x, t = threadIdx.x;
    if (t < M) {
        sk[t] = cand_key[(uint64_t)q * M + t];
        spos[t] = cand_pos[(uint64_t)q * M + t];
    }
    __syncthreads();
    if (t < M) {
        const float kt = sk[t];
        const uint32_t pt = spos[t];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < M; ++j) rank += (sk[j] < kt || (sk[j] == kt && spos[j] < pt)) ? 1u : 0u;
        out_ids[(uint64_t)q * M + rank] = id_base + pt;
        out_d[(uint64_t)q * M + rank] = kt;
    }
    if (t == 0 && out_cnt) out_cnt[q] = M;
}

static int32_t probe_run(vers_ctx* ctx, const RankTable& tb, const ProbePlan& pp, const float* d_queries, uint32_t nq,
                         uint32_t np, uint64_t* out_ids, float* out_d, uint32_t* out_cnt, int family) {
    RowSrc CA{tb.rows, nullptr, tb.ld, tb.n};
    if (pp.stream) {
        ScratchCarver sc(ctx->scratch);
        float* qpad = sc.take<float>((size_t)8 * tb.ld);
        float* dense = sc.take<float>((size_t)8 * tb.n);
        uint32_t* cpos = sc.take<uint32_t>((size_t)nq * np);
        float* ckey = sc.take<float>((size_t)nq * np);
        float* bound = sc.take<float>(nq);
        const int32_t rc = flat_stream_dense(ctx, tb.rows, tb.n, tb.ld, d_queries, nq, qpad, dense, family);
        if (rc == VERS_OK) {
            probe_select_kernel<<<nq, 256, (size_t)tb.n * 4, ctx->stream>>>(dense, (uint32_t)tb.n, np, cpos, ckey, bound);
            VERS_LAUNCH_CHECK(ctx);
            probe_sort_kernel<<<nq, 128, 0, ctx->stream>>>(cpos, ckey, np, tb.id_base, out_ids, out_d, out_cnt);
            VERS_LAUNCH_CHECK(ctx);
            return VERS_OK;
        }
        if (rc != VERS_ERR_UNSUPPORTED) return rc;
    }
    if (!pp.tc) {
        RowSrc QB{d_queries, nullptr, tb.ld, nq};
        return scan_topk_run(ctx, pp.exact, ctx->scratch, CA, QB, nq, tb.ld, np, VERS_METRIC_L2SQ, nullptr, tb.id_base,
                             out_ids, out_d, out_cnt, family);
    }
    ScratchCarver sc(ctx->scratch);
    sc.off = (pp.exact.bytes + 255) & ~size_t(255);
    ProbeBufs b;
    probe_carve(sc, tb, pp, nq, np, b);
    FamilyTimer ft(ctx, family);
    VERS_CUDA(cudaMemsetAsync(b.counter, 0, (size_t)((char*)b.n_fail - (char*)b.counter) + 4, ctx->stream));  // + pstats
    probe_tables_kernel<<<(unsigned)ceil_div(nq + 1, 256), 256, 0, ctx->stream>>>(
        (uint32_t)tb.n, nq, pp.nch, b.seg_off, b.seg_len, b.lq_off, b.item_off, b.pair_chunk_off, b.lq_pair, b.qtau);
    VERS_LAUNCH_CHECK(ctx);
    gather_queries_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_queries, nullptr, b.lq_off, 1, tb.ld, b.gq,
                                                                     b.gq_lo, 1);
    VERS_LAUNCH_CHECK(ctx);
    TcScanParams tp;
    tp.ld = tb.ld;
    tp.C = 1;
    tp.chunk_rows = pp.chunk_rows;
    tp.chunk_rows_tail = pp.chunk_rows;
    tp.tail_list0 = 0xffffffffu;
    tp.seg_off = b.seg_off;
    tp.seg_len = b.seg_len;
    tp.lq_pair = b.lq_pair;
    tp.lq_off = b.lq_off;
    tp.item_off = b.item_off;
    tp.pair_chunk_off = b.pair_chunk_off;
    tp.lm_norm = tb.norm;
    tp.part_d = b.part_d;
    tp.part_p = b.part_p;
    tp.counter = b.counter;
    // 32 candidates: every work item of a query prunes with the query's running 32nd-best key (without it each of the
    // ~8 items per SM restarts from +inf and the selection, not the stream, bounds the kernel: 1M x 300, 32 queries
    // 1.32 ms -> see profiles/README.md).  More than 32 candidates are merged from 32-entry lists: no shared bound.
    tp.qtau = (pp.M == 32 && !pp.dense) ? b.qtau : nullptr;
    tp.lq_query = b.lq_pair;  // grouped pair i is query i
    tp.dense_out = pp.dense ? b.dense : nullptr;
    tp.dense_ld = tb.n;
    tp.key_scale = -2.0f;
    VERS_TRY(launch_tc_scan<1>(ctx, tb.rows, tb.n, tb.ld, b.gq, b.gq_lo, (uint64_t)nq + TC_NQ, tp, -1));
    if (pp.dense) {
        probe_select_kernel<<<nq, 256, (size_t)tb.n * 4, ctx->stream>>>(b.dense, (uint32_t)tb.n, pp.M, b.cand_pos,
                                                                       b.cand_key, b.bound);
        VERS_LAUNCH_CHECK(ctx);
    } else {
        VERS_TRY(launch_cand_merge(ctx, pp.M, b.part_d, b.part_p, b.pair_chunk_off, nq, 1, TC_PARTS, b.cand_pos,
                                   b.cand_key, b.bound, (uint64_t)pp.nch * TC_PARTS));
    }
    VERS_TRY(launch_rerank(ctx, pp.M, tb.rows, nullptr, tb.id_base, tb.ld, d_queries, nq, np, b.cand_pos, b.bound,
                           tb.nmax_bits, b.cand_key, 2, out_ids, out_d, out_cnt, b.fail, tb.stats, b.fail_idx, b.n_fail));
    // exact redo of the uncertified queries (none, typically): the launch is sized for nq, the blocks read n_fail
    RowSrc QF{d_queries, b.fail_idx, tb.ld, nq};
    VERS_TRY(scan_topk_run(ctx, pp.exact, ctx->scratch, CA, QF, nq, tb.ld, np, VERS_METRIC_L2SQ, nullptr, tb.id_base,
                           b.tmp_ids, b.tmp_d, nullptr, -1, b.n_fail));
    probe_scatter_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(b.fail_idx, b.n_fail, b.tmp_ids, b.tmp_d, np, out_ids,
                                                               out_d);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

// ---- exhaustive search of a large batch, table streamed once per 128 queries (flat_tc.cuh): tc_flat_kernel ->
// cand_merge (top-32 by key) -> exact-order rerank + certificate (TF32 error model) -> exact redo of uncertified queries
struct QblockPlan {
    bool ok = false;
    ScanPlan exact;
    uint32_t nk = 0, nqb = 0, nslices = 0, slice_rows = 0, stages = 0;
    size_t smem = 0, bytes = 0;
};
struct QblockBufs {
    uint32_t *qtau, *part_p, *cand_pos, *fail, *fail_idx, *n_fail;
    float *part_d, *cand_key, *bound, *tmp_d;
    uint64_t *pair_chunk_off, *tmp_ids;
};
static void qblock_carve(ScratchCarver& sc, const QblockPlan& qp, uint32_t nq, uint32_t k, QblockBufs& b) {
    b.qtau = sc.take<uint32_t>((size_t)4 * nq);
    b.n_fail = sc.take<uint32_t>(1);
    b.pair_chunk_off = sc.take<uint64_t>((size_t)nq + 1);
    b.part_d = sc.take<float>((size_t)nq * qp.nslices * FT_LIST);
    b.part_p = sc.take<uint32_t>((size_t)nq * qp.nslices * FT_LIST);
    b.cand_pos = sc.take<uint32_t>((size_t)nq * 32);
    b.cand_key = sc.take<float>((size_t)nq * 32);
    b.bound = sc.take<float>(nq);
    b.fail = sc.take<uint32_t>(nq);
    b.fail_idx = sc.take<uint32_t>(nq);
    b.tmp_ids = sc.take<uint64_t>((size_t)nq * k);
    b.tmp_d = sc.take<float>((size_t)nq * k);
}
// smallest batch that takes the query-block kernel (one pass of the table per 128 queries; the list-scan style kernel
// takes one per 32: 1M x 300, 32 queries 0.42 ms, 64: 0.68 ms, 95: 0.94 ms)
static uint32_t qblock_min_batch() {
    static const uint32_t v = [] {
        const char* e = getenv("VERS_QBLOCK_MIN_NQ");
        const int x = e ? atoi(e) : 96;
        return (uint32_t)(x < 1 ? 1 : x);
    }();
    return v;
}
static QblockPlan qblock_plan(const vers_ctx* ctx, uint64_t n, uint32_t ld, uint32_t nq, uint32_t k) {
    QblockPlan qp;
    qp.nk = (ld + FT_KC - 1) / FT_KC;
    // eligible: enough queries to fill 128-wide blocks, rows that fit tensor memory, a top-k the 32 candidates cover
    if (nq < qblock_min_batch() || qp.nk > FT_MAX_KCH || ld < FT_KC || k > 16 || n < 4096 || n >= 0x7fffffffull) return qp;
    qp.nqb = (nq + FT_M - 1) / FT_M;
    const uint32_t want_items = (uint32_t)ctx->sm_count * 4;
    uint64_t nsl = std::max<uint64_t>(1, (want_items + qp.nqb - 1) / qp.nqb);
    nsl = std::min<uint64_t>(nsl, (n + 4 * FT_N - 1) / (4 * FT_N));  // at least 4 tiles per slice
    qp.slice_rows = round_up((uint32_t)((n + nsl - 1) / nsl), (uint32_t)FT_N);
    qp.nslices = (uint32_t)((n + qp.slice_rows - 1) / qp.slice_rows);
    const size_t budget = 227 * 1024;
    uint32_t st = 4;
    while (st > 2 && ft_smem_bytes(qp.nk, st) > budget) --st;
    if (ft_smem_bytes(qp.nk, st) > budget) return qp;
    qp.stages = st;
    qp.smem = ft_smem_bytes(qp.nk, st);
    qp.exact = scan_topk_plan(ctx, n, nq, k);
    ScratchCarver sc(nullptr);
    sc.off = (qp.exact.bytes + 255) & ~size_t(255);
    QblockBufs b;
    qblock_carve(sc, qp, nq, k, b);
    qp.bytes = (sc.off + 255) & ~size_t(255);
    qp.ok = true;
    return qp;
}

__global__ void qblock_tables_kernel(uint32_t nq, uint32_t nslices, uint64_t* pair_chunk_off, uint32_t* qtau) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nq) pair_chunk_off[i] = (uint64_t)i * nslices;
    if (i < nq) reinterpret_cast<uint4*>(qtau)[i] = make_uint4(TAU_INF, TAU_INF, TAU_INF, TAU_INF);
}

static int32_t qblock_run(vers_ctx* ctx, const RankTable& tb, const QblockPlan& qp, const float* row_tiles,
                          const float* d_queries, uint32_t nq, uint32_t k, uint64_t* out_ids, float* out_d,
                          uint32_t* out_cnt, int family) {
    ScratchCarver sc(ctx->scratch);
    sc.off = (qp.exact.bytes + 255) & ~size_t(255);
    QblockBufs b;
    qblock_carve(sc, qp, nq, k, b);
    FamilyTimer ft(ctx, family);
    VERS_CUDA(cudaMemsetAsync(b.n_fail, 0, 4, ctx->stream));
    qblock_tables_kernel<<<(unsigned)ceil_div((uint64_t)nq + 1, 256), 256, 0, ctx->stream>>>(nq, qp.nslices, b.pair_chunk_off,
                                                                                           b.qtau);
    VERS_LAUNCH_CHECK(ctx);
    CUtensorMap tm_q;
    VERS_TRY(make_tmap_2d_f32(&tm_q, d_queries, nq, tb.ld, tb.ld, FT_M, FT_KC));
    TcFlatParams p;
    p.n_rows = tb.n;
    p.nq = nq;
    p.ld = tb.ld;
    p.nk = qp.nk;
    p.nqb = qp.nqb;
    p.nslices = qp.nslices;
    p.slice_rows = qp.slice_rows;
    p.stages = qp.stages;
    p.row_tiles = row_tiles;
    p.row_norm = tb.norm;
    p.part_d = b.part_d;
    p.part_p = b.part_p;
    p.qtau = b.qtau;
    VERS_CUDA(cudaFuncSetAttribute(tc_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qp.smem));
    const uint64_t nitems = (uint64_t)qp.nslices * qp.nqb;
    tc_flat_kernel<<<(unsigned)std::min<uint64_t>(nitems, ctx->sm_count), FT_THREADS, qp.smem, ctx->stream>>>(tm_q, p);
    VERS_LAUNCH_CHECK(ctx);
    VERS_TRY(launch_cand_merge(ctx, 32, b.part_d, b.part_p, b.pair_chunk_off, nq, 1, 1, b.cand_pos, b.cand_key, b.bound,
                               nq <= 256 ? qp.nslices : 0));  // many queries: the blocks are the parallelism
    VERS_TRY(launch_rerank(ctx, 32, tb.rows, nullptr, tb.id_base, tb.ld, d_queries, nq, k, b.cand_pos, b.bound,
                           tb.nmax_bits, b.cand_key, 3, out_ids, out_d, out_cnt, b.fail, tb.stats, b.fail_idx, b.n_fail));
    // exact redo of the uncertified queries: the launch is sized for nq, the blocks read n_fail
    RowSrc CA{tb.rows, nullptr, tb.ld, tb.n};
    RowSrc QF{d_queries, b.fail_idx, tb.ld, nq};
    VERS_TRY(scan_topk_run(ctx, qp.exact, ctx->scratch, CA, QF, nq, tb.ld, k, VERS_METRIC_L2SQ, nullptr, tb.id_base,
                           b.tmp_ids, b.tmp_d, nullptr, -1, b.n_fail));
    probe_scatter_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(b.fail_idx, b.n_fail, b.tmp_ids, b.tmp_d, k, out_ids, out_d);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

// exhaustive search (utils.rs:68-82) of a query batch through the same tensor-core candidate path: the dataset is the
// table.  Declared in scan.cuh, called by vers_flat_search_dev (caller holds ctx->mu).
int32_t flat_search_tc_plan_and_run(vers_ctx* ctx, const float* rows, uint64_t n, uint32_t ld, const float* norm,
                                    const uint32_t* nmax_bits, uint64_t id_base, unsigned long long* stats,
                                    const float* d_queries, uint32_t nq, uint32_t k, uint64_t* d_ids, float* d_d,
                                    uint32_t* d_cnt, bool* used_tc, int flat_path, float** row_tiles_io) {
    RankTable tb;
    tb.rows = rows;
    tb.n = n;
    tb.ld = ld;
    tb.norm = norm;
    tb.nmax_bits = nmax_bits;
    tb.id_base = id_base;
    tb.allow_tc = true;
    tb.stats = stats;
    if (flat_path != 2) {  // large batches: queries resident in tensor memory, the table streamed once per 128 queries
        const QblockPlan qp = qblock_plan(ctx, n, ld, nq, k);
        if (qp.ok && row_tiles_io) {
            if (used_tc) *used_tc = true;
            if (!*row_tiles_io) {  // the tile-major tf32 image of the table, built once per dataset
                float* img = nullptr;
                VERS_CUDA(cudaMalloc(&img, (size_t)ceil_div(n, FT_N) * qp.nk * FT_BOX_BYTES));
                tile_image_tf32_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(rows, n, ld, qp.nk, img);
                ctx->launches += 1;
                const cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) {
                    cudaFree(img);
                    return fail(VERS_ERR_CUDA, "tile image: %s", cudaGetErrorString(e));
                }
                *row_tiles_io = img;
            }
            VERS_TRY(scratch_reserve(ctx, qp.bytes));
            VERS_CUDA(cudaMemsetAsync(stats, 0, 64, ctx->stream));
            return qblock_run(ctx, tb, qp, *row_tiles_io, d_queries, nq, k, d_ids, d_d, d_cnt, KF_FLAT_SCAN);
        }
    }
    const ProbePlan pp = probe_plan(ctx, tb, nq, k);
    if (used_tc) *used_tc = pp.tc;
    if (!pp.tc) return VERS_ERR_UNSUPPORTED;  // the caller runs the exact-order engine
    VERS_TRY(scratch_reserve(ctx, pp.bytes));
    VERS_CUDA(cudaMemsetAsync(stats, 0, 64, ctx->stream));
    return probe_run(ctx, tb, pp, d_queries, nq, k, d_ids, d_d, d_cnt, KF_FLAT_SCAN);
}

// ||row||^2 of n rows + running max (exported for flat.cu)
int32_t launch_rownorm(vers_ctx* ctx, const float* rows, uint32_t ld, uint64_t n, float* norm, uint32_t* nmax_bits) {
    VERS_CUDA(cudaMemsetAsync(nmax_bits, 0, 4, ctx->stream));
    rownorm_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(rows, ld, 0, n, norm, nmax_bits);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

static int32_t ivf_search_dev_locked(vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t k, uint32_t nprobe,
                                     uint64_t* d_ids, float* d_d, uint32_t* d_cnt,
                                     const uint64_t* ext_probe = nullptr) {
    vers_ctx* ctx = ivf->ctx;
    const bool ref_mode = nprobe == 0;
    const uint32_t np = ref_mode ? std::min<uint32_t>(ivf->C, VERS_MAX_TOPK) : std::min<uint32_t>(nprobe, ivf->C);
    if (np > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "nprobe %u > %u", np, VERS_MAX_TOPK);
    // candidate list length of the approximate pass: the private lists are one register per lane
    const uint32_t M = 32;
    const bool approx = !ref_mode && ivf->mode != 1 && k <= 16;
    const bool use_tc = approx && (ivf->mode == 0 || ivf->mode == 3 || ivf->mode == 4) && ivf->ld >= TC_KC &&
                        ivf->cap_total < 0x7fffffffull;
    const bool split3 = use_tc && ivf->mode == 0;
    const bool h16 = use_tc && ivf->mode == 4;  // 16-bit candidate copy
    if (h16) VERS_TRY(ivf_ensure_h16(ivf));
    const uint64_t npairs = (uint64_t)nq * np;
    uint64_t mc[3];
    ivf_max_chunks(ivf, np, mc);
    const uint32_t exact_chunk = nq <= SMALL_BATCH ? LIST_CHUNK_ROWS_SMALL : LIST_CHUNK_ROWS;  // exact-order scans
    const uint64_t max_chunks = std::max<uint64_t>((uint64_t)nq * (nq <= SMALL_BATCH ? mc[2] : mc[0]), 1);
    size_t entries = (size_t)max_chunks * ScanCfg::NSPLIT * k;
    if (approx) entries = std::max(entries, (size_t)std::max<uint64_t>((uint64_t)nq * mc[0], 1) * StreamCfg::NSPLIT * M);
    if (use_tc) entries = std::max(entries, (size_t)nq * std::max<uint64_t>(mc[1], 1) * TC_PARTS * M);

    // the probe carves its buffers from the front of the arena, ours come after it
    const RankTable ctab = centroid_table(ivf);
    const ProbePlan pplan = probe_plan(ctx, ctab, nq, np);
    const size_t probe_reserve = pplan.bytes;
    SearchBufs b;
    auto carve = [&](ScratchCarver& sc) {
        sc.off = probe_reserve;
        b.probe_ids = sc.take<uint64_t>(npairs);
        b.probe_d = sc.take<float>(npairs);
        b.lq_cnt = sc.take<uint32_t>(ivf->C);
        b.cursor = sc.take<uint32_t>(ivf->C);
        b.counter = sc.take<unsigned long long>(2);
        b.item_cnt = sc.take<uint32_t>(ivf->C);
        b.pair_nch = sc.take<uint32_t>(npairs);
        b.lq_off = sc.take<uint64_t>((size_t)ivf->C + 1);
        b.item_off = sc.take<uint64_t>((size_t)ivf->C + 1);
        b.pair_chunk_off = sc.take<uint64_t>(npairs + 1);
        b.lq_query = sc.take<uint32_t>(npairs);
        b.lq_pair = sc.take<uint32_t>(npairs);
        b.used = sc.take<uint32_t>(nq);
        b.fail_flag = sc.take<uint32_t>(nq);
        b.cand_pos = sc.take<uint32_t>((size_t)nq * M);
        b.cand_bound = sc.take<float>(nq);
        b.part_d = sc.take<float>(entries);
        b.part_p = sc.take<uint32_t>(entries);
        b.gq = sc.take<float>(use_tc ? (size_t)(npairs + TC_NQ) * ivf->ld : 4);
        b.gq_lo = sc.take<float>((split3 || h16) ? (size_t)(npairs + TC_NQ) * ivf->ld : 4);
        b.cand_key = sc.take<float>((size_t)nq * M);
        b.qtau = sc.take<uint32_t>((size_t)4 * nq);
    };
    {
        ScratchCarver plan(nullptr);
        carve(plan);
        VERS_TRY(scratch_reserve(ctx, plan.off + 4096));
    }
    ScratchCarver sc(ctx->scratch);
    carve(sc);

    VERS_CUDA(cudaMemsetAsync(ivf->d_stats, 0, 128, ctx->stream));
    // 1. probe: exact-order distances to every centroid, top-np by (distance, centroid index)
    //    (or the caller's probe lists: the multi-GPU driver splits the probe of a batch over the ranks)
    if (ext_probe) {
        VERS_CUDA(cudaMemcpyAsync(b.probe_ids, ext_probe, (size_t)npairs * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        VERS_TRY(probe_run(ctx, ctab, pplan, d_queries, nq, np, b.probe_ids, b.probe_d, nullptr, KF_PROBE));
    }
    VERS_CUDA(cudaMemsetAsync(b.counter, 0, 16, ctx->stream));
    uint32_t* short_flag = reinterpret_cast<uint32_t*>(b.counter + 1);

    if (ref_mode) {
        // the reference's nearest-list-plus-spill semantics, exact order everywhere
        ref_plan_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, ctx->stream>>>(b.probe_ids, ivf->d_seg_len, nq, np, k,
                                                                             b.used, short_flag);
        VERS_LAUNCH_CHECK(ctx);
        VERS_TRY(run_group(ivf, b, nq, np, b.used, nullptr, true, ScanCfg::TB, exact_chunk, exact_chunk));
        VERS_TRY((run_list_scan<ScanCfg, 0>(ivf, b, d_queries, nq, k, nullptr, exact_chunk)));
        ref_assemble_kernel<<<(unsigned)ceil_div(nq, 4), 128, (size_t)4 * k * 8, ctx->stream>>>(
            b.part_d, b.part_p, b.pair_chunk_off, b.used, ivf->d_lm_ids, nq, np, k, ScanCfg::NSPLIT, d_ids, d_d, d_cnt);
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

    const uint32_t* qmask = nullptr;
    if (approx) {
        // 2a. candidate pass (FMA, HBM-streaming) -> top-M per query -> exact-order rerank -> certificate
        uint32_t nsplit;
        if (use_tc) {
            const uint32_t cr = tc_chunk_rows(ivf, nq, np);
            VERS_TRY(run_group(ivf, b, nq, np, nullptr, nullptr, true, TC_NQ, cr, TC_TAIL_CHUNK_ROWS,
                               tc_tail_list0(ivf->C), b.qtau));
            if (h16)
                VERS_TRY(run_list_scan_tc<2>(ivf, b, d_queries, nq, np, cr));
            else if (split3)
                VERS_TRY(run_list_scan_tc<1>(ivf, b, d_queries, nq, np, cr));
            else
                VERS_TRY(run_list_scan_tc<0>(ivf, b, d_queries, nq, np, cr));
            nsplit = TC_PARTS;
        } else {
            VERS_TRY(run_group(ivf, b, nq, np, nullptr, nullptr, true));
            VERS_TRY((run_list_scan<StreamCfg, 1>(ivf, b, d_queries, nq, M)));
            nsplit = StreamCfg::NSPLIT;
        }
        VERS_TRY(launch_cand_merge(ctx, M, b.part_d, b.part_p, b.pair_chunk_off, nq, np, nsplit, b.cand_pos, b.cand_key,
                                   b.cand_bound));
        {
            FamilyTimer ftr(ctx, KF_RERANK);
            VERS_TRY(launch_rerank(ctx, M, ivf->d_lm, ivf->d_lm_ids, 0, ivf->ld, d_queries, nq, k, b.cand_pos, b.cand_bound,
                                   ivf->d_nxmax, b.cand_key, use_tc ? (h16 ? 4 : split3 ? 2 : 1) : 0, d_ids, d_d, d_cnt,
                                   b.fail_flag, ivf->d_stats, nullptr, nullptr, h16 ? ivf->d_h16_stat : nullptr));
        }
        qmask = b.fail_flag;  // 2b. exact-order redo of the (rare) uncertified queries, no host round trip
    }
    // 2b / exact mode: exact-order scan of the probed lists, merge by (distance, id)
    // (approximate path: every kernel of this pass returns at once when no query failed its certificate)
    const uint32_t* skip = approx ? reinterpret_cast<const uint32_t*>(ivf->d_stats + 4) : nullptr;
    VERS_TRY(run_group(ivf, b, nq, np, nullptr, qmask, !approx, ScanCfg::TB, exact_chunk, exact_chunk, 0xffffffffu, nullptr,
                       skip));
    VERS_TRY((run_list_scan<ScanCfg, 0>(ivf, b, d_queries, nq, k, skip, exact_chunk)));
    MergeParams mp;
    mp.part_d = b.part_d;
    mp.part_p = b.part_p;
    mp.seg = b.pair_chunk_off;
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
    mp.qmask = qmask;
    mp.skip_if_zero = skip;
    return launch_merge(ctx, mp);
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
    cudaFree(ivf->d_lm_norm);
    cudaFree(ivf->d_lm16);
    cudaFree(ivf->d_h16_stat);
    ivf->call_graph.reset();
    cudaFree(ivf->d_nxmax);
    cudaFree(ivf->d_cent_norm);
    cudaFree(ivf->d_ncmax);
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

namespace vers {
__global__ void remap_ids_kernel(uint64_t* __restrict__ lm_ids, uint64_t n, uint64_t id_base,
                                 const uint64_t* __restrict__ row_ids) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x)
        lm_ids[j] = row_ids[lm_ids[j] - id_base];
}
}  // namespace vers

extern "C" int32_t vers_ivf_from_parts_dev(vers_dataset* ds, const float* d_centroids, uint32_t num_clusters,
                                           const uint32_t* d_assignments, const uint64_t* d_row_ids, vers_ivf** out) {
    if (!ds || !d_centroids || !out || (!d_assignments && ds->n))
        return fail(VERS_ERR_ARG, "ivf_from_parts_dev: null argument");
    *out = nullptr;
    vers_ctx* ctx = ds->ctx;
    VERS_CUDA(cudaSetDevice(ctx->device));
    vers_kmeans* km = nullptr;
    VERS_TRY(vers_kmeans_create(ds, num_clusters, &km));
    cudaError_t e = cudaMemcpyAsync(km->d_cents, d_centroids, (size_t)num_clusters * ds->ld * 4, cudaMemcpyDeviceToDevice,
                                    ctx->stream);
    if (e == cudaSuccess && ds->n)
        e = cudaMemcpyAsync(km->d_assign, d_assignments, ds->n * 4, cudaMemcpyDeviceToDevice, ctx->stream);
    km->csr_valid = false;
    int32_t rc = e == cudaSuccess ? VERS_OK : fail(VERS_ERR_CUDA, "ivf_from_parts_dev: %s", cudaGetErrorString(e));
    if (rc == VERS_OK) rc = ivf_from_state(km, 0.f, 0, out);
    if (rc == VERS_OK && d_row_ids && ds->n) {
        remap_ids_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>((*out)->d_lm_ids, ds->n, ds->id_base, d_row_ids);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "ivf_from_parts_dev: %s", cudaGetErrorString(e));
    }
    if (rc != VERS_OK) {
        vers_ivf_free(*out);
        *out = nullptr;
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
    std::lock_guard<std::recursive_mutex> lk(ivf->ctx->mu);
    VERS_CUDA(cudaSetDevice(ivf->ctx->device));
    return ivf_search_dev_locked(ivf, d_queries, nq, top_k, nprobe, d_ids, d_dists, d_counts);
}

namespace vers {
int32_t upload_queries(vers_ctx* ctx, const float* q, uint32_t nq, uint32_t stride, uint32_t dim, uint32_t ld,
                       float** d_q);
}

extern "C" int32_t vers_ivf_probe_dev(vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t nprobe,
                                      uint64_t* d_probe_ids) {
    if (!ivf || (!d_queries && nq) || !d_probe_ids) return fail(VERS_ERR_ARG, "ivf_probe_dev: null argument");
    if (nq == 0 || nprobe == 0) return VERS_OK;
    if (nprobe > ivf->C || nprobe > VERS_MAX_TOPK) return fail(VERS_ERR_ARG, "ivf_probe_dev: nprobe %u out of range", nprobe);
    vers_ctx* ctx = ivf->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    const RankTable ctab = centroid_table(ivf);
    const ProbePlan pplan = probe_plan(ctx, ctab, nq, nprobe);
    VERS_TRY(scratch_reserve(ctx, pplan.bytes + (size_t)nq * nprobe * 4 + 256));
    VERS_CUDA(cudaMemsetAsync(ivf->d_stats, 0, 128, ctx->stream));
    float* d_pd = reinterpret_cast<float*>((char*)ctx->scratch + pplan.bytes);
    return probe_run(ctx, ctab, pplan, d_queries, nq, nprobe, d_probe_ids, d_pd, nullptr, KF_PROBE);
}

extern "C" int32_t vers_ivf_search_probed_dev(vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t top_k,
                                              uint32_t nprobe, const uint64_t* d_probe_ids, uint64_t* d_ids,
                                              float* d_dists, uint32_t* d_counts) {
    if (!ivf || (!d_queries && nq) || !d_ids || !d_dists || !d_probe_ids)
        return fail(VERS_ERR_ARG, "ivf_search_probed_dev: null argument");
    if (top_k > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "top_k %u > %u", top_k, VERS_MAX_TOPK);
    if (nprobe == 0 || nprobe > ivf->C) return fail(VERS_ERR_ARG, "ivf_search_probed_dev: nprobe %u out of range", nprobe);
    if (nq == 0 || top_k == 0) return VERS_OK;
    std::lock_guard<std::recursive_mutex> lk(ivf->ctx->mu);
    VERS_CUDA(cudaSetDevice(ivf->ctx->device));
    return ivf_search_dev_locked(ivf, d_queries, nq, top_k, nprobe, d_ids, d_dists, d_counts, d_probe_ids);
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
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
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
    if (q_stride_floats == ivf->ld && ivf->ld == ivf->dim) {  // dense rows: one linear copy
        VERS_CUDA(cudaMemcpyAsync(d_q, queries, (size_t)nq * ivf->ld * 4, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        if (ivf->ld != ivf->dim) VERS_CUDA(cudaMemsetAsync(d_q, 0, (size_t)nq * ivf->ld * 4, ctx->stream));
        VERS_CUDA(cudaMemcpy2DAsync(d_q, (size_t)ivf->ld * 4, queries, (size_t)q_stride_floats * 4,
                                    (size_t)ivf->dim * 4, nq, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (nprobe == 0) {  // reference semantics: reads a flag back mid-call, not capturable
        VERS_TRY(ivf_search_dev_locked(ivf, d_q, nq, top_k, nprobe, d_ids, d_d, d_c));
    } else {
        uint64_t key[12] = {reinterpret_cast<uint64_t>(ivf), nq, top_k, nprobe, 0, 0, 0, 0,
                            reinterpret_cast<uint64_t>(ctx->scratch), reinterpret_cast<uint64_t>(d_q),
                            reinterpret_cast<uint64_t>(ctx->stream), ctx->scratch_bytes};
        ivf_state_stamp(ivf, key + 4);
        VERS_TRY(graph_cached_run(ctx, ivf->call_graph, key, [&]() {
            return ivf_search_dev_locked(ivf, d_q, nq, top_k, nprobe, d_ids, d_d, d_c);
        }));
    }
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

extern "C" int32_t vers_ivf_set_mode(vers_ivf* ivf, int32_t mode) {
    if (!ivf || mode < 0 || mode > 4) return fail(VERS_ERR_ARG, "ivf_set_mode: bad argument");
    ivf->mode = mode;
    return VERS_OK;
}

extern "C" int32_t vers_ivf_last_search_stats(const vers_ivf* ivf, uint64_t out[8]) {
    if (!ivf || !out) return fail(VERS_ERR_ARG, "ivf_last_search_stats: null");
    VERS_CUDA(cudaSetDevice(ivf->ctx->device));
    uint64_t tmp[16];
    VERS_CUDA(cudaMemcpyAsync(tmp, ivf->d_stats, 128, cudaMemcpyDeviceToHost, ivf->ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ivf->ctx->stream));
    for (int i = 0; i < 7; ++i) out[i] = tmp[i];
    out[7] = (tmp[7] & 0xffffffffull) | (tmp[8] << 32);  // probe: uncertified queries | centroids re-ranked
    return VERS_OK;
}

extern "C" int32_t vers_ivf_add(vers_ivf* ivf, const float* embedding, uint64_t vec_id, uint64_t* assigned_id,
                                uint32_t* cluster) {
    (void)vec_id;  // the reference ignores the caller's id: `let vec_id = self.assignments.len()` (ivfflat.rs:209)
    if (!ivf || !embedding) return fail(VERS_ERR_ARG, "ivf_add: null argument");
    vers_ctx* ctx = ivf->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    // nearest centroid, first minimum on ties (min_by, ivfflat.rs:201-207) == top-1 by (d, centroid index): the same
    // probe as a one-query search (staging from the context's arenas: no allocation per add)
    const RankTable ctab = centroid_table(ivf);
    const ProbePlan pplan = probe_plan(ctx, ctab, 1, 1);
    VERS_TRY(scratch_reserve(ctx, pplan.bytes + 256));
    VERS_TRY(io_reserve(ctx, (size_t)ivf->ld * 4 + 512));
    float* d_row = reinterpret_cast<float*>(ctx->io);
    uint64_t* d_best = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(ctx->io) + (((size_t)ivf->ld * 4 + 255) & ~size_t(255)));
    float* d_bd = reinterpret_cast<float*>(d_best + 1);
    int32_t rc = VERS_OK;
    uint64_t best = 0;
    float bd = 0.0f;
    {
        std::vector<float> row(ivf->ld, 0.0f);
        memcpy(row.data(), embedding, (size_t)ivf->dim * 4);
        cudaError_t e = cudaMemcpyAsync(d_row, row.data(), (size_t)ivf->ld * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // `row` goes out of scope
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "ivf_add: %s", cudaGetErrorString(e));
    }
    if (rc == VERS_OK) rc = probe_run(ctx, ctab, pplan, d_row, 1, 1, d_best, d_bd, nullptr, KF_PROBE);
    if (rc == VERS_OK) {
        cudaError_t e = cudaMemcpyAsync(&best, d_best, 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&bd, d_bd, 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "ivf_add: %s", cudaGetErrorString(e));
    }
    if (rc == VERS_OK && bd != bd) best = ivf->C;  // a NaN distance: the select kernel cannot order it
    if (rc == VERS_OK && best >= ivf->C)  // every distance compared false (NaN / inf - inf in the embedding)
        rc = fail(VERS_ERR_PANIC, "ivf_add: a distance is NaN (partial_cmp(..).unwrap() panics, ivfflat.rs:207)");
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
            if (e == cudaSuccess) {
                rownorm_kernel<<<1, 32, 0, ctx->stream>>>(ivf->d_lm, ivf->ld, pos, 1, ivf->d_lm_norm, ivf->d_nxmax);
                ctx->launches += 1;
                e = cudaGetLastError();
            }
            if (e == cudaSuccess && ivf->h16_valid) {
                // keep the 16-bit candidate copy current: convert the one row in place when it fits the copy's scale
                double n2 = 0.0;
                for (uint32_t i = 0; i < ivf->dim; ++i) n2 += (double)embedding[i] * (double)embedding[i];
                if (n2 <= ivf->h16_norm2_limit && pos < ivf->h16_cap) {
                    rows_to_h16_kernel<<<1, 32, 0, ctx->stream>>>(ivf->d_lm, ivf->ld, ivf->ld16, nullptr, nullptr, 0, 1, pos,
                                                                 ivf->h16_scale, 1.0f / ivf->h16_scale, ivf->d_lm16,
                                                                 ivf->d_h16_stat, ivf->d_h16_stat + 1);
                    ctx->launches += 1;
                    e = cudaGetLastError();
                } else {
                    ivf->h16_valid = false;  // rebuilt with a new scale by the next search
                }
            }
            ivf->seg_len[c] += 1;
            ivf->seg_epoch += 1;
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
    return rc;
}

namespace vers {
// appended rows -> their list slots: row, global id, ||row||^2 (any order) and the running maximum
__global__ void add_scatter_kernel(const float* __restrict__ rows, uint32_t ld, const uint64_t* __restrict__ pos,
                                   uint64_t n, uint64_t first_id, float* __restrict__ lm, uint64_t* __restrict__ lm_ids,
                                   float* __restrict__ lm_norm, uint32_t* nxmax) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    float mx = 0.0f;
    for (uint64_t j = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); j < n; j += warps) {
        const float4* src = reinterpret_cast<const float4*>(rows + j * ld);
        float4* dst = reinterpret_cast<float4*>(lm + pos[j] * ld);
        float s = 0.0f;
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            const float4 v = src[c];
            dst[c] = v;
            s = __fmaf_rn(v.x, v.x, s);
            s = __fmaf_rn(v.y, v.y, s);
            s = __fmaf_rn(v.z, v.z, s);
            s = __fmaf_rn(v.w, v.w, s);
        }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
        if (lane == 0) {
            lm_norm[pos[j]] = s;
            lm_ids[pos[j]] = first_id + j;
        }
        mx = fmaxf(mx, s);
    }
    if (lane == 0 && mx > 0.0f) atomicMax(nxmax, __float_as_uint(mx));
}
}  // namespace vers

// Index::add (ivfflat.rs:200-213) for a batch, in order: embedding i gets id assignments.len() + i and goes to its
// nearest centroid (first minimum).  The centroids do not move on add, so this is exactly n sequential adds: one
// exact-order assign launch over the batch, one relayout at most, one scatter into the list slots.
extern "C" int32_t vers_ivf_add_batch(vers_ivf* ivf, const float* embeddings, uint64_t n, uint32_t stride_floats,
                                      uint64_t* assigned_ids, uint32_t* clusters) {
    if (!ivf || (!embeddings && n)) return fail(VERS_ERR_ARG, "ivf_add_batch: null argument");
    if (stride_floats < ivf->dim) return fail(VERS_ERR_ARG, "ivf_add_batch: stride < dim");
    if (n == 0) return VERS_OK;
    if (n >= 0x7fffffffull || ivf->n + n >= 0xffffffffull) return fail(VERS_ERR_UNSUPPORTED, "ivf_add_batch: too many rows");
    vers_ctx* ctx = ivf->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    float* d_rows = nullptr;
    uint32_t *d_assign = nullptr, *d_bad = nullptr;
    uint64_t* d_pos = nullptr;
    auto cleanup = [&]() { cudaFree(d_rows), cudaFree(d_assign), cudaFree(d_bad), cudaFree(d_pos); };
    cudaError_t e = cudaMalloc(&d_rows, n * ivf->ld * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_assign, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_bad, 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_pos, n * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, 4, s);
    if (e == cudaSuccess && ivf->ld != ivf->dim) e = cudaMemsetAsync(d_rows, 0, n * ivf->ld * 4, s);
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(d_rows, (size_t)ivf->ld * 4, embeddings, (size_t)stride_floats * 4, (size_t)ivf->dim * 4, n,
                              cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) {
        cleanup();
        return fail(e == cudaErrorMemoryAllocation ? VERS_ERR_NOMEM : VERS_ERR_CUDA, "ivf_add_batch: %s", cudaGetErrorString(e));
    }
    // nearest centroid of every new row, first minimum on ties (min_by, ivfflat.rs:201-207)
    RowSrc A{d_rows, nullptr, ivf->ld, n};
    int32_t rc = kmeans_assign_rows(ctx, A, ivf->d_cents, ivf->C, ivf->ld, d_assign, KF_ASSIGN, d_bad);
    std::vector<uint32_t> assign(n);
    uint32_t bad = 0;
    if (rc == VERS_OK) {
        e = cudaMemcpyAsync(assign.data(), d_assign, n * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "ivf_add_batch: %s", cudaGetErrorString(e));
    }
    if (rc == VERS_OK && bad)
        rc = fail(VERS_ERR_PANIC, "ivf_add: a distance is NaN (partial_cmp(..).unwrap() panics, ivfflat.rs:207)");
    if (rc != VERS_OK) {
        cleanup();
        return rc;  // nothing was modified
    }
    std::vector<uint32_t> extra(ivf->C, 0);
    for (uint64_t i = 0; i < n; ++i) extra[assign[i]] += 1;
    bool grow = false;
    for (uint32_t c = 0; c < ivf->C; ++c) grow = grow || ivf->seg_len[c] + extra[c] > ivf->seg_cap[c];
    if (grow) rc = ivf_relayout(ivf, extra.data());
    if (rc != VERS_OK) {
        cleanup();
        return rc;
    }
    std::vector<uint64_t> pos(n);
    std::vector<uint32_t> fill(ivf->C, 0);
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t c = assign[i];
        pos[i] = ivf->seg_off[c] + ivf->seg_len[c] + fill[c]++;
    }
    e = cudaMemcpyAsync(d_pos, pos.data(), n * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        add_scatter_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(d_rows, ivf->ld, d_pos, n, ivf->id_base + ivf->n, ivf->d_lm,
                                                            ivf->d_lm_ids, ivf->d_lm_norm, ivf->d_nxmax);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        cleanup();
        return fail(VERS_ERR_CUDA, "ivf_add_batch: %s", cudaGetErrorString(e));
    }
    for (uint32_t c = 0; c < ivf->C; ++c) ivf->seg_len[c] += extra[c];
    ivf->seg_epoch += 1;
    ivf->h16_valid = false;  // bulk change: the 16-bit candidate copy is rebuilt by the next search that wants it
    rc = ivf_upload_segments(ivf);
    for (uint64_t i = 0; i < n; ++i) {
        ivf->assign_tail.push_back(assign[i]);
        if (assigned_ids) assigned_ids[i] = ivf->id_base + ivf->n + i;
        if (clusters) clusters[i] = assign[i];
    }
    ivf->n += n;
    cleanup();
    return rc;
}
