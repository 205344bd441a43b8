// ivf_tc.cuh — tensor-core candidate pass of the inverted-list scan (included by ivf.cu).
//
// The dot products x.q of the candidate pass run on the 5th-generation tensor cores instead of the fp32 SIMT pipe:
//   * TMA (cp.async.bulk.tensor, 128-byte swizzle) streams 128-row x 32-float tiles of a list straight from the fp32
//     list-major rows into shared memory — no 16-bit copy of the database, no register staging;
//   * tcgen05.mma kind::tf32 (M=128, K=8) multiplies them with a group of the list's queries, accumulators in TMEM
//     (two buffers, so the MMAs of tile t+1 overlap the epilogue of tile t).  A group holds up to 32 queries: lists
//     probed by <= 16 queries run with N=16, busier lists with N=32, so a list is streamed from HBM once per 32 of its
//     queries;
//   * SPLIT3 (default): TF32 keeps 10 mantissa bits, too coarse for the rounding-error certificate to separate real
//     neighbours, so every operand is split x = hi + lo (hi = x with the low 13 mantissa bits cleared, lo = x - hi,
//     exact in fp32) and three products are formed per K-step: hi.hi + lo.hi + hi.lo (the dropped lo.lo term and the
//     truncation of lo are <= 3 * 2^-20 |x||q| per product).  The tensor core itself truncates fp32 -> tf32 (measured:
//     bit-identical results with and without clearing the low bits first), so the row tile as loaded IS x_hi; four
//     converter warps compute x_lo from the landed tile and park it in TENSOR MEMORY (tcgen05.st), from where the
//     lo.hi MMA takes its A operand — shared memory sees each row byte only three times (TMA write, converter
//     read, MMA read).  hi.hi and hi.lo are ONE MMA against the concatenated [q_hi; q_lo] tile (N = 2 NQ); the
//     column blocks are summed in the epilogue.  The queries are split once by the gather kernel;
//   * the epilogue (8 warps: 4 TMEM lane groups x 2 column halves) forms the key ||x||^2 - 2 x.q per (row, query) and
//     keeps the 32 smallest (key, row) per (lane group, query): rows that beat the query's current 32nd key are
//     appended to a 32-entry shared-memory queue with one ballot; a full queue is sorted across the lanes (bitonic
//     network) and merged into the sorted list, which tightens the threshold.  Selection cost is per batch of 32
//     survivors, not per survivor, so the epilogue stays far below the HBM time of a tile.
//   * PREC 2 (16-bit candidate copy, BASELINE configs[3] "16-bit candidate / fp32 rerank"): the rows come from an fp16
//     copy of the list-major table (x~ = fp16(x * 2^e), half the HBM bytes per row), 128-row x 64-half boxes, one
//     kind::f16 MMA per 16 dimensions against [q_hi; q_lo] (q_hi = fp16(q), q_lo = fp16((q - q_hi) * 2^11): the query is
//     represented to ~2^-22, so the only real error is (x - x~).q, bounded through max ||x - x~|| by Cauchy-Schwarz in
//     the certificate).  No converter warps, no second issuer; the epilogue reads 8 accumulator columns per load.
// What leaves the kernel is the same (key, position) partial lists as the SIMT candidate pass, so the merge ->
// exact-order rerank -> certificate -> exact redo chain behind it is unchanged.  The kernel does no fp32 SIMT math
// per (row, query, dim): it is an HBM stream.
//
// Warp roles (16 warps): 0 = TMA producer, 1 = TMEM allocator + main MMA issuer, 2 = x_lo MMA issuer,
// 3 = scheduler (claims work items from the atomic counter one item ahead of the producer, decodes them with a
// warp-wide 32-ary search and publishes them DECODED through a shared-memory ring, so no role pays the chain of
// dependent global loads of a decode on its critical path), 4..11 = epilogue (warp w reads TMEM lanes 32*(w%4)..+31 = tile rows; warps 4..7 take the first half of
// the group's query columns, 8..11 the second), 12..15 = hi/lo converters (SPLIT3 only).  mbarriers: full / conv /
// empty per smem stage, tmem_full / tmem_empty per accumulator buffer, sched_full / sched_empty for the work-item
// ring (work items = (list, query group, 4096-row chunk), handed out dynamically through an atomic counter).
#pragma once
#include "tc.cuh"

namespace vers {

constexpr int TC_M = 128, TC_NQ = 32, TC_KC = 32, TC_SCHED = 4;
constexpr int TC_PARTS = 4;       // partial lists per (pair, chunk): one per TMEM lane group
constexpr int TC_SLOT_RANK = 32 / TC_PARTS;  // a list publishes its 8th key into its lane group's slot of the shared bound
static_assert(TC_PARTS == 4, "the shared bound is read as one uint4 per query");
constexpr int TC_EPI_WARP0 = 4, TC_EPI_WARPS = 8, TC_CONV_WARP0 = 12, TC_CONV_WARPS = 4;
constexpr int TC_QPW = TC_NQ / 2;  // query columns per epilogue warp (at most)
constexpr int TC_QCAP = 32;       // queue entries per (epilogue warp, query)
constexpr int TC_A_BYTES = TC_M * TC_KC * 4;

// PREC: 0 = tf32 operands as loaded, 1 = split-precision tf32 (hi/lo, three products), 2 = fp16 row copy x [q_hi; q_lo]
template <int PREC>
struct TcCfg {
    static constexpr bool SPLIT3 = PREC == 1, H16 = PREC == 2;
    static constexpr int STAGES = 7;
    static constexpr int THREADS = SPLIT3 ? 512 : 384;
    static constexpr int KC_ELEMS = H16 ? 64 : TC_KC;              // K elements per stage: one 128-byte swizzle row
    static constexpr int B_ROWS = (SPLIT3 || H16) ? 2 * TC_NQ : TC_NQ;  // B rows per stage at most: [q_hi; q_lo] or q
    // per buffer: [hi.hi | hi.lo | lo.hi], [x~.q_hi | x~.q_lo] or [x.q]
    static constexpr int ACC_COLS = SPLIT3 ? 3 * TC_NQ : (H16 ? 2 * TC_NQ : TC_NQ);
    static constexpr int STAGE_BYTES = TC_A_BYTES + B_ROWS * 128;
    static constexpr int OFF_B = TC_A_BYTES;
    static constexpr int ALO_COL0 = 2 * ACC_COLS;                  // x_lo tiles live in TMEM after the accumulators
    static constexpr uint32_t TMEM_COLS = SPLIT3 ? 512 : (H16 ? 128 : 64);  // 192 + 7*32 = 416 -> 512 / 2*64 / 2*32
    // per epilogue warp: sorted list + queue, keys (fp32) and row offsets (u16), for TC_QPW queries
    static constexpr int SEL_WARP_BYTES = TC_QPW * (32 + TC_QCAP) * 6;
    static constexpr int OFF_SEL = STAGES * STAGE_BYTES;
    static constexpr int OFF_BAR = OFF_SEL + TC_EPI_WARPS * SEL_WARP_BYTES;
    static constexpr int SMEM_BYTES = 1024 + OFF_BAR + 1024;
};

// queries regrouped by list (row i = the i-th grouped (query, list) pair; lq_query == null: row i = query i), split
// into tf32 hi and lo parts
__global__ void gather_queries_kernel(const float* __restrict__ queries, const uint32_t* __restrict__ lq_query,
                                      const uint64_t* __restrict__ lq_off, uint32_t C, uint32_t ld,
                                      float* __restrict__ gq_hi, float* __restrict__ gq_lo, int split) {
    const uint64_t n = lq_off[C];
    const uint32_t ld4 = ld >> 2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * ld4;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / ld4;
        uint32_t c = (uint32_t)(i - r * ld4);
        float4 v = reinterpret_cast<const float4*>(queries)[(uint64_t)(lq_query ? lq_query[r] : (uint32_t)r) * ld4 + c];
        if (split) {
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = __fsub_rn(v.x, h.x);
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = __fsub_rn(v.y, h.y);
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = __fsub_rn(v.z, h.z);
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = __fsub_rn(v.w, h.w);
            reinterpret_cast<float4*>(gq_hi)[i] = h;
            reinterpret_cast<float4*>(gq_lo)[i] = l;
        } else {
            reinterpret_cast<float4*>(gq_hi)[i] = v;
        }
    }
}


// PREC 2: the grouped queries as fp16 [q_hi; q_lo] rows of ld16 halfs (ld16 % 8 == 0, columns past ld zero):
// q_hi = fp16(q), q_lo = fp16((q - q_hi) * 2^11) (the scaling keeps q_lo in the normal fp16 range whenever q_hi is)
constexpr float H16_LO_SCALE = 2048.0f, H16_LO_INV = 1.0f / 2048.0f;
__device__ __forceinline__ void h16_split(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(__fmul_rn(__fsub_rn(v, __half2float(hi)), H16_LO_SCALE));
}
__global__ void gather_queries_h16_kernel(const float* __restrict__ queries, const uint32_t* __restrict__ lq_query,
                                          const uint64_t* __restrict__ lq_off, uint32_t C, uint32_t ld, uint32_t ld16,
                                          __half* __restrict__ gq_hi, __half* __restrict__ gq_lo) {
    const uint64_t n = lq_off[C];
    const uint32_t ld8 = ld16 >> 3;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * ld8;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / ld8;
        const uint32_t c = (uint32_t)(i - r * ld8) * 8;
        const float* q = queries + (uint64_t)(lq_query ? lq_query[r] : (uint32_t)r) * ld + c;
        const float4 a = *reinterpret_cast<const float4*>(q);  // ld % 4 == 0 and c < ld
        const float4 b = c + 4 < ld ? *reinterpret_cast<const float4*>(q + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        __align__(16) __half h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) h16_split(v[e], h[e], l[e]);
        reinterpret_cast<uint4*>(gq_hi)[i] = *reinterpret_cast<const uint4*>(h);
        reinterpret_cast<uint4*>(gq_lo)[i] = *reinterpret_cast<const uint4*>(l);
    }
}

// the fp16 candidate copy of the list-major rows: lm16[pos][c] = fp16(lm[pos][c] * scale) (scale = 2^e chosen by the
// host so that no element can overflow), one block per list (grid-stride), one warp per row.  Also the maximum over the
// live rows of ||x - x~||^2 (x~ = the value the copy represents) for the certificate, and a count of elements that
// did not fit (must stay 0: the host's scale guarantees it; the certificate refuses everything otherwise).
__global__ void __launch_bounds__(256) rows_to_h16_kernel(const float* __restrict__ lm, uint32_t ld, uint32_t ld16,
                                                          const uint64_t* __restrict__ seg_off,
                                                          const uint32_t* __restrict__ seg_len, uint32_t list0,
                                                          uint32_t nlists, uint64_t pos_override, float scale,
                                                          float inv_scale, __half* __restrict__ lm16,
                                                          uint32_t* __restrict__ xlo2max, uint32_t* __restrict__ bad) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t ld8 = ld16 >> 3;
    float mx = 0.0f;
    uint32_t nbad = 0;
    for (uint32_t l = list0 + blockIdx.x; l < list0 + nlists; l += gridDim.x) {
        // seg_off == null: one explicit row at pos_override (in-place update after Index::add)
        const uint64_t off = seg_off ? seg_off[l] : pos_override;
        const uint32_t len = seg_len ? seg_len[l] : 1u;
        for (uint32_t j = warp; j < len; j += nwarps) {
            const float* x = lm + (off + j) * ld;
            __half* y = lm16 + (off + j) * ld16;
            float s = 0.0f;
            for (uint32_t c8 = lane; c8 < ld8; c8 += 32) {
                const uint32_t c = c8 * 8;
                const float4 a = *reinterpret_cast<const float4*>(x + c);
                const float4 b = c + 4 < ld ? *reinterpret_cast<const float4*>(x + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                __align__(16) __half h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    h[e] = __float2half_rn(__fmul_rn(v[e], scale));
                    const float back = __fmul_rn(__half2float(h[e]), inv_scale);
                    const float r = __fsub_rn(v[e], back);
                    s = __fmaf_rn(r, r, s);
                    if (!(fabsf(back) <= 3.0e38f)) ++nbad;  // inf or nan
                }
                reinterpret_cast<uint4*>(y)[c8] = *reinterpret_cast<const uint4*>(h);
            }
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
            mx = fmaxf(mx, s);
        }
    }
    if (lane == 0 && mx > 0.0f) atomicMax(xlo2max, __float_as_uint(mx));
    if (nbad) atomicAdd(bad, nbad);
}


struct TcScanParams {
    uint32_t ld, C;            // ld: K elements of a row (PREC 2: halfs of the fp16 copy)
    uint32_t chunk_rows;       // rows per work item (multiple of 128, <= 65535)
    uint32_t chunk_rows_tail;  // ... for the lists >= tail_list0: the items handed out last are small, so the
    uint32_t tail_list0;       //     persistent CTAs finish within a fraction of a full item of each other
    const uint64_t* seg_off;
    const uint32_t* seg_len;
    const uint32_t* lq_pair;
    const uint64_t* lq_off;
    const uint64_t* item_off;
    const uint64_t* pair_chunk_off;
    const float* lm_norm;
    float* part_d;
    uint32_t* part_p;
    unsigned long long* counter;
    // optional (null: off): per-query running bound shared by all work items of the launch, [nq][TC_PARTS] in an
    // order-preserving uint32 encoding.  Slot g of query q holds the smallest 8th key any list of TMEM lane group g
    // (any item, any CTA) has reached: that list alone holds 8 rows at or below it, the lane groups scan disjoint rows,
    // so 4 x 8 = 32 distinct rows lie at or below the LARGEST of the four slots — a row whose key exceeds that cannot be
    // among the query's 32 best keys.  (One slot fed with whole lists' 32nd keys is the same argument with one list;
    // it stays near the 32/rows-per-list quantile, 3.7 % for 864 rows, where 62 % of all 32-row slices still hold a
    // passing row; the minimum over hundreds of lists of an 8th key is ~0.2 %.)  Later items start selective instead of
    // re-learning the threshold from +inf.  The merged top-32 and its bound do not depend on the timing of these
    // updates.  Requires merged M == 32.
    uint32_t* qtau;
    const uint32_t* lq_query;  // grouped pair -> query
    // optional (null: off): no selection at all, every key is written to dense_out[grouped pair][row position]
    // (the centroid probe of small tables: C keys per query, selected afterwards by probe_select_kernel)
    float* dense_out;
    uint64_t dense_ld;
    float key_scale;  // PREC 2: key = ||x||^2 + key_scale * acc, key_scale = -2 / (row scale 2^e); else unused (-2)
};

__device__ __forceinline__ uint32_t tau_encode(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float tau_decode(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
constexpr uint32_t TAU_INF = 0xff800000u;  // tau_encode(+inf)

struct TcItem {
    uint32_t list, chunk;
    uint32_t nB, nq;  // live queries of the group; MMA width of the group (16 or 32)
    uint64_t q0, base_pos, r0, r1;
};

constexpr uint32_t TC_ITEM_END = 0xffffffffu;  // TcItem::list of the ring entry that ends the launch

// item number -> (list, query group, row chunk) by a whole converged warp: 32-ary search over item_off (3 rounds for
// C <= 32768 instead of 15 dependent loads), then the item's table entries in one round of independent loads
__device__ __forceinline__ TcItem tc_decode_item_warp(const TcScanParams& p, uint64_t it, int lane) {
    uint32_t lo = 0, n = p.C;  // invariant: item_off[lo] <= it, the answer lies in [lo, lo + n)
    while (n > 1) {
        const uint32_t step = (n + 31) / 32, idx = lo + (uint32_t)lane * step;
        const bool ok = idx < lo + n && p.item_off[idx] <= it;  // monotone in the lane; lane 0 always holds
        const unsigned m = __ballot_sync(FULL_MASK, ok);
        const uint32_t j = 31u - (uint32_t)__clz((int)m);
        const uint32_t hi = lo + n;
        lo += j * step;
        n = min(step, hi - lo);
    }
    TcItem t;
    t.list = lo;
    const uint64_t io = p.item_off[lo], l0 = p.lq_off[lo], l1 = p.lq_off[lo + 1], so = p.seg_off[lo];
    const uint32_t len = p.seg_len[lo];
    const uint64_t local = it - io;
    const uint32_t cr = lo >= p.tail_list0 ? p.chunk_rows_tail : p.chunk_rows;
    const uint32_t nch = (len + cr - 1) / cr;
    const uint32_t group = (uint32_t)(local / nch);
    t.chunk = (uint32_t)(local % nch);
    t.q0 = l0 + (uint64_t)group * TC_NQ;
    t.nB = (uint32_t)min((uint64_t)TC_NQ, (l1 - l0) - (uint64_t)group * TC_NQ);
    t.nq = t.nB <= 16 ? 16u : 32u;
    t.base_pos = so;
    t.r0 = (uint64_t)t.chunk * cr;
    t.r1 = min((uint64_t)len, t.r0 + cr);
    return t;
}

// ---- selection helpers of the epilogue: entries are (key, row offset in the item), ordered by key then offset
__device__ __forceinline__ bool sel_less(float d0, uint32_t r0, float d1, uint32_t r1) {
    return d0 < d1 || (d0 == d1 && r0 < r1);
}
// one compare-exchange step of a bitonic network across the lanes: partner = lane ^ j, keep the smaller entry when
// keep_min, else the larger
__device__ __forceinline__ void sel_cmpx(float& d, uint32_t& r, int j, bool keep_min) {
    const float od = __shfl_xor_sync(FULL_MASK, d, j);
    const uint32_t orr = __shfl_xor_sync(FULL_MASK, r, j);
    const bool other_less = sel_less(od, orr, d, r);
    const bool self_less = sel_less(d, r, od, orr);
    if (keep_min ? other_less : self_less) {
        d = od;
        r = orr;
    }
}
// Folds the c (<= 32) queued entries of one query into its sorted 32-entry list; returns the new 32nd key.
//   lk/lr: the list (ascending, lane i = i-th smallest); qk/qr: the queue.  Whole warp, converged.
// returns (the list's 32nd key, its 8th key): +inf while the list holds fewer entries
__device__ __noinline__ float2 sel_flush(float* lk, uint16_t* lr, const float* qk, const uint16_t* qr, uint32_t c, int lane) {
    __syncwarp();  // the queue writes of the other lanes are visible
    float d = (uint32_t)lane < c ? qk[lane] : __int_as_float(0x7f800000);
    uint32_t r = (uint32_t)lane < c ? (uint32_t)qr[lane] : 0xffffu;
    // sort the batch DESCENDING across the lanes (bitonic sorting network, 15 steps)
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool desc_block = (lane & k) == 0 || k == 32;  // final pass: one descending run
            const bool lower = (lane & j) == 0;
            sel_cmpx(d, r, j, desc_block ? !lower : lower);
        }
    }
    // list ascending, batch descending: the lane-wise minimum is the smaller half of the union, as a bitonic sequence
    float ld_ = lk[lane];
    uint32_t lrr = lr[lane];
    if (sel_less(d, r, ld_, lrr)) {
        ld_ = d;
        lrr = r;
    }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) sel_cmpx(ld_, lrr, j, (lane & j) == 0);  // bitonic merge, ascending
    lk[lane] = ld_;
    lr[lane] = (uint16_t)lrr;
    __syncwarp();
    return make_float2(__shfl_sync(FULL_MASK, ld_, 31), __shfl_sync(FULL_MASK, ld_, TC_SLOT_RANK - 1));
}
// the shared bound of query q: the largest of its TC_PARTS slots (see TcScanParams::qtau)
__device__ __forceinline__ uint32_t tau_shared_bits(const uint32_t* qtau, uint32_t q) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(qtau) + q);
    return max(max(v.x, v.y), max(v.z, v.w));
}

template <int PREC>
__global__ void __launch_bounds__(TcCfg<PREC>::THREADS, 1)
    tc_list_scan_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_rows32,
                        const __grid_constant__ CUtensorMap tmap_qhi16,
                        const __grid_constant__ CUtensorMap tmap_qlo16, const __grid_constant__ CUtensorMap tmap_qhi32,
                        const __grid_constant__ CUtensorMap tmap_qlo32, TcScanParams p) {
    using Cfg = TcCfg<PREC>;
    constexpr bool SPLIT3 = Cfg::SPLIT3, H16 = Cfg::H16;
    constexpr bool TWO_B = SPLIT3 || H16;  // the B stage holds [q_hi; q_lo]
    constexpr int S = Cfg::STAGES;
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t raw = tc::smem_u32(tc_smem_raw);
    uint8_t* smem = tc_smem_raw + (((raw + 1023u) & ~1023u) - raw);  // SWIZZLE_128B tiles need 1024-byte alignment
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* conv = full + S;
    uint64_t* empty = conv + S;
    uint64_t* tfull = empty + S;
    uint64_t* tempty = tfull + 2;
    uint64_t* sfull = tempty + 2;
    uint64_t* sempty = sfull + TC_SCHED;
    uint64_t* pgo = sempty + TC_SCHED;  // the producer has started the item it last received
    TcItem* sched = reinterpret_cast<TcItem*>(pgo + 1);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched + TC_SCHED);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t total_items = p.item_off[p.C];
    const uint32_t nk = (p.ld + Cfg::KC_ELEMS - 1) / Cfg::KC_ELEMS;
    // consumers of a scheduled item: producer + MMA thread + epilogue warps (+ lo MMA thread + converters)
    // (single-thread roles arrive once, whole-warp roles with every lane: each lane's read of the slot is released by its
    // own arrive)
    constexpr uint32_t SCHED_CONSUMERS = 2 + 32 * TC_EPI_WARPS + (SPLIT3 ? 1 + 32 * TC_CONV_WARPS : 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&conv[s], TC_CONV_WARPS * 32);  // every converter lane arrives itself
            tc::mbar_init(&empty[s], SPLIT3 ? 2 : 1);  // one commit per MMA issuer
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tfull[b], SPLIT3 ? 2 : 1);
            tc::mbar_init(&tempty[b], TC_EPI_WARPS);
        }
        for (int i = 0; i < TC_SCHED; ++i) {
            tc::mbar_init(&sfull[i], 1);
            tc::mbar_init(&sempty[i], SCHED_CONSUMERS);
        }
        tc::mbar_init(pgo, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmap_rows);
        tc::tma_prefetch_desc(&tmap_rows32);
        tc::tma_prefetch_desc(&tmap_qhi16);
        tc::tma_prefetch_desc(&tmap_qhi32);
        if (TWO_B) {
            tc::tma_prefetch_desc(&tmap_qlo16);
            tc::tma_prefetch_desc(&tmap_qlo32);
        }
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    // every consumer role walks the same ring of scheduled items
    uint32_t sslot = 0, sphase = 0;
    auto next_item = [&](bool leader, bool whole_warp) -> TcItem {
        tc::mbar_wait(&sfull[sslot], sphase);
        const TcItem it = sched[sslot];
        if (whole_warp) __syncwarp();  // keeps the warp converged for the .aligned tcgen05 instructions that follow
        if (whole_warp || leader) tc::mbar_arrive(&sempty[sslot]);
        if (++sslot == TC_SCHED) {
            sslot = 0;
            sphase ^= 1;
        }
        return it;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            while (true) {
                const TcItem t = next_item(true, false);
                if (t.list == TC_ITEM_END) break;
                tc::mbar_arrive(pgo);  // the scheduler may claim (and decode) the next item while this one streams
                const bool wide = t.nq == 32;
                const CUtensorMap* mhi = wide ? &tmap_qhi32 : &tmap_qhi16;
                const CUtensorMap* mlo = wide ? &tmap_qlo32 : &tmap_qlo16;
                const uint32_t b_bytes = t.nq * TC_KC * 4;
                for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                    // the last tile of an item fetches only the 32-row boxes it needs (the rows behind them belong
                    // to the next list: streaming them would be pure waste); the rest of the stage keeps stale,
                    // finite row data whose products land in accumulator rows nobody reads
                    const uint32_t rem = (uint32_t)min((uint64_t)TC_M, t.r1 - a0);
                    const uint32_t nbox = rem > 96 ? 0u : (rem + 31) / 32;  // 0: one full 128-row box
                    const uint32_t tx = (nbox ? nbox * (TC_A_BYTES / 4) : TC_A_BYTES) + (TWO_B ? 2 : 1) * b_bytes;
                    for (uint32_t kc = 0; kc < nk; ++kc) {
                        tc::mbar_wait(&empty[stage], phase ^ 1);
                        tc::mbar_arrive_expect_tx(&full[stage], tx);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        if (nbox == 0) {
                            tc::tma_load_2d(sa, &tmap_rows, &full[stage], (int32_t)(kc * Cfg::KC_ELEMS), (int32_t)(t.base_pos + a0));
                        } else {
                            for (uint32_t bx = 0; bx < nbox; ++bx)
                                tc::tma_load_2d(sa + bx * (TC_A_BYTES / 4), &tmap_rows32, &full[stage],
                                                (int32_t)(kc * Cfg::KC_ELEMS), (int32_t)(t.base_pos + a0 + bx * 32));
                        }
                        tc::tma_load_2d(sa + Cfg::OFF_B, mhi, &full[stage], (int32_t)(kc * Cfg::KC_ELEMS), (int32_t)t.q0);
                        if (TWO_B)
                            tc::tma_load_2d(sa + Cfg::OFF_B + b_bytes, mlo, &full[stage], (int32_t)(kc * Cfg::KC_ELEMS),
                                            (int32_t)t.q0);
                        if (++stage == S) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, tile_ctr = 0;
            while (true) {
                const TcItem t = next_item(true, false);
                if (t.list == TC_ITEM_END) break;
                // x_hi . [q_hi; q_lo]  (or x . q unsplit)
                const uint32_t idesc_main =
                    H16 ? tc::idesc_f16(TC_M, 2 * t.nq) : tc::idesc_tf32(TC_M, SPLIT3 ? 2 * t.nq : t.nq);
                for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                    const uint32_t buf = tile_ctr & 1, tphase = (tile_ctr >> 1) & 1;
                    tc::mbar_wait(&tempty[buf], tphase ^ 1);  // epilogue has drained this accumulator buffer
                    tc::fence_after_thread_sync();
                    const uint32_t d_tmem = tmem_base + buf * Cfg::ACC_COLS;
                    for (uint32_t kc = 0; kc < nk; ++kc) {
                        tc::mbar_wait(&full[stage], phase);
                        tc::fence_after_thread_sync();
                        const uint32_t sa = tc::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                        const uint64_t da = tc::smem_desc_k_sw128(sa), db = tc::smem_desc_k_sw128(sa + Cfg::OFF_B);
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk) {  // 4 x (8 tf32 | 16 fp16) = one 128-byte row, 32 B per step
                            if (H16)
                                tc::mma_f16(d_tmem, da + 2 * kk, db + 2 * kk, idesc_main, (kc | kk) != 0);
                            else
                                tc::mma_tf32(d_tmem, da + 2 * kk, db + 2 * kk, idesc_main, (kc | kk) != 0);
                        }
                        tc::mma_commit(&empty[stage]);  // frees the smem stage once these MMAs have read it
                        if (++stage == S) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    tc::mma_commit(&tfull[buf]);  // accumulator complete
                    ++tile_ctr;
                }
            }
        }
    } else if (warp == 2) {
        // ===================== second MMA issuer: x_lo (tensor memory) . q_hi -> its own accumulator columns ========
        // (a separate thread so that neither issuer's instruction latency per stage exceeds the HBM time per stage)
        if (SPLIT3 && lane == 0) {
            uint32_t stage = 0, phase = 0, tile_ctr = 0;
            while (true) {
                const TcItem t = next_item(true, false);
                if (t.list == TC_ITEM_END) break;
                const uint32_t idesc_lo = tc::idesc_tf32(TC_M, t.nq);
                for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                    const uint32_t buf = tile_ctr & 1, tphase = (tile_ctr >> 1) & 1;
                    tc::mbar_wait(&tempty[buf], tphase ^ 1);
                    tc::fence_after_thread_sync();
                    const uint32_t d_tmem = tmem_base + buf * Cfg::ACC_COLS + 2 * t.nq;
                    for (uint32_t kc = 0; kc < nk; ++kc) {
                        tc::mbar_wait(&conv[stage], phase);  // x_lo of this stage is in tensor memory (implies full)
                        tc::fence_after_thread_sync();
                        const uint64_t db = tc::smem_desc_k_sw128(tc::smem_u32(smem + stage * Cfg::STAGE_BYTES) + Cfg::OFF_B);
                        const uint32_t a_lo = tmem_base + Cfg::ALO_COL0 + stage * TC_KC;
#pragma unroll
                        for (uint32_t kk = 0; kk < TC_KC / 8; ++kk)
                            tc::mma_tf32_ts(d_tmem, a_lo + 8 * kk, db + 2 * kk, idesc_lo, (kc | kk) != 0);
                        tc::mma_commit(&empty[stage]);  // second arrival: q_hi tile and the x_lo columns are free
                        if (++stage == S) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    tc::mma_commit(&tfull[buf]);
                    ++tile_ctr;
                }
            }
        }
    } else if (warp == 3) {
        // ===================== scheduler: claim -> decode -> publish, one item ahead of the producer =============
        uint32_t k = 0;
        while (true) {
            tc::mbar_wait(&sempty[sslot], sphase ^ 1);
            if (k > 0) tc::mbar_wait(pgo, (k - 1) & 1u);  // item k-1 is streaming: claiming further ahead would only
                                                          // hoard work the other CTAs could take at the end
            unsigned long long itu = 0;
            if (lane == 0) itu = atomicAdd(p.counter, 1ull);
            itu = __shfl_sync(FULL_MASK, itu, 0);
            const bool end = itu >= total_items;
            TcItem t;
            if (end) {
                t.list = TC_ITEM_END, t.chunk = 0, t.nB = 0, t.nq = 16, t.q0 = 0, t.base_pos = 0, t.r0 = 0, t.r1 = 0;
            } else {
                t = tc_decode_item_warp(p, itu, lane);
            }
            if (lane == 0) {
                sched[sslot] = t;
                tc::mbar_arrive(&sfull[sslot]);  // release: the slot is visible to the waiters
            }
            __syncwarp();
            if (++sslot == TC_SCHED) {
                sslot = 0;
                sphase ^= 1;
            }
            if (end) break;
            ++k;
        }
    } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + TC_EPI_WARPS) {
        // ===================== epilogue: TMEM -> keys -> 32 smallest (key, row) per (lane group, query) ==========
        const int lane_group = warp & 3;                   // TMEM lanes this warp may touch
        const uint32_t half = (uint32_t)(warp - TC_EPI_WARP0) >> 2;  // which half of the group's query columns
        uint8_t* sel = smem + Cfg::OFF_SEL + (warp - TC_EPI_WARP0) * Cfg::SEL_WARP_BYTES;
        float* lk = reinterpret_cast<float*>(sel);                       // [TC_QPW][32] sorted list keys
        float* qk = lk + TC_QPW * 32;                                     // [TC_QPW][TC_QCAP] queue keys
        uint16_t* lr = reinterpret_cast<uint16_t*>(qk + TC_QPW * TC_QCAP);  // [TC_QPW][32] list row offsets
        uint16_t* qr = lr + TC_QPW * 32;                                  // [TC_QPW][TC_QCAP] queue row offsets
        const uint32_t lt_mask = (1u << lane) - 1u;
        uint32_t tile_ctr = 0;
        while (true) {
            const TcItem t = next_item(lane == 0, true);
            if (t.list == TC_ITEM_END) break;
            // the group's live queries are split evenly between the two warps of a lane group (a list probed by 8
            // queries costs each of them 4 columns per tile, not 8 and 0)
            const uint32_t first = (t.nB + 1) >> 1;
            const uint32_t col0 = half ? first : 0u, nlive = half ? t.nB - first : first;  // this warp's live queries
            if (p.dense_out) {
                // dense mode: keys straight to global memory, lanes = consecutive rows => coalesced
                for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                    const uint32_t buf = tile_ctr & 1, tphase = (tile_ctr >> 1) & 1;
                    tc::mbar_wait(&tfull[buf], tphase);
                    tc::fence_after_thread_sync();
                    const uint32_t tacc = tmem_base + ((uint32_t)(lane_group * 32) << 16) + buf * Cfg::ACC_COLS;
                    const uint64_t row = a0 + (uint64_t)(lane_group * 32 + lane);
                    const bool rowlive = row < t.r1;
                    const float nx = rowlive ? __ldg(p.lm_norm + t.base_pos + row) : 0.0f;
                    // 8 accumulator columns of every block per load: one TMEM round trip per 8 queries (col0 + the
                    // live count rounded up to 8 stays inside the block: both halves hold <= t.nq / 2 columns)
                    for (uint32_t c8 = 0; c8 < nlive; c8 += 8) {
                        uint32_t d0[8], d1[8], d2[8];
                        tc::tmem_ld_8_nowait(tacc + col0 + c8, d0);
                        if (TWO_B) tc::tmem_ld_8_nowait(tacc + t.nq + col0 + c8, d1);
                        if (SPLIT3) tc::tmem_ld_8_nowait(tacc + 2 * t.nq + col0 + c8, d2);
                        if (SPLIT3)
                            tc::tmem_ld_wait_24(d0, d1, d2);
                        else if (TWO_B)
                            tc::tmem_ld_wait_16(d0, d1);
                        else
                            tc::tmem_ld_wait_8(d0);
#pragma unroll
                        for (uint32_t jj = 0; jj < 8; ++jj) {
                            if (c8 + jj < nlive) {  // warp-uniform
                                float dot = __uint_as_float(d0[jj]);
                                if (SPLIT3) dot = __fadd_rn(dot, __fadd_rn(__uint_as_float(d1[jj]), __uint_as_float(d2[jj])));
                                if (H16) dot = __fmaf_rn(__uint_as_float(d1[jj]), H16_LO_INV, dot);
                                if (rowlive)
                                    p.dense_out[(t.q0 + col0 + c8 + jj) * p.dense_ld + t.base_pos + row] =
                                        __fmaf_rn(H16 ? p.key_scale : -2.0f, dot, nx);
                            }
                        }
                    }
                    tc::fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&tempty[buf]);
                    ++tile_ctr;
                }
                continue;
            }
            // lane j keeps query j's threshold (its list's 32nd key, or the query's shared bound) and queue fill
            float my_tau = __int_as_float(0x7f800000);
            float my_pub = __int_as_float(0x7f800000);  // what this list has published so far
            uint32_t my_cnt = 0, my_q = 0;
            uint64_t my_base = 0;  // where query `lane`'s partial list of this (item, lane group) goes: fetched now, so
                                   // the two dependent loads are long done when the item ends
            if ((uint32_t)lane < nlive) {
                const uint32_t pair = p.lq_pair[t.q0 + col0 + lane];
                my_base = ((p.pair_chunk_off[pair] + t.chunk) * TC_PARTS + lane_group) * 32;
                if (p.qtau) {
                    my_q = p.lq_query[t.q0 + col0 + lane];
                    my_tau = tau_decode(tau_shared_bits(p.qtau, my_q));
                }
            }
            for (uint32_t j = 0; j < nlive; ++j) {
                lk[j * 32 + lane] = __int_as_float(0x7f800000);
                lr[j * 32 + lane] = 0xffffu;
            }
            __syncwarp();
            const uint32_t c_hh = col0, c_hl = t.nq + col0, c_lh = 2 * t.nq + col0;
            // ||x||^2 of the tile's row is fetched one tile ahead, the query's shared bound before the wait for the
            // accumulator: neither global load sits on the epilogue's critical path
            const uint64_t row_first = t.r0 + (uint64_t)(lane_group * 32 + lane);
            float nx_next = row_first < t.r1 ? __ldg(p.lm_norm + t.base_pos + row_first) : 0.0f;
            for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                const uint32_t buf = tile_ctr & 1, tphase = (tile_ctr >> 1) & 1;
                uint32_t tau_bits = TAU_INF;
                if (p.qtau && (uint32_t)lane < nlive && a0 != t.r0)  // bounds published by other CTAs meanwhile
                    tau_bits = tau_shared_bits(p.qtau, my_q);
                tc::mbar_wait(&tfull[buf], tphase);
                tc::fence_after_thread_sync();
                const uint32_t tacc = tmem_base + ((uint32_t)(lane_group * 32) << 16) + buf * Cfg::ACC_COLS;
                const uint64_t row = a0 + (uint64_t)(lane_group * 32 + lane);
                const bool rowlive = row < t.r1;
                const uint32_t roff = (uint32_t)(row - t.r0);
                const float nx = nx_next;
                if (row + TC_M < t.r1) nx_next = __ldg(p.lm_norm + t.base_pos + row + TC_M);
                my_tau = fminf(my_tau, tau_decode(tau_bits));
                // one (row, query j) key: rows that beat the query's threshold are appended to its queue
                auto consider = [&](uint32_t j, float key) {
                    float tau = __shfl_sync(FULL_MASK, my_tau, j);
                    bool pass = rowlive && key <= tau;
                    unsigned m = __ballot_sync(FULL_MASK, pass);
                    if (m) {
                        uint32_t c = __shfl_sync(FULL_MASK, my_cnt, j);
                        uint32_t n = __popc(m);
                        if (c + n > TC_QCAP) {  // make room: fold the queue into the list, which tightens tau
                            const float2 f = sel_flush(lk + j * 32, lr + j * 32, qk + j * TC_QCAP, qr + j * TC_QCAP, c, lane);
                            tau = fminf(tau, f.x);
                            if ((uint32_t)lane == j) {
                                if (p.qtau && f.y < my_pub) {
                                    atomicMin(p.qtau + 4 * my_q + lane_group, tau_encode(f.y));
                                    my_pub = f.y;
                                }
                                my_tau = tau;
                            }
                            c = 0;
                            pass = pass && key <= tau;
                            m = __ballot_sync(FULL_MASK, pass);
                            n = __popc(m);
                        }
                        if (pass) {
                            const uint32_t o = c + __popc(m & lt_mask);
                            qk[j * TC_QCAP + o] = key;
                            qr[j * TC_QCAP + o] = (uint16_t)roff;
                        }
                        if ((uint32_t)lane == j) my_cnt = c + n;
                    }
                };
                if (H16) {
                    // 8 accumulator columns of both blocks per load: one TMEM round trip per 8 queries
                    for (uint32_t c8 = 0; c8 < nlive; c8 += 8) {
                        uint32_t hh[8], hl[8];
                        tc::tmem_ld_8_nowait(tacc + c_hh + c8, hh);
                        tc::tmem_ld_8_nowait(tacc + c_hl + c8, hl);
                        tc::tmem_ld_wait_16(hh, hl);
#pragma unroll
                        for (uint32_t jj = 0; jj < 8; ++jj) {
                            if (c8 + jj < nlive) {  // warp-uniform
                                const float dot = __fmaf_rn(__uint_as_float(hl[jj]), H16_LO_INV, __uint_as_float(hh[jj]));
                                consider(c8 + jj, __fmaf_rn(p.key_scale, dot, nx));
                            }
                        }
                    }
                } else {
                    for (uint32_t c8 = 0; c8 < nlive; c8 += 8) {
                        uint32_t hh[8], hl[8], lh[8];
                        tc::tmem_ld_8_nowait(tacc + c_hh + c8, hh);
                        if (SPLIT3) {
                            tc::tmem_ld_8_nowait(tacc + c_hl + c8, hl);  // x_hi . q_lo
                            tc::tmem_ld_8_nowait(tacc + c_lh + c8, lh);  // x_lo . q_hi
                            tc::tmem_ld_wait_24(hh, hl, lh);
                        } else {
                            tc::tmem_ld_wait_8(hh);
                        }
#pragma unroll
                        for (uint32_t jj = 0; jj < 8; ++jj) {
                            if (c8 + jj < nlive) {  // warp-uniform
                                float dot = __uint_as_float(hh[jj]);
                                if (SPLIT3) dot = __fadd_rn(dot, __fadd_rn(__uint_as_float(hl[jj]), __uint_as_float(lh[jj])));
                                consider(c8 + jj, __fmaf_rn(-2.0f, dot, nx));
                            }
                        }
                    }
                }
                tc::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[buf]);
                ++tile_ctr;
            }
            const uint32_t pos0 = (uint32_t)(t.base_pos + t.r0);
            for (uint32_t j = 0; j < nlive; ++j) {
                const uint32_t c = __shfl_sync(FULL_MASK, my_cnt, j);
                if (c) {
                    const float2 f = sel_flush(lk + j * 32, lr + j * 32, qk + j * TC_QCAP, qr + j * TC_QCAP, c, lane);
                    if (p.qtau && (uint32_t)lane == j && f.y < my_pub) atomicMin(p.qtau + 4 * my_q + lane_group, tau_encode(f.y));
                }
                const uint64_t base = __shfl_sync(FULL_MASK, my_base, j);
                const uint32_t r = lr[j * 32 + lane];
                p.part_d[base + lane] = lk[j * 32 + lane];
                p.part_p[base + lane] = r == 0xffffu ? 0xffffffffu : pos0 + r;
            }
            __syncwarp();
        }
    } else if (SPLIT3 && warp >= TC_CONV_WARP0) {
        // ===================== converters: x_lo = x - trunc_tf32(x) of the landed tile -> tensor memory ==========
        const int lane_group = warp & 3;
        const uint32_t row = (uint32_t)(lane_group * 32 + lane);  // tile row == TMEM lane of this thread
        uint32_t stage = 0, phase = 0;
        while (true) {
            const TcItem t = next_item(lane == 0, true);
            if (t.list == TC_ITEM_END) break;
            for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                for (uint32_t kc = 0; kc < nk; ++kc) {
                    tc::mbar_wait(&full[stage], phase);
                    const uint8_t* arow = smem + stage * Cfg::STAGE_BYTES + row * 128;
                    uint32_t lo[TC_KC];
#pragma unroll
                    for (uint32_t c = 0; c < 8; ++c) {  // 128-byte swizzle: logical 16-byte chunk c sits at c ^ (row & 7)
                        const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7u)) << 4));
                        lo[4 * c + 0] = __float_as_uint(__fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u)));
                        lo[4 * c + 1] = __float_as_uint(__fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u)));
                        lo[4 * c + 2] = __float_as_uint(__fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u)));
                        lo[4 * c + 3] = __float_as_uint(__fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u)));
                    }
                    tc::tmem_st_32(tmem_base + ((uint32_t)(lane_group * 32) << 16) + Cfg::ALO_COL0 + stage * TC_KC, lo);
                    tc::fence_before_thread_sync();
                    __syncwarp();
                    tc::mbar_arrive(&conv[stage]);  // per lane: this lane's reads of the stage are done
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace vers
