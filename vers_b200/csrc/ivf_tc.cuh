// ivf_tc.cuh — tensor-core candidate pass of the inverted-list scan (included by ivf.cu).
//
// The dot products x.q of the candidate pass run on the 5th-generation tensor cores instead of the fp32 SIMT pipe:
//   * TMA (cp.async.bulk.tensor, 128-byte swizzle) streams 128-row x 32-float tiles of a list straight from the fp32
//     list-major rows into shared memory — no 16-bit copy of the database, no register staging;
//   * tcgen05.mma kind::tf32 (M=128, N=16, K=8) multiplies them with the group's <= 16 queries, accumulators in TMEM
//     (two 16-column buffers, so the MMAs of tile t+1 overlap the epilogue of tile t);
//   * SPLIT3 (default): TF32 keeps 10 mantissa bits, too coarse for the rounding-error certificate to separate real
//     neighbours, so every operand is split x = hi + lo (hi = x with the low 13 mantissa bits cleared, lo = x - hi,
//     exact in fp32) and three MMAs are issued per K-step: hi.hi + lo.hi + hi.lo (the dropped lo.lo term and the
//     truncation of lo are <= 3 * 2^-20 |x||q| per product).  The tensor core itself truncates fp32 -> tf32 (measured:
//     bit-identical results with and without clearing the low bits first), so the row tile as loaded IS x_hi; four
//     converter warps compute x_lo from the landed tile and park it in TENSOR MEMORY (tcgen05.st), from where the
//     lo.hi MMA takes its A operand — shared memory sees each row byte only three times (TMA write, converter
//     read, MMA read).  hi.hi and hi.lo are ONE MMA against the concatenated [q_hi; q_lo] tile (N = 32); the two
//     16-column halves are summed in the epilogue.  The queries are split once by the gather kernel;
//   * the epilogue warps pull the 128x16 tile with tcgen05.ld, form the key ||x||^2 - 2 x.q and keep a private
//     top-32 per query in registers.
// What leaves the kernel is the same (key, position) partial lists as the SIMT candidate pass, so the merge ->
// exact-order rerank -> certificate -> exact redo chain behind it is unchanged.  The kernel does no fp32 SIMT math
// per (row, query, dim): it is an HBM stream.
//
// Warp roles: 0 = scheduler + TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue (warp w reads TMEM
// lanes 32*(w%4)..+31 = tile rows), 6..9 = hi/lo converters (SPLIT3 only).  mbarriers: full / conv / empty per smem
// stage, tmem_full / tmem_empty per accumulator buffer, sched_full / sched_empty for the work-item ring (work items
// = (list, 16-query group, 4096-row chunk), handed out dynamically through an atomic counter).
#pragma once
#include "tc.cuh"

namespace vers {

constexpr int TC_M = 128, TC_N = 16, TC_KC = 32, TC_EPI_WARPS = 4, TC_CONV_WARPS = 4, TC_CONV_GROUPS = 1, TC_SCHED = 4;
constexpr int TC_A_BYTES = TC_M * TC_KC * 4, TC_B_BYTES = TC_N * TC_KC * 4;

template <bool SPLIT3>
struct TcCfg {
    static constexpr int STAGES = 10;  // ~200 KB of row tiles in flight per SM: covers HBM latency + convert + MMA
    static constexpr int THREADS = SPLIT3 ? (6 + TC_CONV_WARPS * TC_CONV_GROUPS + 1) * 32 : 192;  // + lo-MMA issuer warp
    static constexpr int LO_WARP = 6 + TC_CONV_WARPS * TC_CONV_GROUPS;
    static constexpr int NB = SPLIT3 ? 2 * TC_N : TC_N;         // B rows per stage: [q_hi; q_lo] or q
    static constexpr int ACC_COLS = SPLIT3 ? 3 * TC_N : TC_N;    // per buffer: [hi.hi | hi.lo | lo.hi] or [x.q]
    static constexpr int STAGE_BYTES = TC_A_BYTES + NB * TC_KC * 4;
    static constexpr int TX_BYTES = STAGE_BYTES;
    static constexpr int OFF_B = TC_A_BYTES;
    static constexpr int ALO_COL0 = 2 * ACC_COLS;                // x_lo tiles live in TMEM after the accumulators
    static constexpr uint32_t TMEM_COLS = SPLIT3 ? 512 : 32;     // 96 + 10*32 = 416 -> 512 / 2*16 = 32
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 512;
};

// queries regrouped by list (row i = the i-th grouped (query, list) pair), split into tf32 hi and lo parts
__global__ void gather_queries_kernel(const float* __restrict__ queries, const uint32_t* __restrict__ lq_query,
                                      const uint64_t* __restrict__ lq_off, uint32_t C, uint32_t ld,
                                      float* __restrict__ gq_hi, float* __restrict__ gq_lo, int split) {
    const uint64_t n = lq_off[C];
    const uint32_t ld4 = ld >> 2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * ld4;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / ld4;
        uint32_t c = (uint32_t)(i - r * ld4);
        float4 v = reinterpret_cast<const float4*>(queries)[(uint64_t)lq_query[r] * ld4 + c];
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

struct TcScanParams {
    uint32_t ld, C;
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
};

struct TcItem {
    uint32_t list, chunk;
    uint64_t q0, nB, base_pos, r0, r1;
};

__device__ __forceinline__ TcItem tc_decode_item(const TcScanParams& p, uint64_t it) {
    uint32_t lo = 0, hi = p.C;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (p.item_off[mid] <= it) lo = mid; else hi = mid;
    }
    TcItem t;
    t.list = lo;
    const uint64_t local = it - p.item_off[lo];
    const uint32_t len = p.seg_len[lo];
    const uint32_t nch = (len + LIST_CHUNK_ROWS - 1) / LIST_CHUNK_ROWS;
    const uint32_t group = (uint32_t)(local / nch);
    t.chunk = (uint32_t)(local % nch);
    t.q0 = p.lq_off[lo] + (uint64_t)group * TC_N;
    const uint64_t m_l = p.lq_off[lo + 1] - p.lq_off[lo];
    t.nB = min((uint64_t)TC_N, m_l - (uint64_t)group * TC_N);
    t.base_pos = p.seg_off[lo];
    t.r0 = (uint64_t)t.chunk * LIST_CHUNK_ROWS;
    t.r1 = min((uint64_t)len, t.r0 + LIST_CHUNK_ROWS);
    return t;
}

template <bool SPLIT3>
__global__ void __launch_bounds__(TcCfg<SPLIT3>::THREADS, 1)
    tc_list_scan_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_qhi,
                        const __grid_constant__ CUtensorMap tmap_qlo, TcScanParams p) {
    using Cfg = TcCfg<SPLIT3>;
    constexpr int S = Cfg::STAGES;
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t raw = tc::smem_u32(tc_smem_raw);
    uint8_t* smem = tc_smem_raw + (((raw + 1023u) & ~1023u) - raw);  // SWIZZLE_128B tiles need 1024-byte alignment
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES);
    uint64_t* conv = full + S;
    uint64_t* empty = conv + S;
    uint64_t* tfull = empty + S;
    uint64_t* tempty = tfull + 2;
    uint64_t* sfull = tempty + 2;
    uint64_t* sempty = sfull + TC_SCHED;
    long long* sched = reinterpret_cast<long long*>(sempty + TC_SCHED);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched + TC_SCHED);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t total_items = p.item_off[p.C];
    const uint32_t nk = (p.ld + TC_KC - 1) / TC_KC;
    // consumers of a scheduled item besides the producer: MMA thread + epilogue warps (+ converter warps)
    constexpr uint32_t SCHED_CONSUMERS = 1 + TC_EPI_WARPS + (SPLIT3 ? TC_CONV_WARPS * TC_CONV_GROUPS + 1 : 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&conv[s], TC_CONV_WARPS);
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
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmap_rows);
        tc::tma_prefetch_desc(&tmap_qhi);
        if (SPLIT3) tc::tma_prefetch_desc(&tmap_qlo);
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    // every consumer role walks the same ring of scheduled items
    uint32_t sslot = 0, sphase = 0;
    auto next_item = [&](bool leader, bool whole_warp) -> long long {
        tc::mbar_wait(&sfull[sslot], sphase);
        long long it = sched[sslot];
        if (whole_warp) __syncwarp();  // every lane has read the slot before the leader releases it
        if (leader) tc::mbar_arrive(&sempty[sslot]);
        if (++sslot == TC_SCHED) {
            sslot = 0;
            sphase ^= 1;
        }
        return it;
    };

    if (warp == 0) {
        // ===================== scheduler + TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            while (true) {
                tc::mbar_wait(&sempty[sslot], sphase ^ 1);
                unsigned long long itu = atomicAdd(p.counter, 1ull);
                const long long it = itu < total_items ? (long long)itu : -1;
                sched[sslot] = it;
                tc::mbar_arrive(&sfull[sslot]);  // release: the slot value is visible to the waiters
                if (++sslot == TC_SCHED) {
                    sslot = 0;
                    sphase ^= 1;
                }
                if (it < 0) break;
                const TcItem t = tc_decode_item(p, (uint64_t)it);
                for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                    for (uint32_t kc = 0; kc < nk; ++kc) {
                        tc::mbar_wait(&empty[stage], phase ^ 1);
                        tc::mbar_arrive_expect_tx(&full[stage], Cfg::TX_BYTES);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        tc::tma_load_2d(sa, &tmap_rows, &full[stage], (int32_t)(kc * TC_KC), (int32_t)(t.base_pos + a0));
                        tc::tma_load_2d(sa + Cfg::OFF_B, &tmap_qhi, &full[stage], (int32_t)(kc * TC_KC), (int32_t)t.q0);
                        if (SPLIT3)
                            tc::tma_load_2d(sa + Cfg::OFF_B + TC_B_BYTES, &tmap_qlo, &full[stage], (int32_t)(kc * TC_KC),
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
            const uint32_t idesc_main = tc::idesc_tf32(TC_M, Cfg::NB);  // x_hi . [q_hi; q_lo]  (or x . q unsplit)
            uint32_t stage = 0, phase = 0, tile_ctr = 0;
            while (true) {
                const long long it = next_item(true, false);
                if (it < 0) break;
                const TcItem t = tc_decode_item(p, (uint64_t)it);
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
                        for (uint32_t kk = 0; kk < TC_KC / 8; ++kk)
                            tc::mma_tf32(d_tmem, da + 2 * kk, db + 2 * kk, idesc_main, (kc | kk) != 0);
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
    } else if (warp < 2 + TC_EPI_WARPS) {
        // ===================== epilogue: TMEM -> keys -> private top-32 per query =====================
        const int lane_group = warp & 3;  // TMEM lanes this warp may touch
        const int epi = warp - 2;         // partial-list slot of this warp
        uint32_t tile_ctr = 0;
        while (true) {
            const long long it = next_item(lane == 0, true);
            if (it < 0) break;
            const TcItem t = tc_decode_item(p, (uint64_t)it);
            float rl_d[TC_N], tau_d[TC_N];
            uint32_t rl_p[TC_N], tau_p[TC_N];
#pragma unroll
            for (int j = 0; j < TC_N; ++j) {
                rl_d[j] = tau_d[j] = __int_as_float(0x7f800000);
                rl_p[j] = tau_p[j] = 0xffffffffu;
            }
            for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                const uint32_t buf = tile_ctr & 1, tphase = (tile_ctr >> 1) & 1;
                tc::mbar_wait(&tfull[buf], tphase);
                tc::fence_after_thread_sync();
                float v[TC_N];
                const uint32_t tacc = tmem_base + ((uint32_t)(lane_group * 32) << 16) + buf * Cfg::ACC_COLS;
                tc::tmem_ld_16(tacc, v);
                if (SPLIT3) {
                    float w[TC_N], z[TC_N];
                    tc::tmem_ld_16(tacc + TC_N, w);      // x_hi . q_lo
                    tc::tmem_ld_16(tacc + 2 * TC_N, z);  // x_lo . q_hi
#pragma unroll
                    for (int j = 0; j < TC_N; ++j) v[j] = __fadd_rn(v[j], __fadd_rn(w[j], z[j]));
                }
                tc::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[buf]);
                ++tile_ctr;
                const uint64_t row = a0 + (uint64_t)(lane_group * 32 + lane);
                const bool rowlive = row < t.r1;
                const uint32_t pos = (uint32_t)(t.base_pos + row);
                const float nx = rowlive ? __ldg(p.lm_norm + pos) : 0.0f;
#pragma unroll
                for (int j = 0; j < TC_N; ++j) {
                    const float key = __fmaf_rn(-2.0f, v[j], nx);
                    bool live = rowlive && (uint64_t)j < t.nB;
                    while (true) {
                        bool pass = live && entry_less<uint32_t>(key, pos, tau_d[j], tau_p[j]);
                        unsigned m = __ballot_sync(FULL_MASK, pass);
                        if (!m) break;
                        int src = __ffs(m) - 1;
                        float cv = __shfl_sync(FULL_MASK, key, src);
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
#pragma unroll
            for (int j = 0; j < TC_N; ++j) {
                if ((uint64_t)j < t.nB) {
                    const uint32_t pair = p.lq_pair[t.q0 + j];
                    const uint64_t base = ((p.pair_chunk_off[pair] + t.chunk) * TC_EPI_WARPS + epi) * 32;
                    p.part_d[base + lane] = rl_d[j];
                    p.part_p[base + lane] = rl_p[j];
                }
            }
        }
    } else if (SPLIT3 && warp == Cfg::LO_WARP) {
        // ===================== second MMA issuer: x_lo (tensor memory) . q_hi -> its own 16 accumulator columns ====
        // (a separate thread so that neither issuer's instruction latency per stage exceeds the HBM time per stage)
        if (lane == 0) {
            const uint32_t idesc_lo = tc::idesc_tf32(TC_M, TC_N);
            uint32_t stage = 0, phase = 0, tile_ctr = 0;
            while (true) {
                const long long it = next_item(true, false);
                if (it < 0) break;
                const TcItem t = tc_decode_item(p, (uint64_t)it);
                for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                    const uint32_t buf = tile_ctr & 1, tphase = (tile_ctr >> 1) & 1;
                    tc::mbar_wait(&tempty[buf], tphase ^ 1);
                    tc::fence_after_thread_sync();
                    const uint32_t d_tmem = tmem_base + buf * Cfg::ACC_COLS + 2 * TC_N;
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
    } else if (SPLIT3) {
        // ===================== converters: x_lo = x - trunc_tf32(x) of the landed tile -> tensor memory ==========
        // TC_CONV_GROUPS groups of 4 warps take alternate stage uses, so two tiles are being split at any time
        const int lane_group = warp & 3;
        const uint32_t cgroup = (uint32_t)(warp - (2 + TC_EPI_WARPS)) / TC_CONV_WARPS;
        const uint32_t row = (uint32_t)(lane_group * 32 + lane);  // tile row == TMEM lane of this thread
        uint32_t stage = 0, phase = 0, use = 0;
        while (true) {
            const long long it = next_item(lane == 0, true);
            if (it < 0) break;
            const TcItem t = tc_decode_item(p, (uint64_t)it);
            for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                for (uint32_t kc = 0; kc < nk; ++kc, ++use) {
                    if (use % TC_CONV_GROUPS != cgroup) {
                        if (++stage == S) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
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
                    if (lane == 0) tc::mbar_arrive(&conv[stage]);
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
