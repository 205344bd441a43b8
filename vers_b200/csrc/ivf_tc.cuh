// ivf_tc.cuh — tensor-core candidate pass of the inverted-list scan (included by ivf.cu).
//
// The dot products x.q of the candidate pass move from the fp32 SIMT pipe to the 5th-generation tensor cores:
//   TMA (cp.async.bulk.tensor, 128-byte swizzle) streams 128-row x 32-float tiles of a list straight from the fp32
//   list-major rows into shared memory (no bf16 copy, no register staging); tcgen05.mma kind::tf32 (M=128, N=16, K=8)
//   multiplies them with the group's <= 16 queries, accumulating in TMEM (double buffered); the epilogue warps pull
//   the 128x16 tile with tcgen05.ld, form the candidate key ||x||^2 - 2 x.q and keep a private top-32 per query in
//   registers.  What leaves the kernel is the same (key, position) partial lists as the SIMT candidate pass, so the
//   merge -> exact-order rerank -> certificate -> exact redo chain behind it is unchanged.  The kernel does no fp32
//   SIMT math per (row, query, dim): it is a pure HBM stream.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (warp w reads TMEM lanes 32*(w%4)..+31 = tile rows).  mbarriers: full/empty per smem stage, tmem_full/tmem_empty per
// accumulator buffer.  Work items (list, 16-query group, 4096-row chunk) are dealt round-robin to the CTAs.
#pragma once
#include "tc.cuh"

namespace vers {

constexpr int TC_M = 128, TC_N = 16, TC_KC = 32, TC_STAGES = 6, TC_THREADS = 192, TC_EPI_WARPS = 4;
constexpr int TC_A_BYTES = TC_M * TC_KC * 4, TC_B_BYTES = TC_N * TC_KC * 4, TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_SMEM_BYTES = 1024 + TC_STAGES * TC_STAGE_BYTES + 256;
constexpr uint32_t TC_TMEM_COLS = 32;  // two 16-column accumulator buffers

__global__ void gather_queries_kernel(const float* __restrict__ queries, const uint32_t* __restrict__ lq_query,
                                      const uint64_t* __restrict__ lq_off, uint32_t C, uint32_t ld,
                                      float* __restrict__ gq) {
    const uint64_t n = lq_off[C];
    const uint32_t ld4 = ld >> 2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * ld4;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / ld4;
        uint32_t c = (uint32_t)(i - r * ld4);
        reinterpret_cast<float4*>(gq)[i] = reinterpret_cast<const float4*>(queries)[(uint64_t)lq_query[r] * ld4 + c];
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

__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_list_scan_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_q,
                        TcScanParams p) {
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t raw = tc::smem_u32(tc_smem_raw);
    uint8_t* smem = tc_smem_raw + (((raw + 1023u) & ~1023u) - raw);  // SWIZZLE_128B tiles need 1024-byte alignment
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tfull = empty + TC_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t total_items = p.item_off[p.C];
    const uint32_t nk = (p.ld + TC_KC - 1) / TC_KC;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tfull[b], 1);
            tc::mbar_init(&tempty[b], TC_EPI_WARPS);
        }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmap_rows);
        tc::tma_prefetch_desc(&tmap_q);
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint64_t it = blockIdx.x; it < total_items; it += gridDim.x) {
                const TcItem t = tc_decode_item(p, it);
                for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                    for (uint32_t kc = 0; kc < nk; ++kc) {
                        tc::mbar_wait(&empty[stage], phase ^ 1);
                        tc::mbar_arrive_expect_tx(&full[stage], TC_STAGE_BYTES);
                        uint8_t* sa = smem + stage * TC_STAGE_BYTES;
                        tc::tma_load_2d(sa, &tmap_rows, &full[stage], (int32_t)(kc * TC_KC), (int32_t)(t.base_pos + a0));
                        tc::tma_load_2d(sa + TC_A_BYTES, &tmap_q, &full[stage], (int32_t)(kc * TC_KC), (int32_t)t.q0);
                        if (++stage == TC_STAGES) {
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
            const uint32_t idesc = tc::idesc_tf32(TC_M, TC_N);
            uint32_t stage = 0, phase = 0, tile_ctr = 0;
            for (uint64_t it = blockIdx.x; it < total_items; it += gridDim.x) {
                const TcItem t = tc_decode_item(p, it);
                for (uint64_t a0 = t.r0; a0 < t.r1; a0 += TC_M) {
                    const uint32_t buf = tile_ctr & 1, tphase = (tile_ctr >> 1) & 1;
                    tc::mbar_wait(&tempty[buf], tphase ^ 1);  // epilogue has drained this accumulator buffer
                    tc::fence_after_thread_sync();
                    const uint32_t d_tmem = tmem_base + buf * TC_N;
                    for (uint32_t kc = 0; kc < nk; ++kc) {
                        tc::mbar_wait(&full[stage], phase);
                        tc::fence_after_thread_sync();
                        const uint32_t sa = tc::smem_u32(smem + stage * TC_STAGE_BYTES);
                        const uint64_t da = tc::smem_desc_k_sw128(sa), db = tc::smem_desc_k_sw128(sa + TC_A_BYTES);
#pragma unroll
                        for (uint32_t kk = 0; kk < TC_KC / 8; ++kk)
                            tc::mma_tf32(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (kc | kk) != 0);
                        tc::mma_commit(&empty[stage]);  // frees the smem stage once these MMAs have read it
                        if (++stage == TC_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    tc::mma_commit(&tfull[buf]);  // accumulator complete
                    ++tile_ctr;
                }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> keys -> private top-32 per query =====================
        const int lane_group = warp & 3;  // TMEM lanes this warp may touch
        const int epi = warp - 2;         // partial-list slot of this warp
        uint32_t tile_ctr = 0;
        for (uint64_t it = blockIdx.x; it < total_items; it += gridDim.x) {
            const TcItem t = tc_decode_item(p, it);
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
                tc::tmem_ld_16(tmem_base + ((uint32_t)(lane_group * 32) << 16) + buf * TC_N, v);
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
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

}  // namespace vers
