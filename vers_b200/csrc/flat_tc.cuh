// flat_tc.cuh — exhaustive search of a LARGE query batch (BASELINE configs[1]: 1M x 300, 1000 queries), candidate pass
// on the tensor cores with the table streamed ONCE per 128 queries (included by ivf.cu).
//
// The list-scan kernel (ivf_tc.cuh) keeps the ROWS on the M side of the MMA and at most 32 queries on the N side, so a
// table that every query scans is re-streamed once per 32 queries (38 GB moved for a 1.2 GB table at 1000 queries).
// Here the roles are swapped, like tc_assign1_kernel: a block of 128 QUERIES is converted once (rounded to nearest
// tf32) and stays resident in TENSOR MEMORY as the A operand (ld <= 320 columns next to two 64-column accumulators);
// the table rows are the B operand, streamed in 64-row tiles from a TILE-MAJOR IMAGE of the table (built once per
// dataset by tile_image_tf32_kernel: rows rounded to nearest tf32, laid out exactly as the swizzled shared-memory
// tile), so a tile is one linear cp.async.bulk of nk x 8 KB; ONE tcgen05.mma kind::tf32 per K step.  Work items = (table slice, query block), ordered so that the query
// blocks of one slice run on neighbouring CTAs at the same time: the slice comes from HBM once and from L2 for the
// other blocks.
// Epilogue: thread = query.  key = ||x||^2 - 2 x.q per (query, row); the 32 smallest (key, row) of the item are kept
// in a sorted per-thread list in shared memory (transposed: entry e of thread t at [e][t], conflict-free).  A
// hierarchical min test (tile -> 16 -> 4 -> 1 columns) against min(own 32nd key, the query's shared bound qtau) keeps
// the insertion code off the common path; qtau is shared between all items of the launch through global memory exactly
// like in the list scan (a published value is the 32nd key of a list that already holds 32 rows, so a row above it
// cannot be among the query's 32 best keys — DESIGN.md).
// What leaves the kernel is the same sorted 32-entry partial lists as the list scan's, so merge -> exact-order rerank
// -> certificate (TF32 error model) -> exact redo of uncertified queries behind it are unchanged.
#pragma once
#include "tc.cuh"

namespace vers {

constexpr int FT_M = 128, FT_N = 64, FT_KC = 32, FT_MAX_KCH = 10;
constexpr int FT_EPI_WARP0 = 2, FT_EPI_WARPS = 4, FT_CONV_WARP0 = 6, FT_CONV_WARPS = 4;
constexpr int FT_THREADS = (2 + FT_EPI_WARPS + FT_CONV_WARPS) * 32;
constexpr int FT_BOX_BYTES = FT_N * FT_KC * 4;    // one [64 rows x 32 floats] box
constexpr int FT_ASLOT_BYTES = FT_M * FT_KC * 4;  // one [128 queries x 32 floats] staging slot
constexpr int FT_LIST = 32, FT_QCAP = 16;  // sorted list + pending queue per query (entries)
constexpr uint32_t FT_ACC_COL0 = 0, FT_A_COL0 = 128, FT_TMEM_COLS = 512;

struct TcFlatParams {
    uint64_t n_rows;
    uint32_t nq, ld, nk;   // nk = K chunks of 32 floats (<= FT_MAX_KCH)
    uint32_t nqb;          // query blocks of 128
    uint32_t nslices;
    uint32_t slice_rows;   // multiple of 64
    uint32_t stages;       // B ring depth
    const float* row_tiles;  // tile_image_tf32_kernel's image of the table: [ceil(n/64)][nk][64][32] swizzled, rounded
    const float* row_norm;  // [round_up(n, 64)] ||x||^2 (any order), +inf past n
    float* part_d;          // [nq][nslices][32]
    uint32_t* part_p;
    uint32_t* qtau;         // [nq][4] order-preserving encoding of the query's shared bound slots (TAU_INF at launch)
};

struct FtSmem {
    uint8_t* stages;
    uint8_t* aslot;
    float *lkey, *qkey;
    uint32_t *lpos, *qpos;
    float* nrm;
    uint64_t *full, *empty, *afull, *aempty, *aready, *afree, *tfull, *tempty, *nbar;
    uint32_t* tmem_slot;
};

__device__ __forceinline__ bool ft_elect_one() {
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

inline size_t ft_smem_bytes(uint32_t nk, uint32_t stages) {
    return 1024 + (size_t)stages * nk * FT_BOX_BYTES + FT_ASLOT_BYTES + (size_t)(FT_LIST + FT_QCAP) * FT_M * 8 +
           4 * FT_N * 4 + 512;
}

__global__ void __launch_bounds__(FT_THREADS, 1)
    tc_flat_kernel(const __grid_constant__ CUtensorMap tmap_q, TcFlatParams p) {
    extern __shared__ uint8_t ft_smem_raw[];
    const uint32_t raw = tc::smem_u32(ft_smem_raw);
    uint8_t* smem = ft_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const uint32_t stage_bytes = p.nk * FT_BOX_BYTES;
    FtSmem sm;
    sm.stages = smem;
    sm.aslot = sm.stages + (size_t)p.stages * stage_bytes;
    sm.lkey = reinterpret_cast<float*>(sm.aslot + FT_ASLOT_BYTES);
    sm.lpos = reinterpret_cast<uint32_t*>(sm.lkey + FT_LIST * FT_M);
    sm.qkey = reinterpret_cast<float*>(sm.lpos + FT_LIST * FT_M);
    sm.qpos = reinterpret_cast<uint32_t*>(sm.qkey + FT_QCAP * FT_M);
    sm.nrm = reinterpret_cast<float*>(sm.qpos + FT_QCAP * FT_M);
    sm.full = reinterpret_cast<uint64_t*>(sm.nrm + 4 * FT_N);
    sm.empty = sm.full + 4;
    sm.afull = sm.empty + 4;
    sm.aempty = sm.afull + 1;
    sm.aready = sm.aempty + 1;
    sm.afree = sm.aready + 1;
    sm.tfull = sm.afree + 1;
    sm.tempty = sm.tfull + 2;
    sm.nbar = sm.tempty + 2;
    sm.tmem_slot = reinterpret_cast<uint32_t*>(sm.nbar + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nk = p.nk;
    const uint64_t nitems = (uint64_t)p.nslices * p.nqb;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) {
            tc::mbar_init(&sm.full[s], 1);
            tc::mbar_init(&sm.empty[s], 1);
        }
        tc::mbar_init(sm.afull, 1);
        tc::mbar_init(sm.aempty, FT_CONV_WARPS * 32);  // every converter lane arrives itself
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&sm.tfull[s], 1);
            tc::mbar_init(&sm.tempty[s], FT_EPI_WARPS * 32);  // every epilogue lane arrives itself
        }
        tc::mbar_init(sm.aready, FT_CONV_WARPS);
        tc::mbar_init(sm.afree, 1);
        for (int b = 0; b < 4; ++b) tc::mbar_init(&sm.nbar[b], 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmap_q);
    }
    if (warp == 1) tc::tmem_alloc(sm.tmem_slot, FT_TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *sm.tmem_slot;

    // rows [r0, r1) of item `it`; its query block
    auto item_rows = [&](uint64_t it, uint64_t& r0, uint64_t& r1, uint32_t& qb) {
        const uint64_t slice = it / p.nqb;
        qb = (uint32_t)(it - slice * p.nqb);
        r0 = slice * p.slice_rows;
        r1 = min(p.n_rows, r0 + p.slice_rows);
    };

    if (warp == 0) {
        // TMA producer (converged warp, one elected lane issues)
        uint32_t stage = 0, phase = 0, chunk = 0;
        for (uint64_t it = blockIdx.x; it < nitems; it += gridDim.x) {
            uint64_t r0, r1;
            uint32_t qb;
            item_rows(it, r0, r1, qb);
            for (uint32_t kc = 0; kc < nk; ++kc, ++chunk) {  // one staging slot: the queries of an item load once
                tc::mbar_wait(sm.aempty, (chunk & 1) ^ 1);
                if (ft_elect_one()) {
                    tc::mbar_arrive_expect_tx(sm.afull, FT_ASLOT_BYTES);
                    tc::tma_load_2d(sm.aslot, &tmap_q, sm.afull, (int32_t)(kc * FT_KC), (int32_t)(qb * FT_M));
                }
                __syncwarp();
            }
            for (uint64_t r = r0; r < r1; r += FT_N) {
                tc::mbar_wait(&sm.empty[stage], phase ^ 1);
                if (ft_elect_one()) {  // the whole 64-row tile (all K chunks) is one linear copy of its image
                    tc::mbar_arrive_expect_tx(&sm.full[stage], nk * FT_BOX_BYTES);
                    tc::bulk_load(sm.stages + (size_t)stage * stage_bytes, p.row_tiles + (r / FT_N) * nk * (FT_BOX_BYTES / 4),
                                  nk * FT_BOX_BYTES, &sm.full[stage]);
                }
                __syncwarp();
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer (converged warp, one elected lane issues): per tile nk x 4 MMAs of M=128 N=64 K=8
        const uint32_t idesc = tc::idesc_tf32(FT_M, FT_N);
        uint32_t stage = 0, phase = 0;
        uint64_t t = 0, i = 0;
        for (uint64_t it = blockIdx.x; it < nitems; it += gridDim.x, ++i) {
            uint64_t r0, r1;
            uint32_t qb;
            item_rows(it, r0, r1, qb);
            tc::mbar_wait(sm.aready, (uint32_t)(i & 1));
            tc::fence_after_thread_sync();
            for (uint64_t r = r0; r < r1; r += FT_N, ++t) {
                const uint32_t buf = (uint32_t)(t & 1), tph = (uint32_t)((t >> 1) & 1);
                tc::mbar_wait(&sm.tempty[buf], tph ^ 1);
                tc::mbar_wait(&sm.full[stage], phase);
                tc::fence_after_thread_sync();
                if (ft_elect_one()) {
                    // the tile's ||x||^2: slot t % 4 was last read for tile t-4, which the epilogue finished before
                    // it released tile t-3 (and this tile waited for the release of tile t-2)
                    const uint32_t slot = (uint32_t)(t & 3);
                    tc::mbar_arrive_expect_tx(&sm.nbar[slot], FT_N * 4);
                    tc::bulk_load(sm.nrm + slot * FT_N, p.row_norm + r, FT_N * 4, &sm.nbar[slot]);
                    const uint32_t sb = tc::smem_u32(sm.stages + (size_t)stage * stage_bytes);
                    const uint32_t d_tmem = tmem_base + FT_ACC_COL0 + buf * FT_N;
                    const uint32_t a_tmem = tmem_base + FT_A_COL0;
                    for (uint32_t kc = 0; kc < nk; ++kc) {
                        const uint64_t db = tc::smem_desc_k_sw128(sb + kc * FT_BOX_BYTES);
#pragma unroll
                        for (uint32_t kk = 0; kk < FT_KC / 8; ++kk)
                            tc::mma_tf32_ts(d_tmem, a_tmem + kc * FT_KC + 8 * kk, db + 2 * kk, idesc, (kc | kk) != 0);
                    }
                    tc::mma_commit(&sm.tfull[buf]);
                    tc::mma_commit(&sm.empty[stage]);
                }
                __syncwarp();
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (ft_elect_one()) tc::mma_commit(sm.afree);
            __syncwarp();
        }
    } else if (warp < FT_EPI_WARP0 + FT_EPI_WARPS) {
        // epilogue: thread = query.  Hits go to the thread's pending queue (O(1)); a full queue is merged into the sorted
        // list by the WHOLE WARP (bitonic sort of the queue across the lanes + merge), so an insertion costs the warp
        // ~20 clk amortised instead of a divergent per-thread shift of a shared-memory list.
        const uint32_t lane_group = (uint32_t)warp & 3u;
        const uint32_t tid = lane_group * 32 + lane;  // query of the block == TMEM lane
        const float INF = __int_as_float(0x7f800000);
        uint64_t t = 0;
        // merges the queue of thread `who` (c entries) into its list; returns the new (32nd key, 8th key), +inf while
        // the list holds fewer entries
        auto flush = [&](uint32_t who, uint32_t c) -> float2 {
            const uint32_t tq = lane_group * 32 + who;
            __syncwarp();
            float d = (uint32_t)lane < c ? sm.qkey[lane * FT_M + tq] : INF;
            uint32_t r = (uint32_t)lane < c ? sm.qpos[lane * FT_M + tq] : 0xffffffffu;
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1) {  // bitonic sort, DESCENDING across the lanes
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    const bool desc_block = (lane & k) == 0 || k == 32;
                    const bool lower = (lane & j) == 0;
                    sel_cmpx(d, r, j, desc_block ? !lower : lower);
                }
            }
            float ld_ = sm.lkey[lane * FT_M + tq];
            uint32_t lr_ = sm.lpos[lane * FT_M + tq];
            if (sel_less(d, r, ld_, lr_)) {  // list ascending, queue descending: lane-wise min = smaller half (bitonic)
                ld_ = d;
                lr_ = r;
            }
#pragma unroll
            for (int j = 16; j > 0; j >>= 1) sel_cmpx(ld_, lr_, j, (lane & j) == 0);
            sm.lkey[lane * FT_M + tq] = ld_;
            sm.lpos[lane * FT_M + tq] = lr_;
            __syncwarp();
            const float last_d = __shfl_sync(FULL_MASK, ld_, 31), d8 = __shfl_sync(FULL_MASK, ld_, TC_SLOT_RANK - 1);
            const uint32_t last_p = __shfl_sync(FULL_MASK, lr_, 31), p8 = __shfl_sync(FULL_MASK, lr_, TC_SLOT_RANK - 1);
            return make_float2(last_p != 0xffffffffu ? last_d : INF, p8 != 0xffffffffu ? d8 : INF);
        };
        for (uint64_t it = blockIdx.x; it < nitems; it += gridDim.x) {
            uint64_t r0, r1;
            uint32_t qb;
            item_rows(it, r0, r1, qb);
            const uint32_t q = qb * FT_M + tid;
            const bool qlive = q < p.nq;
#pragma unroll
            for (int e = 0; e < FT_LIST; ++e) {
                sm.lkey[e * FT_M + tid] = INF;
                sm.lpos[e * FT_M + tid] = 0xffffffffu;
            }
            uint32_t qcnt = 0;
            float thr_own = INF, pub_own = INF;
            // shared bound: 4 slots per query, slot = slice mod 4 holds the smallest 8th key of any list of those slices;
            // 4 x 8 distinct rows lie at or below the largest slot (same argument as TcScanParams::qtau, ivf_tc.cuh)
            uint32_t* my_slot = p.qtau + 4 * (uint64_t)q + (uint32_t)((it / p.nqb) & 3u);
            float tg = qlive ? tau_decode(tau_shared_bits(p.qtau, q)) : -INF;  // dead query rows never insert
            for (uint64_t r = r0; r < r1; r += FT_N, ++t) {
                const uint32_t buf = (uint32_t)(t & 1), tph = (uint32_t)((t >> 1) & 1);
                tc::mbar_wait(&sm.nbar[t & 3], (uint32_t)((t >> 2) & 1));
                tc::mbar_wait(&sm.tfull[buf], tph);
                tc::fence_after_thread_sync();
                const uint32_t tacc = tmem_base + ((lane_group * 32u) << 16) + FT_ACC_COL0 + buf * FT_N;
                uint32_t va[32], vb[32];
                tc::tmem_ld_32_nowait(tacc, va);
                tc::tmem_ld_32_nowait(tacc + 32, vb);
                tc::tmem_ld_wait_32(va);
                tc::tmem_ld_wait_32(vb);
                tc::fence_before_thread_sync();
                __syncwarp();  // also the reconvergence point of the divergent selection below: tcgen05.ld is .aligned
                tc::mbar_arrive(&sm.tempty[buf]);  // per lane: also releases this lane's reads of the norm ring
                if ((t & 3) == 3 && qlive) tg = fminf(tg, tau_decode(tau_shared_bits(p.qtau, q)));
                const float* nr = sm.nrm + (t & 3) * FT_N;
                float qm[16];
#pragma unroll
                for (int j = 0; j < 64; j += 4) {
                    const float4 n4 = *reinterpret_cast<const float4*>(nr + j);
                    uint32_t* v = j < 32 ? &va[j] : &vb[j - 32];
                    const float k0 = __fmaf_rn(-2.0f, __uint_as_float(v[0]), n4.x);
                    const float k1 = __fmaf_rn(-2.0f, __uint_as_float(v[1]), n4.y);
                    const float k2 = __fmaf_rn(-2.0f, __uint_as_float(v[2]), n4.z);
                    const float k3 = __fmaf_rn(-2.0f, __uint_as_float(v[3]), n4.w);
                    v[0] = __float_as_uint(k0), v[1] = __float_as_uint(k1), v[2] = __float_as_uint(k2),
                    v[3] = __float_as_uint(k3);
                    qm[j >> 2] = fminf(fminf(k0, k1), fminf(k2, k3));
                }
                float gm[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) gm[g] = fminf(fminf(qm[4 * g], qm[4 * g + 1]), fminf(qm[4 * g + 2], qm[4 * g + 3]));
                const float tm = fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3]));
                float thr = fminf(thr_own, tg);
                int last_col = -1;  // columns <= last_col of this tile are already queued
                bool more;
                do {
                    more = false;
                    if (tm < thr) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (gm[g] < thr) {
#pragma unroll
                                for (int qd = 0; qd < 4; ++qd) {
                                    if (qm[4 * g + qd] < thr) {
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            const int col = 16 * g + 4 * qd + e;
                                            const float key = __uint_as_float(col < 32 ? va[col] : vb[col - 32]);
                                            if (key < thr && col > last_col) {
                                                if (qcnt < FT_QCAP) {
                                                    sm.qkey[qcnt * FT_M + tid] = key;
                                                    sm.qpos[qcnt * FT_M + tid] = (uint32_t)(r + col);
                                                    ++qcnt;
                                                    last_col = col;
                                                } else {
                                                    more = true;  // queue full: flush, then rescan the rest of the tile
                                                }
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                    unsigned fm = __ballot_sync(FULL_MASK, qcnt == FT_QCAP);
                    while (fm) {  // warp-uniform: merge every full queue
                        const uint32_t who = (uint32_t)__ffs(fm) - 1;
                        const float2 f = flush(who, FT_QCAP);
                        if ((uint32_t)lane == who) {
                            qcnt = 0;
                            thr_own = f.x;
                            if (f.y < pub_own) {
                                atomicMin(my_slot, tau_encode(f.y));
                                pub_own = f.y;
                            }
                            thr = fminf(thr_own, tg);
                        }
                        fm &= fm - 1;
                    }
                } while (__any_sync(FULL_MASK, more));
            }
            // end of the item: merge what is still queued, hand the sorted list out
            {
                unsigned fm = __ballot_sync(FULL_MASK, qcnt > 0);
                while (fm) {
                    const uint32_t who = (uint32_t)__ffs(fm) - 1;
                    const uint32_t c = __shfl_sync(FULL_MASK, qcnt, who);
                    const float2 f = flush(who, c);
                    if ((uint32_t)lane == who) {
                        qcnt = 0;
                        if (f.y < pub_own) atomicMin(my_slot, tau_encode(f.y));
                    }
                    fm &= fm - 1;
                }
            }
            if (qlive) {
                const uint64_t slice = it / p.nqb;
                float* od = p.part_d + ((uint64_t)q * p.nslices + slice) * FT_LIST;
                uint32_t* op = p.part_p + ((uint64_t)q * p.nslices + slice) * FT_LIST;
#pragma unroll
                for (int e = 0; e < FT_LIST; e += 4) {
                    *reinterpret_cast<float4*>(od + e) =
                        make_float4(sm.lkey[e * FT_M + tid], sm.lkey[(e + 1) * FT_M + tid], sm.lkey[(e + 2) * FT_M + tid],
                                    sm.lkey[(e + 3) * FT_M + tid]);
                    *reinterpret_cast<uint4*>(op + e) =
                        make_uint4(sm.lpos[e * FT_M + tid], sm.lpos[(e + 1) * FT_M + tid], sm.lpos[(e + 2) * FT_M + tid],
                                   sm.lpos[(e + 3) * FT_M + tid]);
                }
            }
            __syncwarp();
        }
    } else {
        // converters: landed query chunk -> rounded to nearest tf32 -> tensor memory (the A operand of the whole item)
        const uint32_t lane_group = (uint32_t)warp & 3u;
        const uint32_t row = lane_group * 32 + lane;
        uint32_t chunk = 0;
        uint64_t i = 0;
        for (uint64_t it = blockIdx.x; it < nitems; it += gridDim.x, ++i) {
            for (uint32_t kc = 0; kc < nk; ++kc, ++chunk) {
                tc::mbar_wait(sm.afull, chunk & 1);
                const uint8_t* arow = sm.aslot + row * 128;
                uint32_t v[FT_KC];
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    const float4 f = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7u)) << 4));
                    v[4 * c + 0] = round_tf32_bits(__float_as_uint(f.x));
                    v[4 * c + 1] = round_tf32_bits(__float_as_uint(f.y));
                    v[4 * c + 2] = round_tf32_bits(__float_as_uint(f.z));
                    v[4 * c + 3] = round_tf32_bits(__float_as_uint(f.w));
                }
                __syncwarp();
                tc::mbar_arrive(sm.aempty);  // per lane: the staging slot's bytes are in this lane's registers
                if (kc == 0 && i > 0) {  // the previous item's MMAs still read the A columns
                    tc::mbar_wait(sm.afree, (uint32_t)((i - 1) & 1));
                    tc::fence_after_thread_sync();
                }
                tc::tmem_st_32(tmem_base + ((lane_group * 32u) << 16) + FT_A_COL0 + kc * FT_KC, v);
            }
            tc::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(sm.aready);
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, FT_TMEM_COLS);
}

}  // namespace vers
