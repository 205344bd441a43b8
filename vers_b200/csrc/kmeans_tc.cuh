// kmeans_tc.cuh — tensor-core candidate pass of assign_to_clusters (indexes/ivfflat.rs:29-46), included by kmeans.cu.
//
// N x C x D distance evaluations are a GEMM; the reference's answer is an argmin per row.  This kernel computes, for a
// block of 128 rows against ALL centroids (tiles of 128), key(c) = ||c||^2 - 2 x.c on the tensor cores
// (TMA -> shared memory -> tcgen05.mma kind::tf32 M=128 N=128 K=8 -> TMEM, two 128-column accumulators so the MMAs of
// centroid tile t+1 overlap the argmin epilogue of tile t) in SPLIT PRECISION (x = hi + lo, c = hi + lo, three MMAs
// per K step, x_lo parked in tensor memory by converter warps: see ivf_tc.cuh) and keeps the smallest and
// second-smallest key per row.  A row is CERTIFIED when the gap between them exceeds the rounding-error allowance
// of both values: then the reference's exact-order distances must put the same centroid first.  Rows that are not
// certified (ties, near-ties) are appended to a list and re-assigned by the exact-order kernel (assign_kernel with a
// row gather) — same two-stage contract as the inverted-list scan, so assignments stay bit-identical.
#pragma once
#include "tc.cuh"

namespace vers {

constexpr int KA_M = 128, KA_N = 128, KA_KC = 32, KA_STAGES = 4, KA_EPI_WARPS = 4, KA_CONV_WARPS = 4;
constexpr int KA_THREADS = (2 + KA_EPI_WARPS + KA_CONV_WARPS) * 32;
constexpr int KA_A_BYTES = KA_M * KA_KC * 4, KA_B_BYTES = KA_N * KA_KC * 4;
constexpr int KA_STAGE_BYTES = KA_A_BYTES + 2 * KA_B_BYTES;  // x | c_hi | c_lo
constexpr int KA_OFF_NRM = KA_STAGES * KA_STAGE_BYTES;         // [2][KA_N] ||c||^2 of the centroid tile per accumulator
constexpr int KA_OFF_BAR = KA_OFF_NRM + 2 * KA_N * 4;
constexpr int KA_SMEM_BYTES = 1024 + KA_OFF_BAR + 256;
constexpr uint32_t KA_ALO_COL0 = 2 * KA_N;                    // x_lo tiles live in TMEM after the two accumulators,
constexpr uint32_t KA_AHI_COL0 = KA_ALO_COL0 + KA_STAGES * KA_KC;  // then the row tiles themselves (x_hi after truncation)
constexpr uint32_t KA_TMEM_COLS = 512;                        // 2*128 + 4*32 + 4*32 = 512

struct TcAssignParams {
    uint64_t n_rows;
    uint32_t C, ld;
    const float* row_norm;   // [n] ||x||^2 (any order)
    const float* cent_norm;  // [round_up(C, 128)] ||c||^2 (any order), +inf past C
    const uint32_t* ncmax_bits;  // max ||c||^2
    uint32_t* assign;        // [n] candidate argmin
    uint32_t* flagged;       // [n] compacted list of uncertified rows
    uint32_t* n_flagged;     // [1]
};

// centroids split into tf32 hi / lo parts (the tensor core truncates fp32 -> tf32, so hi = c with 13 low bits cleared)
__global__ void split_tf32_kernel(const float* __restrict__ in, uint64_t n4, float* __restrict__ hi,
                                  float* __restrict__ lo) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(in)[i], h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = __fsub_rn(v.x, h.x);
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = __fsub_rn(v.y, h.y);
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = __fsub_rn(v.z, h.z);
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = __fsub_rn(v.w, h.w);
        reinterpret_cast<float4*>(hi)[i] = h;
        reinterpret_cast<float4*>(lo)[i] = l;
    }
}

// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = argmin epilogue, 6..9 = converters.
// Per 32-wide K chunk: x_hi.c_hi + x_hi.c_lo + x_lo.c_hi, all three accumulated into the same 128-column accumulator.
// ALL THREE take their A operand from TENSOR MEMORY: the converter warps park the landed row tile (the tensor core
// truncates it to x_hi) next to x_lo.  With M = N = 128 an MMA that reads A and B from shared memory needs
// 8 KB / 64 cycles = the whole 128 B/clk of the SM, and the TMA writes and converter reads come on top: measured 67 %
// tensor-pipe activity.  With A in TMEM shared memory only serves the B tiles.
__global__ void __launch_bounds__(KA_THREADS, 1)
    tc_assign_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_chi,
                     const __grid_constant__ CUtensorMap tmap_clo, TcAssignParams p) {
    extern __shared__ uint8_t ka_smem_raw[];
    const uint32_t raw = tc::smem_u32(ka_smem_raw);
    uint8_t* smem = ka_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    float* nrm = reinterpret_cast<float*>(smem + KA_OFF_NRM);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + KA_OFF_BAR);
    uint64_t* conv = full + KA_STAGES;
    uint64_t* empty = conv + KA_STAGES;
    uint64_t* tfull = empty + KA_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* nbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nbar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nk = (p.ld + KA_KC - 1) / KA_KC;
    const uint32_t nct = (p.C + KA_N - 1) / KA_N;
    const uint64_t nrb = (p.n_rows + KA_M - 1) / KA_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < KA_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&conv[s], KA_CONV_WARPS * 32);  // every converter lane arrives itself
            tc::mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tfull[b], 1);
            tc::mbar_init(&tempty[b], KA_EPI_WARPS * 32);  // every epilogue lane arrives itself
            tc::mbar_init(&nbar[b], 1);
        }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmap_rows);
        tc::tma_prefetch_desc(&tmap_chi);
        tc::tma_prefetch_desc(&tmap_clo);
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, KA_TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
                for (uint32_t ct = 0; ct < nct; ++ct) {
                    for (uint32_t kc = 0; kc < nk; ++kc) {
                        tc::mbar_wait(&empty[stage], phase ^ 1);
                        tc::mbar_arrive_expect_tx(&full[stage], KA_STAGE_BYTES);
                        uint8_t* sa = smem + stage * KA_STAGE_BYTES;
                        tc::tma_load_2d(sa, &tmap_rows, &full[stage], (int32_t)(kc * KA_KC), (int32_t)(rb * KA_M));
                        tc::tma_load_2d(sa + KA_A_BYTES, &tmap_chi, &full[stage], (int32_t)(kc * KA_KC),
                                        (int32_t)(ct * KA_N));
                        tc::tma_load_2d(sa + KA_A_BYTES + KA_B_BYTES, &tmap_clo, &full[stage], (int32_t)(kc * KA_KC),
                                        (int32_t)(ct * KA_N));
                        if (++stage == KA_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_tf32(KA_M, KA_N);
            uint32_t stage = 0, phase = 0, tile_ctr = 0;
            for (uint64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
                for (uint32_t ct = 0; ct < nct; ++ct) {
                    const uint32_t buf = tile_ctr & 1, tphase = (tile_ctr >> 1) & 1;
                    tc::mbar_wait(&tempty[buf], tphase ^ 1);
                    tc::fence_after_thread_sync();
                    // the tile's ||c||^2 ride along into shared memory for the epilogue (the buffer is free: the
                    // epilogue released it together with the accumulator)
                    tc::mbar_arrive_expect_tx(&nbar[buf], KA_N * 4);
                    tc::bulk_load(nrm + buf * KA_N, p.cent_norm + (size_t)ct * KA_N, KA_N * 4, &nbar[buf]);
                    const uint32_t d_tmem = tmem_base + buf * KA_N;
                    for (uint32_t kc = 0; kc < nk; ++kc) {
                        tc::mbar_wait(&full[stage], phase);
                        tc::fence_after_thread_sync();
                        const uint32_t sa = tc::smem_u32(smem + stage * KA_STAGE_BYTES);
                        const uint64_t dbh = tc::smem_desc_k_sw128(sa + KA_A_BYTES);
                        const uint64_t dbl = tc::smem_desc_k_sw128(sa + KA_A_BYTES + KA_B_BYTES);
                        tc::mbar_wait(&conv[stage], phase);  // x and x_lo of this stage are in tensor memory
                        tc::fence_after_thread_sync();
                        const uint32_t a_hi = tmem_base + KA_AHI_COL0 + stage * KA_KC;
                        const uint32_t a_lo = tmem_base + KA_ALO_COL0 + stage * KA_KC;
#pragma unroll
                        for (uint32_t kk = 0; kk < KA_KC / 8; ++kk) {
                            tc::mma_tf32_ts(d_tmem, a_hi + 8 * kk, dbh + 2 * kk, idesc, (kc | kk) != 0);
                            tc::mma_tf32_ts(d_tmem, a_hi + 8 * kk, dbl + 2 * kk, idesc, 1);
                            tc::mma_tf32_ts(d_tmem, a_lo + 8 * kk, dbh + 2 * kk, idesc, 1);
                        }
                        tc::mma_commit(&empty[stage]);
                        if (++stage == KA_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    tc::mma_commit(&tfull[buf]);
                    ++tile_ctr;
                }
            }
        }
    } else if (warp < 2 + KA_EPI_WARPS) {
        // epilogue: thread = row; running smallest / second smallest key over all centroid tiles
        const int lane_group = warp & 3;
        uint32_t tile_ctr = 0;
        const double u = 5.9604644775390625e-08;  // 2^-24
        const double ncmax = (double)__uint_as_float(*p.ncmax_bits);
        for (uint64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
            const uint64_t row = rb * KA_M + (uint64_t)(lane_group * 32 + lane);
            float b1 = __int_as_float(0x7f800000), b2 = __int_as_float(0x7f800000);
            uint32_t c1 = 0xffffffffu;
            for (uint32_t ct = 0; ct < nct; ++ct) {
                const uint32_t buf = tile_ctr & 1, tphase = (tile_ctr >> 1) & 1;
                tc::mbar_wait(&nbar[buf], tphase);
                tc::mbar_wait(&tfull[buf], tphase);
                tc::fence_after_thread_sync();
                const uint32_t tacc = tmem_base + ((uint32_t)(lane_group * 32) << 16) + buf * KA_N;
                const float* nr = nrm + buf * KA_N;
                // branch-free running (smallest, second smallest, index of the smallest) over the 128 columns; two
                // 32-column register buffers so the next tcgen05.ld is in flight while one is consumed.  Strict `<`:
                // the lowest index among equal keys stays first; an equal key lands in b2 (gap 0 => uncertified)
                const float b1_in = b1;
                uint32_t loc = 0;
                uint32_t va[32], vb[32];
                tc::tmem_ld_32_nowait(tacc, va);
#pragma unroll
                for (int h = 0; h < KA_N / 32; ++h) {
                    if (h & 1) tc::tmem_ld_wait_32(vb); else tc::tmem_ld_wait_32(va);
                    if (h + 1 < KA_N / 32) {
                        if (h & 1) tc::tmem_ld_32_nowait(tacc + 32 * (h + 1), va);
                        else tc::tmem_ld_32_nowait(tacc + 32 * (h + 1), vb);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 n4 = *reinterpret_cast<const float4*>(nr + 32 * h + j);
                        const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float acc = __uint_as_float((h & 1) ? vb[j + e] : va[j + e]);
                            const float key = __fmaf_rn(-2.0f, acc, nn[e]);
                            b2 = fminf(b2, fmaxf(b1, key));
                            const bool lt = key < b1;
                            b1 = lt ? key : b1;
                            loc = lt ? (uint32_t)(32 * h + j + e) : loc;
                        }
                    }
                }
                if (b1 < b1_in) c1 = ct * KA_N + loc;
                tc::fence_before_thread_sync();
                __syncwarp();  // also the reconvergence point of the divergent selection above: tcgen05.ld is .aligned
                tc::mbar_arrive(&tempty[buf]);  // per lane: each lane's own reads (accumulator, norm ring) are released
                ++tile_ctr;
            }
            if (row < p.n_rows) {
                p.assign[row] = c1;
                // certificate: every other centroid has key >= b2; |d~ - d_true| <= E, d_ref within (1 +- rho) d_true.
                // E: fp32 norm/key terms + split-precision product error 3*2^-20 + accumulation allowance (n+8) 2^-22,
                // each per unit of (||x||^2 + ||c||^2_max)  (same model as the list-scan certificate, ivf.cu)
                const double nx = (double)__ldg(p.row_norm + row);
                const double S = nx + ncmax;
                const double E = (1.01 * (2.0 * p.ld + 8.0) * u + 3.003 / 1048576.0 + (p.ld + 8.0) * 2.384185791015625e-07) * S;
                const double rho = (p.ld + 3.0) * u;
                const bool certified = ((double)b2 + nx - E) * (1.0 - rho) > ((double)b1 + nx + E) * (1.0 + rho);
                if (!certified) {
                    uint32_t at = atomicAdd(p.n_flagged, 1u);
                    p.flagged[at] = (uint32_t)row;
                }
            }
        }
    } else {
        // converters: x_lo = x - trunc_tf32(x) of the landed row tile -> tensor memory (see ivf_tc.cuh)
        const int lane_group = warp & 3;
        const uint32_t row = (uint32_t)(lane_group * 32 + lane);
        uint32_t stage = 0, phase = 0;
        for (uint64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
            for (uint32_t ct = 0; ct < nct; ++ct) {
                for (uint32_t kc = 0; kc < nk; ++kc) {
                    tc::mbar_wait(&full[stage], phase);
                    const uint8_t* arow = smem + stage * KA_STAGE_BYTES + row * 128;
                    uint32_t lo[KA_KC], hi[KA_KC];
#pragma unroll
                    for (uint32_t c = 0; c < 8; ++c) {
                        const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7u)) << 4));
                        hi[4 * c + 0] = __float_as_uint(v.x);
                        hi[4 * c + 1] = __float_as_uint(v.y);
                        hi[4 * c + 2] = __float_as_uint(v.z);
                        hi[4 * c + 3] = __float_as_uint(v.w);
                        lo[4 * c + 0] = __float_as_uint(__fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u)));
                        lo[4 * c + 1] = __float_as_uint(__fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u)));
                        lo[4 * c + 2] = __float_as_uint(__fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u)));
                        lo[4 * c + 3] = __float_as_uint(__fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u)));
                    }
                    const uint32_t tl = tmem_base + ((uint32_t)(lane_group * 32) << 16);
                    tc::tmem_st_32(tl + KA_AHI_COL0 + stage * KA_KC, hi);  // the tensor core truncates it to x_hi
                    tc::tmem_st_32(tl + KA_ALO_COL0 + stage * KA_KC, lo);
                    tc::fence_before_thread_sync();
                    __syncwarp();
                    tc::mbar_arrive(&conv[stage]);
                    if (++stage == KA_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, KA_TMEM_COLS);
}

// =====================================================================================================================
// tc_assign1_kernel — TF32-FIRST candidate pass for ld <= 128 (BASELINE configs[4]: 50M x 128, 16384 centroids).
//
// The split-precision kernel above is bound by L2 -> SM operand traffic, not by the tensor pipe: per 128 x 128 tile it
// moves 192 KB (the row tile re-fetched per K chunk + c_hi + c_lo) against an L2 delivery of ~42 B/clk/SM, i.e.
// ~4500 clk for 3072 clk of MMAs (ncu: 67 % tensor-pipe activity).  This kernel removes both factors:
//   * ONE tcgen05.mma kind::tf32 per K step instead of three.  Both operands are rounded to nearest tf32 (centroids by
//     a prep kernel, rows by the converter warps), so |x.c - x~.c~| <= 2^-10 (1 + 2^-12) |x||c|.  That is too coarse
//     to certify the argmin from the candidate values alone, so per row the FOUR smallest keys are kept (+ the fifth
//     key as the bound), the four candidates are re-ranked with the reference's exact-order arithmetic
//     (indexes/base.rs:119-126) inside the same kernel, and the certificate compares the exact best distance with the
//     fifth key: (k5 + ||x||^2 - E)(1 - rho) > d_ref(best)  =>  no centroid outside the four can win or tie.
//     Uncertified rows (near-ties within the tf32 error) are appended to the flagged list and redone exactly.
//   * the rows are loaded and converted ONCE per row block and stay resident in TENSOR MEMORY as the A operand
//     (ld <= 128 fp32 columns); a CTA owns TWO 128-row blocks (2 x 128 A columns + 2 x 2 x 64 accumulator columns =
//     all 512 TMEM columns), so a 64-centroid B tile (32 KB, the only thing streamed from L2) feeds 2 x 16 MMAs of
//     M=128 N=64 K=8: 32 B/clk/SM at full tensor rate.
//   * epilogue: key = ||c||^2 - 2 acc per element (one FFMA), a min-tree per 16 columns and ONE compare against the
//     row's current fifth key; the insertion code runs only for groups that contain a new top-5 entry.
// Warp roles (14 warps): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = epilogue (warps 2..5 row block 0,
// 6..9 row block 1; warp w reads TMEM lanes 32*(w%4)..+31), 10..13 = converters.  (Measured: splitting a row's columns
// over two epilogue threads — 16 epilogue warps — is SLOWER, 323 -> 368 ms per C5 pass with kind::f16: the extra
// warps compete with the single MMA-issuing thread for issue slots and with the MMA's A operand for TMEM reads.)
constexpr int K1_M = 128, K1_N = 64, K1_KC = 32, K1_STAGES = 5, K1_MAX_KCH = 4;
constexpr int K1_EPI_WARP0 = 2, K1_EPI_WARPS = 8, K1_CONV_WARP0 = 10, K1_CONV_WARPS = 4;
constexpr int K1_THREADS = (2 + K1_EPI_WARPS + K1_CONV_WARPS) * 32;
constexpr int K1_BOX_BYTES = K1_N * K1_KC * 4;                 // one [64 centroids x 32 floats] box
constexpr int K1_STAGE_BYTES = K1_MAX_KCH * K1_BOX_BYTES;      // a whole 64-centroid tile (all K chunks)
constexpr int K1_ASLOT_BYTES = K1_M * K1_KC * 4;               // one [128 rows x 32 floats] staging slot
constexpr int K1_OFF_ASLOT = K1_STAGES * K1_STAGE_BYTES;
constexpr int K1_OFF_NRM = K1_OFF_ASLOT + 2 * K1_ASLOT_BYTES;  // [4][64] ||c||^2 ring
constexpr int K1_OFF_BAR = K1_OFF_NRM + 4 * K1_N * 4;
constexpr int K1_SMEM_BYTES = 1024 + K1_OFF_BAR + 512;
constexpr uint32_t K1_ACC_COL0 = 0;      // accumulator of (row block rb, buffer b) at column (2 rb + b) * 64
constexpr uint32_t K1_A_COL0 = 256;      // rows of row block rb at column 256 + 128 rb
constexpr uint32_t K1_TMEM_COLS = 512;

struct TcAssign1Params {
    uint64_t n_rows;
    uint32_t C, ld;
    const float* rows;       // [n][ld] the rows themselves (exact rerank)
    const float* cents;      // [C][ld] the centroids themselves (exact rerank)
    const float* cent_tiles; // tile_image_tf32_kernel's image of the centroids: [ceil(C/64)][nk][64][32] swizzled
                             // (F16: tile_image_f16_kernel's, [ceil(C/64)][nk16][64 rows][64 halfs], values * scale)
    float scale;             // F16: the power of two both operands are multiplied with before the fp16 rounding
    float key_scale;         // F16: -2 / scale^2 (key = ||c||^2 + key_scale * acc); tf32: -2
    const uint32_t* f16_bad; // F16: elements of the centroid image that did not fit fp16 (nothing is certified then)
    const float* row_norm;   // [n] ||x||^2 (any order)
    const float* cent_norm;  // [round_up(C, 128)] ||c||^2 (any order), +inf past C
    const uint32_t* ncmax_bits;
    uint32_t* assign;
    uint32_t* flagged;
    uint32_t* n_flagged;
};

// sorted insertion into the five smallest keys (ids for the first four); strict `<`: an equal key stays behind
__device__ __forceinline__ void k1_insert(float (&k)[5], uint32_t (&id)[4], float key, uint32_t idx) {
    const bool c3 = key < k[3], c2 = key < k[2], c1 = key < k[1], c0 = key < k[0];
    k[4] = c3 ? k[3] : key;
    k[3] = c3 ? (c2 ? k[2] : key) : k[3];
    id[3] = c3 ? (c2 ? id[2] : idx) : id[3];
    k[2] = c2 ? (c1 ? k[1] : key) : k[2];
    id[2] = c2 ? (c1 ? id[1] : idx) : id[2];
    k[1] = c1 ? (c0 ? k[0] : key) : k[1];
    id[1] = c1 ? (c0 ? id[0] : idx) : id[1];
    k[0] = c0 ? key : k[0];
    id[0] = c0 ? idx : id[0];
}

// one elected lane of a converged warp (the compiler then emits the tcgen05 / TMA instructions straight, without the
// per-instruction ELECT ... BRA.U.ANY loop it wraps around them under a divergent `if (lane == 0)`: measured ~50
// clk per tcgen05.mma of issue overhead, more than the 32 clk an M=128 N=64 K=8 MMA runs)
__device__ __forceinline__ bool k1_elect_one() {
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

// F16: the same kernel with fp16 operands (kind::f16, K = 16 per MMA): fp16 rounds to the same 11-bit significand as
// tf32, so the certificate's error model is unchanged, while the tensor pipe runs at twice the rate and the centroid
// tiles are half the bytes.  Both operands are scaled by one power of two (host: from max ||row||^2; centroids are means
// of rows, so they fit as well — an image with elements that do not fit certifies nothing, f16_bad).  The rows are
// packed two per 32-bit tensor-memory column (even dimension in the low half).
template <int NK, bool F16>  // K chunks of 32 floats: ld <= 32 NK
__global__ void __launch_bounds__(K1_THREADS, 1)
    tc_assign1_kernel(const __grid_constant__ CUtensorMap tmap_rows, TcAssign1Params p) {
    extern __shared__ uint8_t k1_smem_raw[];
    const uint32_t raw = tc::smem_u32(k1_smem_raw);
    uint8_t* smem = k1_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    float* nrm = reinterpret_cast<float*>(smem + K1_OFF_NRM);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + K1_OFF_BAR);
    uint64_t* empty = full + K1_STAGES;
    uint64_t* afull = empty + K1_STAGES;   // [2] staging slot landed
    uint64_t* aempty = afull + 2;          // [2] staging slot read by all converter warps
    uint64_t* aready = aempty + 2;         // [1] both row blocks of this pair are in tensor memory
    uint64_t* afree = aready + 1;          // [1] every MMA of the pair has completed: A may be overwritten
    uint64_t* tfull = afree + 1;           // [2 rb][2 buf]
    uint64_t* tempty = tfull + 4;          // [2 rb][2 buf]
    uint64_t* nbar = tempty + 4;           // [4] norm ring
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nbar + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t nk = NK;
    constexpr uint32_t nk16 = (NK + 1) / 2;                          // F16: B boxes / MMA groups of 64 dimensions
    constexpr uint32_t tile_bytes = F16 ? nk16 * K1_BOX_BYTES : nk * K1_BOX_BYTES;
    constexpr uint32_t a_cols = F16 ? 64u : 128u;                    // tensor-memory columns of one 128-row block
    const uint32_t nct = (p.C + K1_N - 1) / K1_N;
    const uint64_t nrbp = (p.n_rows + 2 * K1_M - 1) / (2 * K1_M);
    const uint64_t my_pairs = blockIdx.x < nrbp ? (nrbp - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint64_t total_tiles = my_pairs * nct;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K1_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&afull[s], 1);
            tc::mbar_init(&aempty[s], K1_CONV_WARPS * 32);  // every converter lane arrives itself
        }
        tc::mbar_init(aready, K1_CONV_WARPS);
        tc::mbar_init(afree, 1);
        for (int b = 0; b < 4; ++b) {
            tc::mbar_init(&tfull[b], 1);
            tc::mbar_init(&tempty[b], K1_EPI_WARPS / 2 * 32);  // every epilogue lane arrives itself
            tc::mbar_init(&nbar[b], 1);
        }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmap_rows);
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, K1_TMEM_COLS);
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // TMA producer: the whole warp walks the loops (and waits), one elected lane issues
        uint32_t stage = 0, phase = 0, chunk = 0;
        for (uint64_t rbp = blockIdx.x; rbp < nrbp; rbp += gridDim.x) {
            for (uint32_t rb = 0; rb < 2; ++rb) {
#pragma unroll
                for (uint32_t kc = 0; kc < nk; ++kc, ++chunk) {
                    const uint32_t slot = chunk & 1, ph = (chunk >> 1) & 1;
                    tc::mbar_wait(&aempty[slot], ph ^ 1);
                    if (k1_elect_one()) {
                        tc::mbar_arrive_expect_tx(&afull[slot], K1_ASLOT_BYTES);
                        tc::tma_load_2d(smem + K1_OFF_ASLOT + slot * K1_ASLOT_BYTES, &tmap_rows, &afull[slot],
                                        (int32_t)(kc * K1_KC), (int32_t)((rbp * 2 + rb) * K1_M));
                    }
                    __syncwarp();
                }
            }
            for (uint32_t ct = 0; ct < nct; ++ct) {
                tc::mbar_wait(&empty[stage], phase ^ 1);
                if (k1_elect_one()) {  // the whole 64-centroid tile (all K chunks) is one linear copy of its image
                    tc::mbar_arrive_expect_tx(&full[stage], tile_bytes);
                    tc::bulk_load(smem + stage * K1_STAGE_BYTES,
                                  reinterpret_cast<const uint8_t*>(p.cent_tiles) + (size_t)ct * tile_bytes, tile_bytes,
                                  &full[stage]);
                }
                __syncwarp();
                if (++stage == K1_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: converged warp, one elected lane issues.  Per tile: 2 row blocks x NK x 4 MMAs (M=128 N=64 K=8)
        const uint32_t idesc = F16 ? tc::idesc_f16(K1_M, K1_N) : tc::idesc_tf32(K1_M, K1_N);
        uint32_t stage = 0, phase = 0;
        uint64_t t = 0;
        if (total_tiles && k1_elect_one()) {
            // the norms of tile 0; tile t+1's are requested when tile t starts (slot (t+1) % 4 was last read for tile
            // t-3, which both epilogues finished before releasing tile t-2)
            tc::mbar_arrive_expect_tx(&nbar[0], K1_N * 4);
            tc::bulk_load(nrm, p.cent_norm, K1_N * 4, &nbar[0]);
        }
        __syncwarp();
        for (uint64_t i = 0; i < my_pairs; ++i) {
            tc::mbar_wait(aready, (uint32_t)(i & 1));
            tc::fence_after_thread_sync();
            for (uint32_t ct = 0; ct < nct; ++ct, ++t) {
                const uint32_t buf = (uint32_t)(t & 1), tph = (uint32_t)((t >> 1) & 1);
                tc::mbar_wait(&tempty[buf], tph ^ 1);
                tc::mbar_wait(&tempty[2 + buf], tph ^ 1);
                tc::mbar_wait(&full[stage], phase);
                tc::fence_after_thread_sync();
                if (k1_elect_one()) {
                    if (t + 1 < total_tiles) {
                        const uint32_t slot = (uint32_t)((t + 1) & 3), nct1 = (ct + 1 == nct) ? 0 : ct + 1;
                        tc::mbar_arrive_expect_tx(&nbar[slot], K1_N * 4);
                        tc::bulk_load(nrm + slot * K1_N, p.cent_norm + (size_t)nct1 * K1_N, K1_N * 4, &nbar[slot]);
                    }
                    const uint32_t sb = tc::smem_u32(smem + stage * K1_STAGE_BYTES);
#pragma unroll
                    for (uint32_t rb = 0; rb < 2; ++rb) {
                        const uint32_t d_tmem = tmem_base + K1_ACC_COL0 + (2 * rb + buf) * K1_N;
                        const uint32_t a_tmem = tmem_base + K1_A_COL0 + rb * a_cols;
                        if (F16) {
#pragma unroll
                            for (uint32_t kc = 0; kc < nk16; ++kc) {  // 64 dimensions = 32 packed columns = 4 MMAs of K 16
                                const uint64_t db = tc::smem_desc_k_sw128(sb + kc * K1_BOX_BYTES);
#pragma unroll
                                for (uint32_t kk = 0; kk < 4; ++kk)
                                    tc::mma_f16_ts(d_tmem, a_tmem + kc * 32 + 8 * kk, db + 2 * kk, idesc, (kc | kk) != 0);
                            }
                        } else {
#pragma unroll
                            for (uint32_t kc = 0; kc < nk; ++kc) {
                                const uint64_t db = tc::smem_desc_k_sw128(sb + kc * K1_BOX_BYTES);
#pragma unroll
                                for (uint32_t kk = 0; kk < K1_KC / 8; ++kk)
                                    tc::mma_tf32_ts(d_tmem, a_tmem + kc * K1_KC + 8 * kk, db + 2 * kk, idesc, (kc | kk) != 0);
                            }
                        }
                        tc::mma_commit(&tfull[2 * rb + buf]);
                    }
                    tc::mma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == K1_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (k1_elect_one()) tc::mma_commit(afree);
            __syncwarp();
        }
    } else if (warp < K1_EPI_WARP0 + K1_EPI_WARPS) {
        // epilogue: thread = row.  Five smallest keys (ids of the first four) over all centroid tiles, then the exact
        // rerank of the four and the certificate.
        const uint32_t rb = (uint32_t)(warp - K1_EPI_WARP0) >> 2, lane_group = (uint32_t)warp & 3u;
        const double u = 5.9604644775390625e-08;  // 2^-24
        const double ncmax = (double)__uint_as_float(*p.ncmax_bits);
        const float INF = __int_as_float(0x7f800000);
        const float ks = F16 ? p.key_scale : -2.0f;
        const bool image_ok = !F16 || *p.f16_bad == 0u;
        uint64_t t = 0;
        for (uint64_t rbp = blockIdx.x; rbp < nrbp; rbp += gridDim.x) {
            const uint64_t row = (rbp * 2 + rb) * K1_M + (uint64_t)(lane_group * 32 + lane);
            float k[5] = {INF, INF, INF, INF, INF};
            uint32_t id[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
            for (uint32_t ct = 0; ct < nct; ++ct, ++t) {
                const uint32_t buf = (uint32_t)(t & 1), tph = (uint32_t)((t >> 1) & 1);
                tc::mbar_wait(&nbar[t & 3], (uint32_t)((t >> 2) & 1));
                tc::mbar_wait(&tfull[2 * rb + buf], tph);
                tc::fence_after_thread_sync();
                const uint32_t tacc = tmem_base + ((lane_group * 32u) << 16) + K1_ACC_COL0 + (2 * rb + buf) * K1_N;
                uint32_t va[32], vb[32];
                tc::tmem_ld_32_nowait(tacc, va);
                tc::tmem_ld_32_nowait(tacc + 32, vb);
                tc::tmem_ld_wait_32(va);
                tc::tmem_ld_wait_32(vb);
                // the accumulator is in registers: hand the buffer back before the selection work
                tc::fence_before_thread_sync();
                __syncwarp();  // also the reconvergence point of the divergent selection below: tcgen05.ld is .aligned
                tc::mbar_arrive(&tempty[2 * rb + buf]);  // per lane: also releases this lane's reads of the norm ring
                const float* nr = nrm + (t & 3) * K1_N;
                const uint32_t c0 = ct * K1_N;
                // phase 1 (straight-line): the 64 keys in place, the min of every quad, of every 16-column group and of
                // the tile
                float qm[16];
#pragma unroll
                for (int j = 0; j < 64; j += 4) {
                    const float4 n4 = *reinterpret_cast<const float4*>(nr + j);
                    uint32_t* v = j < 32 ? &va[j] : &vb[j - 32];
                    const float k0 = __fmaf_rn(ks, __uint_as_float(v[0]), n4.x);
                    const float k1 = __fmaf_rn(ks, __uint_as_float(v[1]), n4.y);
                    const float k2 = __fmaf_rn(ks, __uint_as_float(v[2]), n4.z);
                    const float k3 = __fmaf_rn(ks, __uint_as_float(v[3]), n4.w);
                    v[0] = __float_as_uint(k0), v[1] = __float_as_uint(k1), v[2] = __float_as_uint(k2),
                    v[3] = __float_as_uint(k3);
                    qm[j >> 2] = fminf(fminf(k0, k1), fminf(k2, k3));
                }
                float gm[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) gm[g] = fminf(fminf(qm[4 * g], qm[4 * g + 1]), fminf(qm[4 * g + 2], qm[4 * g + 3]));
                const float tm = fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3]));
                // phase 2: descend only where a new top-5 entry can be (tile -> group of 16 -> quad -> element)
                if (tm < k[4]) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (gm[g] < k[4]) {
#pragma unroll
                            for (int qd = 0; qd < 4; ++qd) {
                                if (qm[4 * g + qd] < k[4]) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const int col = 16 * g + 4 * qd + e;
                                        const float key = __uint_as_float(col < 32 ? va[col] : vb[col - 32]);
                                        if (key < k[4]) k1_insert(k, id, key, c0 + col);
                                    }
                                }
                            }
                        }
                    }
                }
            }
            if (row < p.n_rows) {
                // exact-order distances of the (up to) four candidates: sequential over the dimensions like
                // squared_euclidean (base.rs:119-126), four independent chains per thread
                const float4* xr = reinterpret_cast<const float4*>(p.rows + row * p.ld);
                const float4* cr[4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    cr[r] = reinterpret_cast<const float4*>(p.cents + (size_t)(id[r] < p.C ? id[r] : 0u) * p.ld);
                float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                for (uint32_t j = 0; j < (p.ld >> 2); ++j) {
                    const float4 x = __ldg(xr + j);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float4 c = __ldg(cr[r] + j);
                        float e0 = __fsub_rn(x.x, c.x), e1 = __fsub_rn(x.y, c.y), e2 = __fsub_rn(x.z, c.z),
                              e3 = __fsub_rn(x.w, c.w);
                        d[r] = __fadd_rn(d[r], __fmul_rn(e0, e0));
                        d[r] = __fadd_rn(d[r], __fmul_rn(e1, e1));
                        d[r] = __fadd_rn(d[r], __fmul_rn(e2, e2));
                        d[r] = __fadd_rn(d[r], __fmul_rn(e3, e3));
                    }
                }
                float bd = INF;
                uint32_t bc = 0xffffffffu;
                bool any = false, ordered = true;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (id[r] < p.C) {
                        // a NaN distance (the reference panics on it) sends the row to the exact path, which reports it
                        ordered = ordered && (d[r] == d[r]);
                        if (!any || d[r] < bd || (d[r] == bd && id[r] < bc)) {
                            bd = d[r];
                            bc = id[r];
                            any = true;
                        }
                    }
                }
                p.assign[row] = bc;
                // every centroid outside the four has key >= k[4]; |key + ||x||^2 - d_true| <= E and d_ref >=
                // d_true (1 - rho).  E per unit of (||x||^2 + ||c||^2_max): fp32 norm/key terms, the tf32 rounding of
                // both operands 2^-10 (1 + 2^-12), the accumulation allowance (n + 8) 2^-22 (same model as above)
                const double nx = (double)__ldg(p.row_norm + row);
                const double S = nx + ncmax;
                // (fp16: the same 2^-11 per operand in the normal range; an element below the normal range of the
                // scaled value is off by <= 2^-25 / scale <= 2^-38 max||x||, 2^-30 S covers every such element)
                const double E = (1.01 * (2.0 * p.ld + 8.0) * u + 1.001 / 1024.0 + (p.ld + 8.0) * 2.384185791015625e-07 +
                                  (F16 ? 9.313225746154785e-10 : 0.0)) * S;
                const double rho = (p.ld + 3.0) * u;
                const bool certified = image_ok && any && ordered && ((double)k[4] + nx - E) * (1.0 - rho) > (double)bd;
                if (!certified) {
                    uint32_t at = atomicAdd(p.n_flagged, 1u);
                    p.flagged[at] = (uint32_t)row;
                }
            }
        }
    } else {
        // converters: landed row chunk -> rounded to nearest tf32 -> tensor memory (the A operand of every MMA of the pair)
        const uint32_t lane_group = (uint32_t)warp & 3u;
        const uint32_t row = lane_group * 32 + lane;
        uint32_t chunk = 0;
        for (uint64_t i = 0; i < my_pairs; ++i) {
            for (uint32_t rb = 0; rb < 2; ++rb) {
                for (uint32_t kc = 0; kc < nk; ++kc, ++chunk) {
                    const uint32_t slot = chunk & 1, ph = (chunk >> 1) & 1;
                    tc::mbar_wait(&afull[slot], ph);
                    const uint8_t* arow = smem + K1_OFF_ASLOT + slot * K1_ASLOT_BYTES + row * 128;
                    uint32_t v[K1_KC];
#pragma unroll
                    for (uint32_t c = 0; c < 8; ++c) {
                        const float4 f = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7u)) << 4));
                        if (F16) {  // two dimensions per column, the even one in the low half
                            const __half2 h0 = __floats2half2_rn(__fmul_rn(f.x, p.scale), __fmul_rn(f.y, p.scale));
                            const __half2 h1 = __floats2half2_rn(__fmul_rn(f.z, p.scale), __fmul_rn(f.w, p.scale));
                            v[2 * c + 0] = *reinterpret_cast<const uint32_t*>(&h0);
                            v[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                        } else {
                            v[4 * c + 0] = round_tf32_bits(__float_as_uint(f.x));
                            v[4 * c + 1] = round_tf32_bits(__float_as_uint(f.y));
                            v[4 * c + 2] = round_tf32_bits(__float_as_uint(f.z));
                            v[4 * c + 3] = round_tf32_bits(__float_as_uint(f.w));
                        }
                    }
                    __syncwarp();
                    tc::mbar_arrive(&aempty[slot]);  // per lane: the staging slot's bytes are in this lane's registers
                    if (rb == 0 && kc == 0 && i > 0) {  // the previous pair's MMAs still read the A columns
                        tc::mbar_wait(afree, (uint32_t)((i - 1) & 1));
                        tc::fence_after_thread_sync();
                    }
                    if (F16) {
                        uint32_t h[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) h[e] = v[e];
                        const uint32_t a0 = tmem_base + ((lane_group * 32u) << 16) + K1_A_COL0 + rb * a_cols + kc * 16;
                        tc::tmem_st_16(a0, h);
                        if ((NK & 1) && kc == nk - 1) {  // an odd number of 32-float chunks: the MMAs read 64 dimensions
#pragma unroll
                            for (int e = 0; e < 16; ++e) h[e] = 0u;
                            tc::tmem_st_16(a0 + 16, h);
                        }
                    } else {
                        tc::tmem_st_32(tmem_base + ((lane_group * 32u) << 16) + K1_A_COL0 + rb * 128 + kc * K1_KC, v);
                    }
                }
            }
            tc::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(aready);
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, K1_TMEM_COLS);
}

// squared norms of rows (warp per row, any order) + running max; used for rows and centroids
__global__ void sqnorm_kernel(const float* __restrict__ rows, uint32_t ld, uint64_t count, float* __restrict__ norm,
                              uint32_t* nmax) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    float mx = 0.0f;
    for (uint64_t j = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); j < count; j += warps) {
        const float4* r = reinterpret_cast<const float4*>(rows + j * ld);
        float s = 0.0f;
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            float4 v = r[c];
            s = __fmaf_rn(v.x, v.x, s);
            s = __fmaf_rn(v.y, v.y, s);
            s = __fmaf_rn(v.z, v.z, s);
            s = __fmaf_rn(v.w, v.w, s);
        }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
        if (lane == 0) norm[j] = s;
        mx = fmaxf(mx, s);
    }
    if (nmax && lane == 0 && mx > 0.0f) atomicMax(nmax, __float_as_uint(mx));
}

// scatter the exact re-assignments of the flagged rows back
__global__ void scatter_assign_kernel(const uint32_t* __restrict__ flagged, const uint32_t* __restrict__ n_flagged,
                                      const uint32_t* __restrict__ exact, uint32_t* __restrict__ assign) {
    const uint32_t n = *n_flagged;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        assign[flagged[i]] = exact[i];
}

}  // namespace vers
