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
            tc::mbar_init(&conv[s], KA_CONV_WARPS);
            tc::mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tfull[b], 1);
            tc::mbar_init(&tempty[b], KA_EPI_WARPS);
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
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[buf]);
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
                    if (lane == 0) tc::mbar_arrive(&conv[stage]);
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
