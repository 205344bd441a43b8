// pairdist.cu — HNSW distance offload.
//
// The reference's HNSW keeps graph build and traversal on the CPU (hnsw.rs) and spends its time evaluating distances
// between one vector and a handful of neighbours: hnsw.rs:146 (heuristic neighbour selection), :258 and :273 (layer
// search) call Vector::cosine_similarity_simd, base.rs:225-294 holds its squared-euclidean twin.  These are NOT the
// scalar left-to-right sums of the IVF / LSH path: base.rs:158-223 sums in SIMD chunks,
//
//     res = 0
//     for every 64-wide chunk:  res += reduce_sum(u[64] * v[64])     (std::simd f32x64, base.rs:178-187)
//     for every  4-wide chunk of what is left:  res += reduce_sum(u[4] * v[4])        (base.rs:195-210)
//     for the last < 4 elements:                res += u[i] * v[i]                    (base.rs:213-221)
//     cosine: return 1.0 - res                                                        (base.rs:223)
//
// where reduce_sum is portable-simd's ORDERED reduction (core::intrinsics::simd::simd_reduce_add_ordered: the lanes are
// added left to right onto the start value; the start value is 0.0 or -0.0 depending on the toolchain's age, which
// cannot change `res`: x + p = p for the first product either way, and an all-(-0.0) chunk is absorbed by res = +0.0).
// So every chunk is an independent sequential chain and the chains are combined sequentially: a batch of (query, row)
// pairs maps to G lanes per pair with lane = chunk, then one short serial fold — bit-identical to the reference.
//
// The traversal itself stays on the host (north-star); a host that walks many queries at once (or evaluates a whole
// candidate frontier) hands the (query, neighbour id) pairs of a step to vers_pair_distances_simd[_dev].
#include <algorithm>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace vers {

enum { PD_L2SQ = 0, PD_DOT = 1 };

template <int OP>  // PD_L2SQ: (u - v)^2, PD_DOT: u * v — separately rounded like the reference (no FMA)
__device__ __forceinline__ float pd_term(float u, float v) {
    if (OP == PD_L2SQ) {
        const float t = __fsub_rn(u, v);
        return __fmul_rn(t, t);
    }
    return __fmul_rn(u, v);
}

// G lanes per pair (G = the power of two >= the number of 64-wide chunks, so 32 / G pairs share a warp and every lane
// has a chunk to chew on: dim 300 -> G = 4, 8 pairs per warp; dim 768 -> G = 16).  Lane `sub` of a group owns the 64-wide
// chunks sub, sub + G, ... and the 4-wide chunks sub, sub + G, ...; every chunk is its own sequential chain
// (reduce_sum is ordered), the chunk sums are folded into res in chunk order by every lane of the group.
// rows: [n][ld] fp32 (pad columns are never read: the chunking follows dim like the reference's const N);
// queries: [nq][q_ld]; pair_query == null: every pair uses query 0.
template <int OP, int G>
__global__ void __launch_bounds__(256)
    pair_distances_simd_kernel(const float* __restrict__ rows, uint64_t n, uint32_t ld, uint32_t dim, uint64_t id_base,
                               const float* __restrict__ queries, uint32_t nq, uint32_t q_ld,
                               const uint32_t* __restrict__ pair_query, const uint64_t* __restrict__ pair_row,
                               uint64_t n_pairs, float* __restrict__ out, uint32_t* __restrict__ bad) {
    constexpr int PPW = 32 / G;  // pairs per warp
    const int lane = threadIdx.x & 31, grp = lane / G, sub = lane % G, lane0 = grp * G;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t n64 = dim / 64, rem = dim - n64 * 64, n4 = rem / 4, tail0 = n64 * 64 + n4 * 4;
    const uint64_t n_iter = (n_pairs + PPW - 1) / PPW;
    for (uint64_t it = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_iter; it += warps) {
        const uint64_t pi = it * PPW + grp;
        const bool have = pi < n_pairs;
        uint64_t rid = 0;
        uint32_t qi = 0;
        bool ok = true;
        if (have) {
            rid = pair_row[pi] - id_base;  // wraps for ids below id_base: caught by the range test
            qi = pair_query ? pair_query[pi] : 0u;
            ok = rid < n && qi < nq;  // else: id_to_vec.get(..).unwrap() would panic (hnsw.rs:133, 270)
            if (!ok) rid = 0, qi = 0;
        }
        const float* u = queries + (uint64_t)qi * q_ld;
        const float* v = rows + rid * ld;
        float res = 0.0f;
        // 64-wide chunks (ld % 4 == 0 and 64-element chunk starts keep the float4 loads aligned)
        for (uint32_t c0 = 0; c0 < n64; c0 += G) {
            const uint32_t c = c0 + sub;
            float s = 0.0f;
            if (c < n64) {
                const float4* u4 = reinterpret_cast<const float4*>(u + (size_t)c * 64);
                const float4* v4 = reinterpret_cast<const float4*>(v + (size_t)c * 64);
#pragma unroll 8
                for (int i = 0; i < 16; ++i) {
                    const float4 a = u4[i], b = v4[i];
                    s = __fadd_rn(s, pd_term<OP>(a.x, b.x));
                    s = __fadd_rn(s, pd_term<OP>(a.y, b.y));
                    s = __fadd_rn(s, pd_term<OP>(a.z, b.z));
                    s = __fadd_rn(s, pd_term<OP>(a.w, b.w));
                }
            }
            const uint32_t live = min((uint32_t)G, n64 - c0);  // warp-uniform
            for (uint32_t l = 0; l < live; ++l) res = __fadd_rn(res, __shfl_sync(FULL_MASK, s, lane0 + (int)l));  // res += chunk
        }
        // 4-wide chunks of the remainder (at most 15)
        for (uint32_t c0 = 0; c0 < n4; c0 += G) {
            const uint32_t c = c0 + sub;
            float s = 0.0f;
            if (c < n4) {
                const float4 a = *reinterpret_cast<const float4*>(u + n64 * 64 + c * 4);
                const float4 b = *reinterpret_cast<const float4*>(v + n64 * 64 + c * 4);
                s = __fadd_rn(s, pd_term<OP>(a.x, b.x));
                s = __fadd_rn(s, pd_term<OP>(a.y, b.y));
                s = __fadd_rn(s, pd_term<OP>(a.z, b.z));
                s = __fadd_rn(s, pd_term<OP>(a.w, b.w));
            }
            const uint32_t live = min((uint32_t)G, n4 - c0);
            for (uint32_t l = 0; l < live; ++l) res = __fadd_rn(res, __shfl_sync(FULL_MASK, s, lane0 + (int)l));
        }
        for (uint32_t i = tail0; i < dim; ++i) res = __fadd_rn(res, pd_term<OP>(u[i], v[i]));  // < 4 scalars, every lane
        if (have && sub == 0) {
            if (ok) {
                out[pi] = OP == PD_DOT ? __fsub_rn(1.0f, res) : res;
            } else {
                out[pi] = __int_as_float(0x7fc00000);
                atomicAdd(bad, 1u);
            }
        }
    }
}

template <int OP>
static void pair_distances_dispatch(int G, unsigned grid, cudaStream_t st, const float* rows, uint64_t n, uint32_t ld,
                                    uint32_t dim, uint64_t id_base, const float* q, uint32_t nq, uint32_t q_ld,
                                    const uint32_t* pq, const uint64_t* pr, uint64_t np, float* out, uint32_t* bad) {
#define VERS_PD(GG)                                                                                                     \
    pair_distances_simd_kernel<OP, GG><<<grid, 256, 0, st>>>(rows, n, ld, dim, id_base, q, nq, q_ld, pq, pr, np, out, bad)
    switch (G) {
        case 1: VERS_PD(1); break;
        case 2: VERS_PD(2); break;
        case 4: VERS_PD(4); break;
        case 8: VERS_PD(8); break;
        case 16: VERS_PD(16); break;
        default: VERS_PD(32); break;
    }
#undef VERS_PD
}

static int32_t pair_distances_launch(vers_dataset* ds, const float* d_queries, uint32_t nq, uint32_t q_ld,
                                     const uint32_t* d_pair_query, const uint64_t* d_pair_row, uint64_t n_pairs,
                                     uint32_t metric, float* d_out, uint32_t* d_bad) {
    vers_ctx* ctx = ds->ctx;
    // lanes per pair: enough for the 64-wide chunks (dims below 64 only have 4-wide chunks: 2 lanes share them)
    const uint32_t n64 = ds->dim / 64;
    int G = n64 ? 1 : 2;
    while ((uint32_t)G < n64 && G < 32) G <<= 1;
    const uint64_t iters = ceil_div(n_pairs, (uint64_t)(32 / G));
    const unsigned grid = (unsigned)std::min<uint64_t>(ceil_div(iters, 8), (uint64_t)ctx->sm_count * 8);
    if (metric == VERS_METRIC_L2SQ)
        pair_distances_dispatch<PD_L2SQ>(G, grid, ctx->stream, ds->d_rows, ds->n, ds->ld, ds->dim, ds->id_base, d_queries, nq,
                                         q_ld, d_pair_query, d_pair_row, n_pairs, d_out, d_bad);
    else
        pair_distances_dispatch<PD_DOT>(G, grid, ctx->stream, ds->d_rows, ds->n, ds->ld, ds->dim, ds->id_base, d_queries, nq,
                                        q_ld, d_pair_query, d_pair_row, n_pairs, d_out, d_bad);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

}  // namespace vers

using namespace vers;

extern "C" int32_t vers_pair_distances_simd_dev(vers_dataset* ds, const float* d_queries, uint32_t nq,
                                                uint32_t q_stride_floats, const uint32_t* d_pair_query,
                                                const uint64_t* d_pair_row, uint64_t n_pairs, uint32_t metric,
                                                float* d_out, uint32_t* d_bad_count) {
    if (!ds || (!d_queries && nq) || (!d_pair_row && n_pairs) || (!d_out && n_pairs) || !d_bad_count)
        return fail(VERS_ERR_ARG, "pair_distances_simd_dev: null argument");
    if (metric > VERS_METRIC_COSINE) return fail(VERS_ERR_ARG, "pair_distances_simd_dev: unknown metric %u", metric);
    if (q_stride_floats < ds->dim || (q_stride_floats & 3u) || (reinterpret_cast<uintptr_t>(d_queries) & 15u))
        return fail(VERS_ERR_ARG, "pair_distances_simd_dev: queries must be 16-byte aligned with a stride that is a "
                                  "multiple of 4 floats and >= dim");
    if (n_pairs == 0) return VERS_OK;
    if (nq == 0) return fail(VERS_ERR_ARG, "pair_distances_simd_dev: pairs without queries");
    std::lock_guard<std::recursive_mutex> lk(ds->ctx->mu);
    VERS_CUDA(cudaSetDevice(ds->ctx->device));
    return pair_distances_launch(ds, d_queries, nq, q_stride_floats, d_pair_query, d_pair_row, n_pairs, metric, d_out,
                                 d_bad_count);
}

extern "C" int32_t vers_pair_distances_simd(vers_dataset* ds, const float* queries, uint32_t nq, uint32_t q_stride_floats,
                                            const uint32_t* pair_query, const uint64_t* pair_row, uint64_t n_pairs,
                                            uint32_t metric, float* out) {
    if (!ds || (!queries && nq) || (!pair_row && n_pairs) || (!out && n_pairs))
        return fail(VERS_ERR_ARG, "pair_distances_simd: null argument");
    if (metric > VERS_METRIC_COSINE) return fail(VERS_ERR_ARG, "pair_distances_simd: unknown metric %u", metric);
    if (q_stride_floats < ds->dim) return fail(VERS_ERR_ARG, "pair_distances_simd: query stride < dim");
    if (n_pairs == 0) return VERS_OK;
    if (nq == 0) return fail(VERS_ERR_ARG, "pair_distances_simd: pairs without queries");
    vers_ctx* ctx = ds->ctx;
    VERS_CUDA(cudaSetDevice(ctx->device));
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    ScratchCarver plan(nullptr);
    plan.plan<float>((size_t)nq * ds->ld);
    plan.plan<uint64_t>(n_pairs);
    plan.plan<uint32_t>(pair_query ? n_pairs : 1);
    plan.plan<float>(n_pairs);
    plan.plan<uint32_t>(1);
    VERS_TRY(io_reserve(ctx, plan.off + 256));
    ScratchCarver io(ctx->io);
    float* d_q = io.take<float>((size_t)nq * ds->ld);
    uint64_t* d_row = io.take<uint64_t>(n_pairs);
    uint32_t* d_pq = io.take<uint32_t>(pair_query ? n_pairs : 1);
    float* d_out = io.take<float>(n_pairs);
    uint32_t* d_bad = io.take<uint32_t>(1);
    if (ds->ld != ds->dim) VERS_CUDA(cudaMemsetAsync(d_q, 0, (size_t)nq * ds->ld * 4, ctx->stream));
    VERS_CUDA(cudaMemcpy2DAsync(d_q, (size_t)ds->ld * 4, queries, (size_t)q_stride_floats * 4, (size_t)ds->dim * 4, nq,
                                cudaMemcpyHostToDevice, ctx->stream));
    VERS_CUDA(cudaMemcpyAsync(d_row, pair_row, n_pairs * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (pair_query) VERS_CUDA(cudaMemcpyAsync(d_pq, pair_query, n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    VERS_CUDA(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    VERS_TRY(pair_distances_launch(ds, d_q, nq, ds->ld, pair_query ? d_pq : nullptr, d_row, n_pairs, metric, d_out, d_bad));
    uint32_t bad = 0;
    VERS_CUDA(cudaMemcpyAsync(out, d_out, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad)
        return fail(VERS_ERR_PANIC, "pair_distances_simd: %u pairs name a row or query that does not exist "
                                    "(id_to_vec.get(..).unwrap() panics, hnsw.rs:133, 270)", bad);
    return VERS_OK;
}
