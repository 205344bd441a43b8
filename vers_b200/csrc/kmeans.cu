// kmeans.cu — IVFFlatIndex k-means (indexes/ivfflat.rs:18-100, 138-149) on one GPU's row shard.
//   assign  : exact-order pairwise tiles + per-row first-minimum argmin                 (ivfflat.rs:29-46)
//   sums    : rows grouped by cluster in ascending row order, one thread per (cluster, 4 dims) walks its members
//             left to right so the association is the reference's `sums[c] = sums[c].add(x)` (ivfflat.rs:52-55);
//             the running sums are an in/out argument so row shards can be chained in row order across GPUs
//   finalize: sum / (count as f32), zero vector for empty clusters, bitwise convergence test (ivfflat.rs:57-93)
//   cost    : per-row exact distance, then a strictly sequential fold in row order     (ivfflat.rs:138-149)
#include <cub/device/device_radix_sort.cuh>

#include "kmeans.cuh"
#include "scan.cuh"
#include "kmeans_tc.cuh"

namespace vers {

// ---------------------------------------------------------------- assign
template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, 2)
    assign_kernel(RowSrc A, RowSrc B, uint32_t ld, uint32_t* __restrict__ out_assign, uint32_t* __restrict__ bad) {
    extern __shared__ __align__(16) float smem[];
    constexpr int MA = Cfg::MA, MB = Cfg::MB, NTA = Cfg::NTA, NTB = Cfg::NTB, TA = Cfg::TA, TB = Cfg::TB;
    const int tid = threadIdx.x, ta = tid % NTA, tb = tid / NTA;
    const uint64_t a0 = (uint64_t)blockIdx.x * TA;
    float best_d[MA];
    uint32_t best_c[MA];
#pragma unroll
    for (int i = 0; i < MA; ++i) {
        best_d[i] = __int_as_float(0x7f800000);
        best_c[i] = 0xffffffffu;
    }
    for (uint64_t b0 = 0; b0 < B.n; b0 += TB) {
        float acc[MA][MB];
        tile_compute<Cfg, OP_L2SQ>(acc, A, a0, B, b0, ld, smem);
#pragma unroll
        for (int j = 0; j < MB; ++j) {
            uint64_t c = b0 + (uint64_t)(tb + j * NTB);
            if (c < B.n) {
#pragma unroll
                for (int i = 0; i < MA; ++i) {
                    if (entry_less<uint32_t>(acc[i][j], (uint32_t)c, best_d[i], best_c[i])) {
                        best_d[i] = acc[i][j];
                        best_c[i] = (uint32_t)c;
                    }
                }
            }
        }
    }
    // tile_compute ended with __syncthreads(): the staging buffers are free, reuse them for the row reduction
    float* red_d = smem;
    uint32_t* red_c = reinterpret_cast<uint32_t*>(smem + TA * NTB);
#pragma unroll
    for (int i = 0; i < MA; ++i) {
        int r = ta + i * NTA;
        red_d[r * NTB + tb] = best_d[i];
        red_c[r * NTB + tb] = best_c[i];
    }
    __syncthreads();
    for (int r = tid; r < TA; r += Cfg::NT) {
        float bd = red_d[r * NTB];
        uint32_t bc = red_c[r * NTB];
        for (int t = 1; t < NTB; ++t) {
            float d = red_d[r * NTB + t];
            uint32_t c = red_c[r * NTB + t];
            if (entry_less<uint32_t>(d, c, bd, bc)) {
                bd = d;
                bc = c;
            }
        }
        if (a0 + r < A.n) {
            out_assign[a0 + r] = bc;
            // every distance of this row compared false (NaN, or inf - inf): the reference panics on
            // partial_cmp(..).unwrap() (ivfflat.rs:39); flag it instead of handing out an index past the table
            if (bc == 0xffffffffu && bad) atomicOr(bad, 1u);
        }
    }
}

template <class Cfg>
static int32_t launch_assign(vers_ctx* ctx, const RowSrc& A, const RowSrc& B, uint32_t ld, uint32_t* d_assign,
                             uint32_t* d_bad) {
    static_assert(Cfg::TILE_FLOATS >= Cfg::TA * Cfg::NTB * 2, "reduction fits in the staging buffers");
    auto kern = assign_kernel<Cfg>;
    size_t smem = (size_t)Cfg::TILE_FLOATS * 4;
    VERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)ceil_div(A.n, Cfg::TA), Cfg::NT, smem, ctx->stream>>>(A, B, ld, d_assign, d_bad);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

int32_t kmeans_assign_rows(vers_ctx* ctx, const RowSrc& rows, const float* d_cents, uint32_t C, uint32_t ld,
                           uint32_t* d_assign, int family, uint32_t* d_bad) {
    if (rows.n == 0) return VERS_OK;
    RowSrc B{d_cents, nullptr, ld, C};
    FamilyTimer ft(ctx, family);
    if (C <= 8) return launch_assign<NarrowCfg>(ctx, rows, B, ld, d_assign, d_bad);
    return launch_assign<WideCfg>(ctx, rows, B, ld, d_assign, d_bad);
}

// reads the "a row compared false against every centroid" flag of the assign kernels (synchronises)
static int32_t kmeans_check_bad(vers_kmeans* km) {
    uint32_t bad = 0;
    VERS_CUDA(cudaMemcpyAsync(&bad, km->d_bad, 4, cudaMemcpyDeviceToHost, km->ds->ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(km->ds->ctx->stream));
    if (bad)
        return fail(VERS_ERR_PANIC, "assign_to_clusters: a distance is NaN (partial_cmp(..).unwrap() panics, "
                                    "ivfflat.rs:39)");
    return VERS_OK;
}

// ---------------------------------------------------------------- CSR (rows grouped by cluster, stable)
__global__ void iota_kernel(uint32_t* p, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = (uint32_t)i;
}
__global__ void hist_kernel(const uint32_t* __restrict__ assign, uint64_t n, uint32_t* hist) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[assign[i]], 1u);
}

int32_t kmeans_build_csr(vers_kmeans* km) {
    if (km->csr_valid) return VERS_OK;
    vers_ctx* ctx = km->ds->ctx;
    const uint64_t n = km->ds->n;
    VERS_CUDA(cudaMemsetAsync(km->d_hist, 0, sizeof(uint32_t) * km->C, ctx->stream));
    if (n) {
        hist_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(km->d_assign, n, km->d_hist);
        VERS_LAUNCH_CHECK(ctx);
    }
    VERS_TRY(launch_exclusive_scan(ctx, km->d_hist, km->C, km->d_off));
    if (n) {
        int end_bit = 1;
        while ((1ull << end_bit) < km->C) ++end_bit;
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, km->d_assign, km->d_sorted_keys, km->d_iota, km->d_sorted_rows,
                                        (int64_t)n, 0, end_bit, ctx->stream);
        if (need > km->cub_bytes) {
            VERS_CUDA(cudaStreamSynchronize(ctx->stream));
            if (km->d_cub) cudaFree(km->d_cub);
            km->d_cub = nullptr;
            VERS_CUDA(cudaMalloc(&km->d_cub, need));
            km->cub_bytes = need;
        }
        size_t bytes = km->cub_bytes;
        // LSD radix sort is stable: inside one cluster the rows stay in ascending row order
        VERS_CUDA(cub::DeviceRadixSort::SortPairs(km->d_cub, bytes, km->d_assign, km->d_sorted_keys, km->d_iota,
                                                  km->d_sorted_rows, (int64_t)n, 0, end_bit, ctx->stream));
        ctx->launches += 1;
    }
    km->csr_valid = true;
    return VERS_OK;
}

// ---------------------------------------------------------------- ordered sums
// one warp = one cluster x 128 consecutive dims (32 lanes x float4); members walked in order, 8 loads in flight
__global__ void __launch_bounds__(256)
    sums_kernel(const float* __restrict__ rows, uint32_t ld, const uint32_t* __restrict__ sorted_rows,
                const uint64_t* __restrict__ off, uint32_t C, float* sums_io, uint64_t* counts_io) {
    const uint32_t ld4 = ld >> 2;
    const uint32_t groups = (ld4 + 31) / 32;
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (uint64_t)C * groups) return;
    const uint32_t c = (uint32_t)(w / groups), g = (uint32_t)(w % groups);
    const uint32_t col4 = g * 32 + lane;
    const uint64_t m0 = off[c], m1 = off[c + 1];
    if (g == 0 && lane == 0) counts_io[c] += (m1 - m0);
    if (col4 >= ld4) return;
    float4* sp = reinterpret_cast<float4*>(sums_io + (uint64_t)c * ld) + col4;
    float4 s = *sp;
    const float4* base = reinterpret_cast<const float4*>(rows) + col4;
    uint64_t m = m0;
    for (; m + 8 <= m1; m += 8) {
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = __ldg(base + (uint64_t)sorted_rows[m + u] * ld4);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s.x = __fadd_rn(s.x, x[u].x);
            s.y = __fadd_rn(s.y, x[u].y);
            s.z = __fadd_rn(s.z, x[u].z);
            s.w = __fadd_rn(s.w, x[u].w);
        }
    }
    for (; m < m1; ++m) {
        float4 x = __ldg(base + (uint64_t)sorted_rows[m] * ld4);
        s.x = __fadd_rn(s.x, x.x);
        s.y = __fadd_rn(s.y, x.y);
        s.z = __fadd_rn(s.z, x.z);
        s.w = __fadd_rn(s.w, x.w);
    }
    *sp = s;
}

__global__ void finalize_kernel(const float* __restrict__ sums, const uint64_t* __restrict__ counts,
                                const float* __restrict__ cur, float* __restrict__ next, uint32_t C, uint32_t dim,
                                uint32_t ld, uint32_t* flag) {
    uint64_t total = (uint64_t)C * ld;
    bool diff = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c = (uint32_t)(i / ld), d = (uint32_t)(i % ld);
        float v = 0.0f;
        uint64_t cnt = counts[c];
        if (d < dim && cnt > 0) v = __fdiv_rn(sums[i], __ull2float_rn(cnt));  // `count as f32`
        next[i] = v;
        if (__float_as_uint(v) != __float_as_uint(cur[i])) diff = true;  // to_hashkey compare, base.rs:113-117
    }
    if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(flag, 1u);
}

// ---------------------------------------------------------------- cost
__global__ void rowdist_kernel(const float* __restrict__ rows, uint64_t n, uint32_t ld,
                               const float* __restrict__ cents, const uint32_t* __restrict__ assign,
                               float* __restrict__ out) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float4* x = reinterpret_cast<const float4*>(rows + r * ld);
    const float4* c = reinterpret_cast<const float4*>(cents + (uint64_t)assign[r] * ld);
    float s = 0.0f;
    for (uint32_t i = 0; i < (ld >> 2); ++i) {
        float4 a = x[i], b = __ldg(c + i);
        float t;
        t = __fsub_rn(a.x, b.x); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(a.y, b.y); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(a.z, b.z); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(a.w, b.w); s = __fadd_rn(s, __fmul_rn(t, t));
    }
    out[r] = s;
}

// strictly sequential fold in row order (fold(0.0, |acc, v| acc + v), ivfflat.rs:148) by one warp:
// lanes fetch 32 values coalesced, every lane replays the same 32 dependent adds via shuffles
__global__ void __launch_bounds__(32) seqsum_kernel(const float* __restrict__ v, uint64_t n, float* acc_io) {
    const int lane = threadIdx.x;
    float s = *acc_io;
    float cur = lane < n ? v[lane] : 0.0f;
    for (uint64_t base = 0; base < n; base += 32) {
        uint64_t nx = base + 32 + lane;
        float nxt = nx < n ? v[nx] : 0.0f;
        int cnt = (int)min((uint64_t)32, n - base);
        if (cnt == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) s = __fadd_rn(s, __shfl_sync(FULL_MASK, cur, j));
        } else {
            for (int j = 0; j < cnt; ++j) s = __fadd_rn(s, __shfl_sync(FULL_MASK, cur, j));
        }
        cur = nxt;
    }
    if (lane == 0) *acc_io = s;
}

__global__ void gather_rows_kernel(const float* __restrict__ rows, uint32_t ld, const uint64_t* __restrict__ idx,
                                   uint32_t cnt, float* __restrict__ out) {
    uint64_t total = (uint64_t)cnt * ld;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t j = i / ld, d = i % ld;
        out[i] = rows[idx[j] * ld + d];
    }
}

__global__ void widen_assign_kernel(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}
__global__ void narrow_assign_kernel(const uint64_t* __restrict__ in, uint64_t n, uint32_t C, uint32_t* __restrict__ out,
                                     uint32_t* bad) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = in[i];
        if (v >= C) {
            atomicOr(bad, 1u);
            v = 0;
        }
        out[i] = (uint32_t)v;
    }
}

}  // namespace vers

using namespace vers;

extern "C" int32_t vers_kmeans_free(vers_kmeans* km) {
    if (!km) return VERS_OK;
    cudaSetDevice(km->ds->ctx->device);
    cudaStreamSynchronize(km->ds->ctx->stream);
    cudaFree(km->d_cents);
    cudaFree(km->d_next);
    cudaFree(km->d_assign);
    cudaFree(km->d_sorted_rows);
    cudaFree(km->d_sorted_keys);
    cudaFree(km->d_iota);
    cudaFree(km->d_hist);
    cudaFree(km->d_off);
    cudaFree(km->d_sums);
    cudaFree(km->d_counts);
    cudaFree(km->d_rowdist);
    cudaFree(km->d_flag);
    cudaFree(km->d_bad);
    cudaFree(km->d_cub);
    cudaFree(km->d_row_norm);
    cudaFree(km->d_cent_norm);
    cudaFree(km->d_cent_hi);
    cudaFree(km->d_cent_lo);
    cudaFree(km->d_cent_tiles);
    cudaFree(km->d_ncmax);
    cudaFree(km->d_nxmax);
    cudaFree(km->d_flagged);
    cudaFree(km->d_nflagged);
    cudaFree(km->d_exact);
    delete km;
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_create(vers_dataset* ds, uint32_t num_clusters, vers_kmeans** out) {
    if (!ds || !out) return fail(VERS_ERR_ARG, "kmeans_create: null argument");
    *out = nullptr;
    if (num_clusters == 0) return fail(VERS_ERR_PANIC, "kmeans_create: 0 clusters (min_by on empty iterator, ivfflat.rs:42)");
    if (ds->n >= 0xffffffffull) return fail(VERS_ERR_UNSUPPORTED, "more than 2^32-2 rows per GPU shard");
    vers_ctx* ctx = ds->ctx;
    VERS_CUDA(cudaSetDevice(ctx->device));
    vers_kmeans* km = new vers_kmeans();
    km->ds = ds;
    km->C = num_clusters;
    const size_t cl = (size_t)num_clusters * ds->ld, n1 = ds->n ? ds->n : 1;
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    };
    A((void**)&km->d_cents, cl * 4);
    A((void**)&km->d_next, cl * 4);
    A((void**)&km->d_sums, cl * 4);
    A((void**)&km->d_counts, (size_t)num_clusters * 8);
    A((void**)&km->d_assign, n1 * 4);
    A((void**)&km->d_sorted_rows, n1 * 4);
    A((void**)&km->d_sorted_keys, n1 * 4);
    A((void**)&km->d_iota, n1 * 4);
    A((void**)&km->d_hist, (size_t)num_clusters * 4);
    A((void**)&km->d_off, ((size_t)num_clusters + 1) * 8);
    A((void**)&km->d_flag, 4);
    A((void**)&km->d_bad, 4);
    if (e != cudaSuccess) {
        vers_kmeans_free(km);
        return fail(VERS_ERR_NOMEM, "kmeans_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaMemsetAsync(km->d_cents, 0, cl * 4, ctx->stream));
    if (ds->n) {
        iota_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(km->d_iota, ds->n);
        VERS_LAUNCH_CHECK(ctx);
    }
    *out = km;
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_init_from_rows(vers_kmeans* km, const uint64_t* init_rows) {
    if (!km || !init_rows) return fail(VERS_ERR_ARG, "kmeans_init_from_rows: null argument");
    vers_dataset* ds = km->ds;
    vers_ctx* ctx = ds->ctx;
    if (ds->n == 0) return fail(VERS_ERR_PANIC, "kmeans init on an empty dataset (gen_range(0..0), ivfflat.rs:23)");
    for (uint32_t j = 0; j < km->C; ++j)
        if (init_rows[j] >= ds->n) return fail(VERS_ERR_ARG, "kmeans_init_from_rows: row %llu out of range",
                                               (unsigned long long)init_rows[j]);
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    uint64_t* d_idx = nullptr;
    VERS_CUDA(cudaMalloc(&d_idx, (size_t)km->C * 8));
    cudaError_t e = cudaMemcpyAsync(d_idx, init_rows, (size_t)km->C * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        gather_rows_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ds->d_rows, ds->ld, d_idx, km->C, km->d_cents);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_idx);
    if (e != cudaSuccess) return fail(VERS_ERR_CUDA, "kmeans_init_from_rows: %s", cudaGetErrorString(e));
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_set_centroids(vers_kmeans* km, const float* centroids, uint32_t stride_floats) {
    if (!km || !centroids) return fail(VERS_ERR_ARG, "kmeans_set_centroids: null argument");
    vers_dataset* ds = km->ds;
    if (stride_floats < ds->dim) return fail(VERS_ERR_ARG, "kmeans_set_centroids: stride < dim");
    VERS_CUDA(cudaSetDevice(ds->ctx->device));
    VERS_CUDA(cudaMemsetAsync(km->d_cents, 0, (size_t)km->C * ds->ld * 4, ds->ctx->stream));
    VERS_CUDA(cudaMemcpy2DAsync(km->d_cents, (size_t)ds->ld * 4, centroids, (size_t)stride_floats * 4,
                                (size_t)ds->dim * 4, km->C, cudaMemcpyHostToDevice, ds->ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ds->ctx->stream));
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_get_centroids(vers_kmeans* km, float* centroids, uint32_t stride_floats) {
    if (!km || !centroids) return fail(VERS_ERR_ARG, "kmeans_get_centroids: null argument");
    vers_dataset* ds = km->ds;
    if (stride_floats < ds->dim) return fail(VERS_ERR_ARG, "kmeans_get_centroids: stride < dim");
    VERS_CUDA(cudaSetDevice(ds->ctx->device));
    VERS_CUDA(cudaMemcpy2DAsync(centroids, (size_t)stride_floats * 4, km->d_cents, (size_t)ds->ld * 4,
                                (size_t)ds->dim * 4, km->C, cudaMemcpyDeviceToHost, ds->ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ds->ctx->stream));
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_get_assignments(vers_kmeans* km, uint64_t* assignments) {
    if (!km || (!assignments && km->ds->n)) return fail(VERS_ERR_ARG, "kmeans_get_assignments: null argument");
    vers_ctx* ctx = km->ds->ctx;
    const uint64_t n = km->ds->n;
    if (n == 0) return VERS_OK;
    VERS_CUDA(cudaSetDevice(ctx->device));
    uint64_t* d_wide = nullptr;
    VERS_CUDA(cudaMalloc(&d_wide, n * 8));
    widen_assign_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(km->d_assign, n, d_wide);
    ctx->launches += 1;
    cudaError_t e = cudaMemcpyAsync(assignments, d_wide, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_wide);
    if (e != cudaSuccess) return fail(VERS_ERR_CUDA, "kmeans_get_assignments: %s", cudaGetErrorString(e));
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_centroids_device_ptr(vers_kmeans* km, void** ptr, uint32_t* ld) {
    if (!km || !ptr) return fail(VERS_ERR_ARG, "kmeans_centroids_device_ptr: null argument");
    *ptr = km->d_cents;
    if (ld) *ld = km->ds->ld;
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_assign_device_ptr(vers_kmeans* km, void** ptr) {
    if (!km || !ptr) return fail(VERS_ERR_ARG, "kmeans_assign_device_ptr: null argument");
    *ptr = km->d_assign;
    return VERS_OK;
}

// tensor-core candidate argmin + certificate; uncertified rows redone by the exact-order kernel.
// ld <= 128 (mode 0 / 3): tc_assign1_kernel (one MMA per K step — kind::f16 by default, kind::tf32 in mode 3 — rows
// resident in tensor memory, top-4 + exact rerank inside the kernel); otherwise / mode 2: tc_assign_kernel (split
// precision, three MMAs per K step, top-2 gap).
static int32_t kmeans_tc_buffers(vers_kmeans* km) {
    vers_dataset* ds = km->ds;
    vers_ctx* ctx = ds->ctx;
    cudaStream_t s = ctx->stream;
    if (km->d_row_norm && km->norm_epoch == ds->epoch) return VERS_OK;
    if (!km->d_row_norm) {
        // all or nothing: a failed allocation must not leave a half-initialised state behind
        float *row_norm = nullptr, *cent_norm = nullptr, *cent_hi = nullptr, *cent_lo = nullptr, *cent_tiles = nullptr;
        uint32_t *ncmax = nullptr, *nxmax = nullptr, *flagged = nullptr, *nflagged = nullptr, *exact = nullptr;
        cudaError_t e = cudaSuccess;
        auto A = [&](void** p, size_t bytes) {
            if (e == cudaSuccess) e = cudaMalloc(p, bytes);
        };
        A((void**)&row_norm, ds->n * 4);
        A((void**)&cent_norm, ((size_t)km->C + KA_N) * 4);
        A((void**)&cent_hi, (size_t)km->C * ds->ld * 4);
        A((void**)&cent_lo, (size_t)km->C * ds->ld * 4);
        if (ds->ld <= K1_MAX_KCH * K1_KC)  // image of the centroid tiles for tc_assign1_kernel
            A((void**)&cent_tiles, (size_t)ceil_div(km->C, K1_N) * ((ds->ld + K1_KC - 1) / K1_KC) * K1_BOX_BYTES);
        A((void**)&ncmax, 4);
        A((void**)&nxmax, 8);
        A((void**)&flagged, ds->n * 4);
        A((void**)&nflagged, 4);
        A((void**)&exact, ds->n * 4);
        if (e == cudaSuccess) {  // +inf past C: the epilogue reads whole tiles of norms; a padded column can never win
            std::vector<float> inf(KA_N, __builtin_inff());
            e = cudaMemcpyAsync(cent_norm + km->C, inf.data(), KA_N * 4, cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        }
        if (e != cudaSuccess) {
            cudaFree(row_norm), cudaFree(cent_norm), cudaFree(cent_hi), cudaFree(cent_lo), cudaFree(cent_tiles);
            cudaFree(ncmax), cudaFree(nxmax), cudaFree(flagged), cudaFree(nflagged), cudaFree(exact);
            return fail(e == cudaErrorMemoryAllocation ? VERS_ERR_NOMEM : VERS_ERR_CUDA, "kmeans assign buffers: %s",
                        cudaGetErrorString(e));
        }
        km->d_row_norm = row_norm, km->d_cent_norm = cent_norm, km->d_cent_hi = cent_hi, km->d_cent_lo = cent_lo;
        km->d_cent_tiles = cent_tiles;
        km->d_ncmax = ncmax, km->d_nxmax = nxmax, km->d_flagged = flagged, km->d_nflagged = nflagged, km->d_exact = exact;
    }
    // (re)computed whenever the rows changed since (vers_dataset_normalize bumps ds->epoch)
    VERS_CUDA(cudaMemsetAsync(km->d_nxmax, 0, 8, s));
    sqnorm_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(ds->d_rows, ds->ld, ds->n, km->d_row_norm, km->d_nxmax);
    VERS_LAUNCH_CHECK(ctx);
    // the fp16 kernel's scale: |x_i| <= ||x|| < 2^ex  =>  |x_i| 2^(14 - ex) < 2^14 for every row, and for every centroid
    // that is a mean of rows (user-set centroids beyond that are caught by the image's overflow count)
    uint32_t nxbits = 0;
    VERS_CUDA(cudaMemcpyAsync(&nxbits, km->d_nxmax, 4, cudaMemcpyDeviceToHost, s));
    VERS_CUDA(cudaStreamSynchronize(s));
    float nxmax;
    memcpy(&nxmax, &nxbits, 4);
    int ex = 0;
    if (nxmax > 0.0f && std::isfinite(nxmax)) (void)std::frexp(std::sqrt((double)nxmax) * 1.0001, &ex);
    km->f16_scale = std::ldexp(1.0f, std::max(-60, std::min(60, 14 - ex)));
    km->norm_epoch = ds->epoch;
    return VERS_OK;
}

static int32_t kmeans_assign_tc(vers_kmeans* km) {
    vers_dataset* ds = km->ds;
    vers_ctx* ctx = ds->ctx;
    cudaStream_t s = ctx->stream;
    VERS_TRY(kmeans_tc_buffers(km));
    const bool tf32_first = (km->mode == 0 || km->mode == 3) && ds->ld <= K1_MAX_KCH * K1_KC;
    const bool f16 = tf32_first && km->mode == 0;  // default: the same kernel on kind::f16 (mode 3 keeps kind::tf32)
    VERS_CUDA(cudaMemsetAsync(km->d_ncmax, 0, 4, s));
    VERS_CUDA(cudaMemsetAsync(km->d_nflagged, 0, 4, s));
    VERS_CUDA(cudaMemsetAsync(km->d_bad, 0, 4, s));
    sqnorm_kernel<<<(unsigned)ceil_div((uint64_t)km->C * 32, 256), 256, 0, s>>>(km->d_cents, ds->ld, km->C,
                                                                               km->d_cent_norm, km->d_ncmax);
    VERS_LAUNCH_CHECK(ctx);
    if (tf32_first) {
        const uint32_t nk1 = (ds->ld + K1_KC - 1) / K1_KC;
        if (f16) {
            VERS_CUDA(cudaMemsetAsync(km->d_nxmax + 1, 0, 4, s));
            tile_image_f16_kernel<<<ctx->sm_count * 2, 256, 0, s>>>(km->d_cents, km->C, ds->ld, (nk1 + 1) / 2, km->f16_scale,
                                                                   reinterpret_cast<uint4*>(km->d_cent_tiles),
                                                                   km->d_nxmax + 1);
        } else {
            tile_image_tf32_kernel<<<ctx->sm_count * 2, 256, 0, s>>>(km->d_cents, km->C, ds->ld, nk1, km->d_cent_tiles);
        }
        VERS_LAUNCH_CHECK(ctx);
        CUtensorMap tm_rows;
        VERS_TRY(make_tmap_2d_f32(&tm_rows, ds->d_rows, ds->n, ds->ld, ds->ld, K1_M, K1_KC));
        TcAssign1Params p;
        p.n_rows = ds->n;
        p.C = km->C;
        p.ld = ds->ld;
        p.rows = ds->d_rows;
        p.cents = km->d_cents;
        p.cent_tiles = km->d_cent_tiles;
        p.row_norm = km->d_row_norm;
        p.cent_norm = km->d_cent_norm;
        p.ncmax_bits = km->d_ncmax;
        p.assign = km->d_assign;
        p.flagged = km->d_flagged;
        p.n_flagged = km->d_nflagged;
        p.scale = km->f16_scale;
        p.key_scale = -2.0f / (km->f16_scale * km->f16_scale);
        p.f16_bad = km->d_nxmax + 1;
        void (*kern)(CUtensorMap, TcAssign1Params) = nullptr;
        switch ((ds->ld + K1_KC - 1) / K1_KC) {
            case 1: kern = f16 ? tc_assign1_kernel<1, true> : tc_assign1_kernel<1, false>; break;
            case 2: kern = f16 ? tc_assign1_kernel<2, true> : tc_assign1_kernel<2, false>; break;
            case 3: kern = f16 ? tc_assign1_kernel<3, true> : tc_assign1_kernel<3, false>; break;
            default: kern = f16 ? tc_assign1_kernel<4, true> : tc_assign1_kernel<4, false>; break;
        }
        VERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES));
        FamilyTimer ft(ctx, KF_ASSIGN);
        const uint64_t nrbp = ceil_div(ds->n, 2 * K1_M);
        kern<<<(unsigned)std::min<uint64_t>(nrbp, ctx->sm_count), K1_THREADS, K1_SMEM_BYTES, s>>>(tm_rows, p);
        VERS_LAUNCH_CHECK(ctx);
    } else {
        split_tf32_kernel<<<ctx->sm_count * 2, 256, 0, s>>>(km->d_cents, (uint64_t)km->C * (ds->ld >> 2), km->d_cent_hi,
                                                            km->d_cent_lo);
        VERS_LAUNCH_CHECK(ctx);
        CUtensorMap tm_rows, tm_chi, tm_clo;
        VERS_TRY(make_tmap_2d_f32(&tm_rows, ds->d_rows, ds->n, ds->ld, ds->ld, KA_M, KA_KC));
        VERS_TRY(make_tmap_2d_f32(&tm_chi, km->d_cent_hi, km->C, ds->ld, ds->ld, KA_N, KA_KC));
        VERS_TRY(make_tmap_2d_f32(&tm_clo, km->d_cent_lo, km->C, ds->ld, ds->ld, KA_N, KA_KC));
        TcAssignParams p;
        p.n_rows = ds->n;
        p.C = km->C;
        p.ld = ds->ld;
        p.row_norm = km->d_row_norm;
        p.cent_norm = km->d_cent_norm;
        p.ncmax_bits = km->d_ncmax;
        p.assign = km->d_assign;
        p.flagged = km->d_flagged;
        p.n_flagged = km->d_nflagged;
        VERS_CUDA(cudaFuncSetAttribute(tc_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KA_SMEM_BYTES));
        FamilyTimer ft(ctx, KF_ASSIGN);
        const uint64_t nrb = ceil_div(ds->n, KA_M);
        tc_assign_kernel<<<(unsigned)std::min<uint64_t>(nrb, ctx->sm_count), KA_THREADS, KA_SMEM_BYTES, s>>>(tm_rows,
                                                                                                          tm_chi, tm_clo, p);
        VERS_LAUNCH_CHECK(ctx);
    }
    uint32_t nf = 0;
    VERS_CUDA(cudaMemcpyAsync(&nf, km->d_nflagged, 4, cudaMemcpyDeviceToHost, s));
    VERS_CUDA(cudaStreamSynchronize(s));
    km->last_flagged = nf;
    if (nf) {
        RowSrc A{ds->d_rows, km->d_flagged, ds->ld, nf};
        VERS_TRY(kmeans_assign_rows(ctx, A, km->d_cents, km->C, ds->ld, km->d_exact, KF_LIST_SCAN, km->d_bad));  // own family: the redo
        scatter_assign_kernel<<<ctx->sm_count * 2, 256, 0, s>>>(km->d_flagged, km->d_nflagged, km->d_exact, km->d_assign);
        VERS_LAUNCH_CHECK(ctx);
        VERS_TRY(kmeans_check_bad(km));
    }
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_assign_step(vers_kmeans* km) {
    if (!km) return fail(VERS_ERR_ARG, "kmeans_assign_step: null");
    vers_dataset* ds = km->ds;
    std::lock_guard<std::recursive_mutex> lk(ds->ctx->mu);
    VERS_CUDA(cudaSetDevice(ds->ctx->device));
    km->csr_valid = false;
    if (km->mode != 1 && ds->ld >= KA_KC && ds->n >= 1 && ds->n < 0x7fffffffull) return kmeans_assign_tc(km);
    km->last_flagged = 0;
    VERS_CUDA(cudaMemsetAsync(km->d_bad, 0, 4, ds->ctx->stream));
    RowSrc A{ds->d_rows, nullptr, ds->ld, ds->n};
    VERS_TRY(kmeans_assign_rows(ds->ctx, A, km->d_cents, km->C, ds->ld, km->d_assign, KF_ASSIGN, km->d_bad));
    return kmeans_check_bad(km);
}

extern "C" int32_t vers_kmeans_set_mode(vers_kmeans* km, int32_t mode) {
    if (!km || mode < 0 || mode > 3) return fail(VERS_ERR_ARG, "kmeans_set_mode: bad argument");
    km->mode = mode;
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_last_assign_stats(vers_kmeans* km, uint64_t* uncertified_rows) {
    if (!km || !uncertified_rows) return fail(VERS_ERR_ARG, "kmeans_last_assign_stats: null");
    *uncertified_rows = km->last_flagged;
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_sums_step_dev(vers_kmeans* km, float* d_sums_io, uint64_t* d_counts_io) {
    if (!km || !d_sums_io || !d_counts_io) return fail(VERS_ERR_ARG, "kmeans_sums_step_dev: null argument");
    vers_dataset* ds = km->ds;
    vers_ctx* ctx = ds->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    VERS_TRY(kmeans_build_csr(km));
    const uint32_t groups = ((ds->ld >> 2) + 31) / 32;
    const uint64_t warps = (uint64_t)km->C * groups;
    FamilyTimer ft(ctx, KF_SUMS);
    sums_kernel<<<(unsigned)ceil_div(warps * 32, 256), 256, 0, ctx->stream>>>(ds->d_rows, ds->ld, km->d_sorted_rows,
                                                                             km->d_off, km->C, d_sums_io, d_counts_io);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_finalize_step_dev(vers_kmeans* km, const float* d_sums, const uint64_t* d_counts,
                                                 uint32_t* changed) {
    if (!km || !d_sums || !d_counts) return fail(VERS_ERR_ARG, "kmeans_finalize_step_dev: null argument");
    vers_dataset* ds = km->ds;
    vers_ctx* ctx = ds->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    VERS_CUDA(cudaMemsetAsync(km->d_flag, 0, 4, ctx->stream));
    finalize_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(d_sums, d_counts, km->d_cents, km->d_next, km->C,
                                                                ds->dim, ds->ld, km->d_flag);
    VERS_LAUNCH_CHECK(ctx);
    uint32_t flag = 0;
    VERS_CUDA(cudaMemcpyAsync(&flag, km->d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        float* t = km->d_cents;
        km->d_cents = km->d_next;
        km->d_next = t;
    }
    if (changed) *changed = flag;
    return VERS_OK;
}

extern "C" int32_t vers_kmeans_cost_step(vers_kmeans* km, float* cost_io) {
    if (!km || !cost_io) return fail(VERS_ERR_ARG, "kmeans_cost_step: null argument");
    vers_dataset* ds = km->ds;
    vers_ctx* ctx = ds->ctx;
    if (ds->n == 0) return VERS_OK;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    if (!km->d_rowdist) VERS_CUDA(cudaMalloc(&km->d_rowdist, ds->n * 4 + 4));
    float* d_acc = km->d_rowdist + ds->n;
    VERS_CUDA(cudaMemcpyAsync(d_acc, cost_io, 4, cudaMemcpyHostToDevice, ctx->stream));
    rowdist_kernel<<<(unsigned)ceil_div(ds->n, 128), 128, 0, ctx->stream>>>(ds->d_rows, ds->n, ds->ld, km->d_cents,
                                                                           km->d_assign, km->d_rowdist);
    VERS_LAUNCH_CHECK(ctx);
    seqsum_kernel<<<1, 32, 0, ctx->stream>>>(km->d_rowdist, ds->n, d_acc);
    VERS_LAUNCH_CHECK(ctx);
    VERS_CUDA(cudaMemcpyAsync(cost_io, d_acc, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    return VERS_OK;
}

static int32_t kmeans_update_local(vers_kmeans* km, uint32_t* changed) {
    vers_ctx* ctx = km->ds->ctx;
    VERS_CUDA(cudaMemsetAsync(km->d_sums, 0, (size_t)km->C * km->ds->ld * 4, ctx->stream));
    VERS_CUDA(cudaMemsetAsync(km->d_counts, 0, (size_t)km->C * 8, ctx->stream));
    VERS_TRY(vers_kmeans_sums_step_dev(km, km->d_sums, km->d_counts));
    return vers_kmeans_finalize_step_dev(km, km->d_sums, km->d_counts, changed);
}

extern "C" int32_t vers_kmeans_fit(vers_kmeans* km, uint32_t max_iterations, uint32_t* iterations_run) {
    if (!km) return fail(VERS_ERR_ARG, "kmeans_fit: null");
    uint32_t it = 0;
    for (; it < max_iterations; ++it) {
        VERS_TRY(vers_kmeans_assign_step(km));
        uint32_t changed = 0;
        VERS_TRY(kmeans_update_local(km, &changed));
        if (!changed) {
            ++it;
            break;
        }
    }
    if (iterations_run) *iterations_run = it;
    return vers_kmeans_assign_step(km);  // the final assign, ivfflat.rs:98
}

extern "C" int32_t vers_kmeans_assign(vers_dataset* ds, const float* centroids, uint32_t num_clusters,
                                      uint32_t stride_floats, uint64_t* assignments) {
    vers_kmeans* km = nullptr;
    VERS_TRY(vers_kmeans_create(ds, num_clusters, &km));
    int32_t rc = vers_kmeans_set_centroids(km, centroids, stride_floats);
    if (rc == VERS_OK) rc = vers_kmeans_assign_step(km);
    if (rc == VERS_OK) rc = vers_kmeans_get_assignments(km, assignments);
    vers_kmeans_free(km);
    return rc;
}

extern "C" int32_t vers_kmeans_update(vers_dataset* ds, const uint64_t* assignments, uint32_t num_clusters,
                                      float* centroids, uint64_t* counts) {
    if (!ds || (!assignments && ds->n) || !centroids) return fail(VERS_ERR_ARG, "kmeans_update: null argument");
    vers_kmeans* km = nullptr;
    VERS_TRY(vers_kmeans_create(ds, num_clusters, &km));
    vers_ctx* ctx = ds->ctx;
    int32_t rc = VERS_OK;
    uint64_t* d_wide = nullptr;
    if (ds->n) {
        cudaError_t e = cudaMalloc(&d_wide, ds->n * 8);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_wide, assignments, ds->n * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(km->d_flag, 0, 4, ctx->stream);
        if (e == cudaSuccess) {
            narrow_assign_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_wide, ds->n, num_clusters, km->d_assign,
                                                                             km->d_flag);
            ctx->launches += 1;
            uint32_t bad = 0;
            e = cudaMemcpyAsync(&bad, km->d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e == cudaSuccess && bad)
                rc = fail(VERS_ERR_PANIC, "kmeans_update: assignment >= num_clusters (index out of bounds, ivfflat.rs:53)");
        }
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "kmeans_update: %s", cudaGetErrorString(e));
        cudaFree(d_wide);
    }
    km->csr_valid = false;
    uint32_t changed = 0;
    if (rc == VERS_OK) rc = kmeans_update_local(km, &changed);
    if (rc == VERS_OK) {
        // after finalize the freshly computed centroids are in d_cents when they differ from the zeros we started
        // with, else d_next holds the same values (all zeros); read from whichever is current
        const float* src = changed ? km->d_cents : km->d_next;
        cudaError_t e = cudaMemcpy2DAsync(centroids, (size_t)ds->dim * 4, src, (size_t)ds->ld * 4, (size_t)ds->dim * 4,
                                          num_clusters, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && counts)
            e = cudaMemcpyAsync(counts, km->d_counts, (size_t)num_clusters * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "kmeans_update: %s", cudaGetErrorString(e));
    }
    vers_kmeans_free(km);
    return rc;
}
