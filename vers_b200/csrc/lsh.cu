// lsh.cu — "LSH" random-hyperplane forest (indexes/lsh.rs).
//
// hash: Hyperplane::point_is_above (lsh.rs:27-29) for every row x every plane,
//   bit = (dot(coef_p, row) + const_p) >= 0.0, dot summed left to right in fp32 with no FMA (base.rs:91-93).
//   Same exact-order tile engine as the distance paths (OP_DOT, 2 fp32 instructions per pair-dimension); the sign
//   epilogue adds the constant with one rounded add, exactly the reference's `dot + constant`.
#include "scan.cuh"

namespace vers {

struct HashParams {
    RowSrc A;  // rows
    RowSrc B;  // planes
    uint32_t ld;
    const float* consts;
    uint8_t* bits;  // [nA][P]
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, 2) lsh_hash_kernel(HashParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int MA = Cfg::MA, MB = Cfg::MB, NTA = Cfg::NTA, NTB = Cfg::NTB;
    const int tid = threadIdx.x, ta = tid % NTA, tb = tid / NTA;
    const uint64_t a0 = (uint64_t)blockIdx.x * Cfg::TA;
    const uint64_t P = p.B.n;
    for (uint64_t b0 = 0; b0 < P; b0 += Cfg::TB) {
        float acc[MA][MB];
        tile_compute<Cfg, OP_DOT>(acc, p.A, a0, p.B, b0, p.ld, smem);
#pragma unroll
        for (int j = 0; j < MB; ++j) {
            uint64_t c = b0 + (uint64_t)(tb + j * NTB);
            if (c < P) {
                float k = p.consts[c];
#pragma unroll
                for (int i = 0; i < MA; ++i) {
                    uint64_t r = a0 + (uint64_t)(ta + i * NTA);
                    if (r < p.A.n) p.bits[r * P + c] = (__fadd_rn(acc[i][j], k) >= 0.0f) ? 1 : 0;
                }
            }
        }
    }
}

template <class Cfg>
static int32_t launch_hash(vers_ctx* ctx, const HashParams& p) {
    auto kern = lsh_hash_kernel<Cfg>;
    size_t smem = (size_t)Cfg::TILE_FLOATS * 4;
    VERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)ceil_div(p.A.n, Cfg::TA), Cfg::NT, smem, ctx->stream>>>(p);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

int32_t upload_queries(vers_ctx* ctx, const float* q, uint32_t nq, uint32_t stride, uint32_t dim, uint32_t ld,
                       float** d_q);

}  // namespace vers

using namespace vers;

extern "C" int32_t vers_lsh_hash_dev(vers_dataset* ds, const float* d_planes, uint32_t num_planes,
                                     const float* d_consts, uint8_t* d_bits) {
    if (!ds || !d_planes || !d_consts || !d_bits) return fail(VERS_ERR_ARG, "lsh_hash_dev: null argument");
    if (ds->n == 0 || num_planes == 0) return VERS_OK;
    vers_ctx* ctx = ds->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    HashParams p;
    p.A = RowSrc{ds->d_rows, nullptr, ds->ld, ds->n};
    p.B = RowSrc{d_planes, nullptr, ds->ld, num_planes};
    p.ld = ds->ld;
    p.consts = d_consts;
    p.bits = d_bits;
    FamilyTimer ft(ctx, KF_LSH_HASH);
    if (num_planes <= 16) return launch_hash<NarrowCfg>(ctx, p);
    return launch_hash<WideCfg>(ctx, p);
}

extern "C" int32_t vers_lsh_hash(vers_dataset* ds, const float* planes, uint32_t num_planes,
                                 uint32_t plane_stride_floats, const float* consts, uint8_t* bits) {
    if (!ds || (!planes && num_planes) || (!consts && num_planes) || (!bits && num_planes && ds->n))
        return fail(VERS_ERR_ARG, "lsh_hash: null argument");
    if (plane_stride_floats < ds->dim) return fail(VERS_ERR_ARG, "lsh_hash: plane stride < dim");
    if (ds->n == 0 || num_planes == 0) return VERS_OK;
    vers_ctx* ctx = ds->ctx;
    VERS_CUDA(cudaSetDevice(ctx->device));
    float* d_planes = nullptr;
    float* d_consts = nullptr;
    uint8_t* d_bits = nullptr;
    size_t nb = (size_t)ds->n * num_planes;
    int32_t rc = upload_queries(ctx, planes, num_planes, plane_stride_floats, ds->dim, ds->ld, &d_planes);
    if (rc == VERS_OK && cudaMalloc(&d_consts, (size_t)num_planes * 4) != cudaSuccess)
        rc = fail(VERS_ERR_NOMEM, "lsh_hash: cudaMalloc");
    if (rc == VERS_OK && cudaMalloc(&d_bits, nb) != cudaSuccess) rc = fail(VERS_ERR_NOMEM, "lsh_hash: cudaMalloc");
    if (rc == VERS_OK) {
        cudaError_t e = cudaMemcpyAsync(d_consts, consts, (size_t)num_planes * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "lsh_hash: %s", cudaGetErrorString(e));
    }
    if (rc == VERS_OK) rc = vers_lsh_hash_dev(ds, d_planes, num_planes, d_consts, d_bits);
    if (rc == VERS_OK) {
        cudaError_t e = cudaMemcpyAsync(bits, d_bits, nb, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "lsh_hash: %s", cudaGetErrorString(e));
    }
    cudaFree(d_planes);
    cudaFree(d_consts);
    cudaFree(d_bits);
    return rc;
}

// ---- forest entry points (build / search / add / flatten): lsh_forest.cu
