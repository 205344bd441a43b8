// api.cu — context, errors, scratch arena and datasets (Vec<Vector<N>> resident in HBM).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace vers {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int32_t scratch_reserve(vers_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return VERS_OK;
    // the old arena may still be referenced by kernels in flight on our stream
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes + bytes / 4 + (1u << 20);
    VERS_CUDA(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return VERS_OK;
}

int32_t io_reserve(vers_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->io_bytes) return VERS_OK;
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->io) cudaFree(ctx->io);
    ctx->io = nullptr;
    ctx->io_bytes = 0;
    size_t want = bytes + bytes / 4 + (1u << 16);
    VERS_CUDA(cudaMalloc(&ctx->io, want));
    ctx->io_bytes = want;
    return VERS_OK;
}

int32_t upload_queries(vers_ctx* ctx, const float* q, uint32_t nq, uint32_t stride, uint32_t dim, uint32_t ld,
                       float** d_q) {
    *d_q = nullptr;
    if (nq == 0) return VERS_OK;
    VERS_CUDA(cudaMalloc(d_q, (size_t)nq * ld * sizeof(float)));
    if (ld != dim) VERS_CUDA(cudaMemsetAsync(*d_q, 0, (size_t)nq * ld * sizeof(float), ctx->stream));
    VERS_CUDA(cudaMemcpy2DAsync(*d_q, (size_t)ld * 4, q, (size_t)stride * 4, (size_t)dim * 4, nq,
                                cudaMemcpyHostToDevice, ctx->stream));
    return VERS_OK;
}

// cuTensorMapEncodeTiled is a driver-API entry point; resolve it through the runtime so the library only links cudart
static int32_t make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, uint32_t elem_bytes, const void* base,
                            uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(VERS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_elems * elem_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, dtype, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VERS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return VERS_OK;
}
int32_t make_tmap_2d_f32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_floats,
                         uint32_t box_rows, uint32_t box_cols) {
    return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, rows, cols, row_stride_floats, box_rows, box_cols);
}
int32_t make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_halfs,
                         uint32_t box_rows, uint32_t box_cols) {
    return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, rows, cols, row_stride_halfs, box_rows, box_cols);
}

// ---------------------------------------------------------------- kernels
__global__ void synth_kernel(float* rows, uint64_t n, uint32_t dim, uint32_t ld, uint64_t seed, uint64_t center_seed,
                             uint32_t kind, uint32_t n_centers, uint64_t row0) {
    uint64_t total = n * (uint64_t)ld;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / ld;
        uint32_t c = (uint32_t)(i - r * ld);
        rows[i] = c < dim ? vers_synth_elem(seed, center_seed, kind, n_centers, row0 + r, c, dim) : 0.0f;
    }
}

// Vector::normalize (indexes/base.rs:95-105): m = sqrt(Σ x_i*x_i) summed left to right; unchanged if m < 1e-6;
// else x_i / m (IEEE division).  One thread owns one row's sequential sum; rows are staged through shared memory
// in 32-column chunks so global reads stay coalesced.
constexpr int NORM_ROWS = 128;
__global__ void __launch_bounds__(NORM_ROWS) normalize_kernel(float* rows, uint64_t n, uint32_t dim, uint32_t ld) {
    __shared__ float tile[NORM_ROWS][33];
    __shared__ float mag[NORM_ROWS];
    const uint64_t r0 = (uint64_t)blockIdx.x * NORM_ROWS;
    const int t = threadIdx.x;
    float s = 0.0f;
    for (uint32_t k0 = 0; k0 < dim; k0 += 32) {
        for (int f = t; f < NORM_ROWS * 32; f += NORM_ROWS) {
            int r = f >> 5, c = f & 31;
            float v = 0.0f;
            if (r0 + r < n && k0 + c < dim) v = rows[(r0 + r) * (uint64_t)ld + k0 + c];
            tile[r][c] = v;
        }
        __syncthreads();
        uint32_t kmax = min(32u, dim - k0);
        for (uint32_t c = 0; c < kmax; ++c) {
            float x = tile[t][c];
            s = __fadd_rn(s, __fmul_rn(x, x));
        }
        __syncthreads();
    }
    mag[t] = __fsqrt_rn(s);
    __syncthreads();
    for (uint64_t f = t; f < (uint64_t)NORM_ROWS * dim; f += NORM_ROWS) {
        uint32_t r = (uint32_t)(f / dim), c = (uint32_t)(f % dim);
        if (r0 + r < n) {
            float m = mag[r];
            if (!(m < 1e-6f)) {
                float* p = rows + (r0 + r) * (uint64_t)ld + c;
                *p = __fdiv_rn(*p, m);
            }
        }
    }
}

}  // namespace vers

using namespace vers;

extern "C" const char* vers_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t vers_abi_version(void) { return 1; }

extern "C" int32_t vers_ctx_create(int32_t device, vers_ctx** out) {
    if (!out) return fail(VERS_ERR_ARG, "ctx_create: out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(VERS_ERR_CUDA, "no CUDA device available (%s); libvers_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) return fail(VERS_ERR_ARG, "ctx_create: device %d out of range [0,%d)", device, count);
    VERS_CUDA(cudaSetDevice(device));
    vers_ctx* ctx = new vers_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    VERS_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    VERS_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
    *out = ctx;
    return VERS_OK;
}

extern "C" int32_t vers_ctx_destroy(vers_ctx* ctx) {
    if (!ctx) return VERS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scan_flags) cudaFree(ctx->scan_flags);
    if (ctx->io) cudaFree(ctx->io);
    for (int i = 0; i < KF_COUNT; ++i) {
        for (cudaEvent_t e : ctx->ev0[i]) cudaEventDestroy(e);
        for (cudaEvent_t e : ctx->ev1[i]) cudaEventDestroy(e);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
    delete ctx;
    return VERS_OK;
}

extern "C" int32_t vers_ctx_set_stream(vers_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(VERS_ERR_ARG, "ctx_set_stream: null ctx");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    // switching INTO a capturing stream (CUDA graph capture of a search step): a synchronize would invalidate the
    // capture; the caller has drained the old stream before starting the capture
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing((cudaStream_t)cuda_stream, &cap) != cudaSuccess) {
        cudaGetLastError();
        cap = cudaStreamCaptureStatusNone;
    }
    if (cap == cudaStreamCaptureStatusNone) cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && cap == cudaStreamCaptureStatusNone) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return VERS_OK;
}

extern "C" int32_t vers_ctx_sync(vers_ctx* ctx) {
    if (!ctx) return fail(VERS_ERR_ARG, "ctx_sync: null ctx");
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    return VERS_OK;
}

extern "C" int32_t vers_ctx_launch_count(vers_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return fail(VERS_ERR_ARG, "ctx_launch_count: null");
    *out = ctx->launches;
    return VERS_OK;
}

extern "C" int32_t vers_ctx_enable_timing(vers_ctx* ctx, int32_t on) {
    if (!ctx) return fail(VERS_ERR_ARG, "ctx_enable_timing: null");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    if (on && ctx->ev0[0].empty()) {
        for (int i = 0; i < KF_COUNT; ++i) {
            ctx->ev0[i].resize(vers_ctx::EV_RING);
            ctx->ev1[i].resize(vers_ctx::EV_RING);
            for (int j = 0; j < vers_ctx::EV_RING; ++j) {
                VERS_CUDA(cudaEventCreate(&ctx->ev0[i][j]));
                VERS_CUDA(cudaEventCreate(&ctx->ev1[i][j]));
            }
        }
    }
    for (int i = 0; i < KF_COUNT; ++i) ctx->ev_used[i] = 0;  // (re)start the measurement window
    ctx->timing = on != 0;
    return VERS_OK;
}

extern "C" int32_t vers_ctx_kernel_ms(vers_ctx* ctx, int32_t which, float* total_ms, uint64_t* timed_launches) {
    if (!ctx || which < 0 || which >= KF_COUNT) return fail(VERS_ERR_ARG, "ctx_kernel_ms: bad argument");
    VERS_CUDA(cudaSetDevice(ctx->device));
    float total = 0.f;
    for (uint64_t j = 0; j < ctx->ev_used[which]; ++j) {
        float ms = 0.f;
        VERS_CUDA(cudaEventSynchronize(ctx->ev1[which][j]));
        VERS_CUDA(cudaEventElapsedTime(&ms, ctx->ev0[which][j], ctx->ev1[which][j]));
        total += ms;
    }
    if (total_ms) *total_ms = total;
    if (timed_launches) *timed_launches = ctx->ev_used[which];
    return VERS_OK;
}

// ---------------------------------------------------------------- datasets
static int32_t dataset_alloc(vers_ctx* ctx, uint64_t n, uint32_t dim, uint64_t id_base, vers_dataset** out) {
    if (!ctx || !out) return fail(VERS_ERR_ARG, "dataset: null argument");
    if (dim == 0) return fail(VERS_ERR_ARG, "dataset: dim must be > 0");
    *out = nullptr;
    VERS_CUDA(cudaSetDevice(ctx->device));
    vers_dataset* ds = new vers_dataset();
    ds->ctx = ctx;
    ds->n = n;
    ds->dim = dim;
    ds->ld = round_up(dim, 4);
    ds->id_base = id_base;
    size_t bytes = (size_t)(n ? n : 1) * ds->ld * sizeof(float);
    cudaError_t e = cudaMalloc(&ds->d_rows, bytes);
    if (e != cudaSuccess) {
        delete ds;
        return fail(VERS_ERR_NOMEM, "dataset: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    *out = ds;
    return VERS_OK;
}

extern "C" int32_t vers_dataset_upload(vers_ctx* ctx, const float* rows, uint64_t n, uint32_t dim,
                                       uint32_t stride_floats, uint64_t id_base, vers_dataset** out) {
    if (!rows && n) return fail(VERS_ERR_ARG, "dataset_upload: rows is null");
    if (stride_floats < dim) return fail(VERS_ERR_ARG, "dataset_upload: stride %u < dim %u", stride_floats, dim);
    VERS_TRY(dataset_alloc(ctx, n, dim, id_base, out));
    vers_dataset* ds = *out;
    if (n) {
        if (ds->ld != dim) VERS_CUDA(cudaMemsetAsync(ds->d_rows, 0, (size_t)n * ds->ld * 4, ctx->stream));
        VERS_CUDA(cudaMemcpy2DAsync(ds->d_rows, (size_t)ds->ld * 4, rows, (size_t)stride_floats * 4, (size_t)dim * 4,
                                    n, cudaMemcpyHostToDevice, ctx->stream));
        VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return VERS_OK;
}

extern "C" int32_t vers_dataset_synth(vers_ctx* ctx, uint64_t seed, uint64_t center_seed, uint32_t kind,
                                      uint32_t n_centers, uint64_t row0, uint64_t n, uint32_t dim, int32_t normalize,
                                      vers_dataset** out) {
    if (kind > VERS_SYNTH_CLUSTERED) return fail(VERS_ERR_ARG, "dataset_synth: unknown kind %u", kind);
    if (kind == VERS_SYNTH_CLUSTERED && n_centers == 0) return fail(VERS_ERR_ARG, "dataset_synth: n_centers == 0");
    VERS_TRY(dataset_alloc(ctx, n, dim, row0, out));
    vers_dataset* ds = *out;
    if (n) {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        synth_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ds->d_rows, n, dim, ds->ld, seed, center_seed, kind,
                                                                 n_centers, row0);
        VERS_LAUNCH_CHECK(ctx);
    }
    if (normalize) VERS_TRY(vers_dataset_normalize(ds));
    return VERS_OK;
}

extern "C" int32_t vers_dataset_normalize(vers_dataset* ds) {
    if (!ds) return fail(VERS_ERR_ARG, "dataset_normalize: null");
    if (ds->n == 0) return VERS_OK;
    vers_ctx* ctx = ds->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    normalize_kernel<<<(unsigned)ceil_div(ds->n, NORM_ROWS), NORM_ROWS, 0, ctx->stream>>>(ds->d_rows, ds->n, ds->dim,
                                                                                         ds->ld);
    VERS_LAUNCH_CHECK(ctx);
    ds->epoch += 1;    // k-means states over this dataset recompute their ||row||^2 on the next assign
    if (ds->d_norm) {  // cached ||row||^2 of the tensor-core exhaustive search are stale now
        VERS_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ds->d_norm);
        cudaFree(ds->d_nmax);
        cudaFree(ds->d_stats);
        cudaFree(ds->d_tiles);
        ds->d_tiles = nullptr;
        ds->d_norm = nullptr;
        ds->d_nmax = nullptr;
        ds->d_stats = nullptr;
    }
    return VERS_OK;
}

extern "C" int32_t vers_dataset_info(const vers_dataset* ds, uint64_t* n, uint32_t* dim, uint32_t* ld,
                                     uint64_t* id_base) {
    if (!ds) return fail(VERS_ERR_ARG, "dataset_info: null");
    if (n) *n = ds->n;
    if (dim) *dim = ds->dim;
    if (ld) *ld = ds->ld;
    if (id_base) *id_base = ds->id_base;
    return VERS_OK;
}

extern "C" int32_t vers_dataset_download(const vers_dataset* ds, uint64_t row0, uint64_t n, float* out,
                                         uint32_t stride_floats) {
    if (!ds || (!out && n)) return fail(VERS_ERR_ARG, "dataset_download: null");
    if (row0 + n > ds->n) return fail(VERS_ERR_ARG, "dataset_download: rows [%llu,%llu) out of range",
                                      (unsigned long long)row0, (unsigned long long)(row0 + n));
    if (stride_floats < ds->dim) return fail(VERS_ERR_ARG, "dataset_download: stride < dim");
    if (n == 0) return VERS_OK;
    VERS_CUDA(cudaSetDevice(ds->ctx->device));
    VERS_CUDA(cudaMemcpy2DAsync(out, (size_t)stride_floats * 4, ds->d_rows + row0 * ds->ld, (size_t)ds->ld * 4,
                                (size_t)ds->dim * 4, n, cudaMemcpyDeviceToHost, ds->ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ds->ctx->stream));
    return VERS_OK;
}

extern "C" int32_t vers_dataset_device_ptr(const vers_dataset* ds, void** ptr) {
    if (!ds || !ptr) return fail(VERS_ERR_ARG, "dataset_device_ptr: null");
    *ptr = ds->d_rows;
    return VERS_OK;
}

extern "C" int32_t vers_dataset_wrap_device(vers_ctx* ctx, const float* d_rows, uint64_t n, uint32_t dim,
                                            uint64_t id_base, vers_dataset** out) {
    if (!ctx || !out || (!d_rows && n)) return fail(VERS_ERR_ARG, "dataset_wrap_device: null argument");
    if (dim == 0) return fail(VERS_ERR_ARG, "dataset_wrap_device: dim == 0");
    vers_dataset* ds = new vers_dataset();
    ds->ctx = ctx;
    ds->d_rows = const_cast<float*>(d_rows);
    ds->n = n;
    ds->dim = dim;
    ds->ld = round_up(dim, 4);
    ds->id_base = id_base;
    ds->owned = false;  // the caller keeps the memory alive for the life of the handle
    *out = ds;
    return VERS_OK;
}

extern "C" int32_t vers_dataset_free(vers_dataset* ds) {
    if (!ds) return VERS_OK;
    cudaSetDevice(ds->ctx->device);
    cudaStreamSynchronize(ds->ctx->stream);
    if (ds->owned && ds->d_rows) cudaFree(ds->d_rows);
    cudaFree(ds->d_norm);
    cudaFree(ds->d_nmax);
    cudaFree(ds->d_stats);
    cudaFree(ds->d_tiles);
    delete ds;
    return VERS_OK;
}
