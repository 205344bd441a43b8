// common.cuh — shared host/device plumbing of libvers_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vers_device.h"
#include "../../include/vers_synth.h"

namespace vers {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int32_t fail(int32_t code, const char* fmt, ...);

#define VERS_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::vers::fail(_e == cudaErrorMemoryAllocation ? VERS_ERR_NOMEM : VERS_ERR_CUDA,    \
                                "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define VERS_TRY(expr)               \
    do {                             \
        int32_t _rc = (expr);        \
        if (_rc != VERS_OK) return _rc; \
    } while (0)

// kernel families for vers_ctx_last_kernel_ms
enum KernelFamily { KF_LIST_SCAN = 0, KF_FLAT_SCAN = 1, KF_ASSIGN = 2, KF_SUMS = 3, KF_LSH_HASH = 4, KF_PROBE = 5, KF_CAND_SCAN = 6, KF_RERANK = 7, KF_COUNT = 8 };

}  // namespace vers

// ---------------------------------------------------------------- handles
struct vers_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    bool timing = false;
    // per kernel family: ring of CUDA event pairs recorded on `stream` around each launch while timing is on
    static constexpr int EV_RING = 512;
    std::vector<cudaEvent_t> ev0[vers::KF_COUNT];
    std::vector<cudaEvent_t> ev1[vers::KF_COUNT];
    uint64_t ev_used[vers::KF_COUNT] = {};
    uint64_t fam_launches[vers::KF_COUNT] = {};
    // grow-only device staging for the host-pointer entry points (queries in, results out)
    void* io = nullptr;
    size_t io_bytes = 0;
    // scratch arena (grow-only), owned by ctx, used by search calls; guarded by mu
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    unsigned long long* scan_flags = nullptr;  // [256] inter-block carries of the chained exclusive scan
    cudaStream_t cap_stream = nullptr;  // private stream the call graphs are captured on (the caller's stream may be
                                        // the legacy default stream, which cannot be captured)
    std::recursive_mutex mu;  // recursive: a host-buffer entry point holds it across the device-pointer calls it is made of
};

struct vers_dataset {
    vers_ctx* ctx = nullptr;
    float* d_rows = nullptr;  // [n][ld]
    uint64_t n = 0;
    uint32_t dim = 0;
    uint32_t ld = 0;  // round_up(dim, 4), pad columns are zero
    uint64_t id_base = 0;
    bool owned = true;
    uint64_t epoch = 1;  // bumped whenever the rows are rewritten in place (normalize): cached ||row||^2 become stale
    // tensor-core exhaustive search: ||row||^2 (any order), their max, 8 counters; built on first use, dropped when the
    // rows change (normalize)
    float* d_norm = nullptr;
    uint32_t* d_nmax = nullptr;
    unsigned long long* d_stats = nullptr;
    float* d_tiles = nullptr;  // tile-major tf32 image of the rows (tc_flat_kernel's B operand), built on first use
    int flat_mode = 0;  // 0: tensor-core candidate paths for eligible batches, 1: exact-order engine only, 2: like 0
                        // without the query-block kernel (the list-scan style kernel for every eligible batch)
};

namespace vers {

inline uint32_t round_up(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }
inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// host: 2D fp32 TMA tensor map {inner = cols, outer = rows}, box {box_cols, box_rows}, 128-byte swizzle, zero fill
int32_t make_tmap_2d_f32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_floats,
                         uint32_t box_rows, uint32_t box_cols);
// the same over fp16 elements (box_cols = 64 halfs = one 128-byte swizzle row)
int32_t make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_halfs,
                         uint32_t box_rows, uint32_t box_cols);

// grow-only scratch; caller holds ctx->mu
int32_t scratch_reserve(vers_ctx* ctx, size_t bytes);
// grow-only device staging for host-pointer calls; caller holds ctx->mu
int32_t io_reserve(vers_ctx* ctx, size_t bytes);

// A host-buffer entry point that is called again and again with the same shapes (a serving loop) replays its device
// work from a CUDA graph instead of re-launching ~25 small kernels: first call with a key runs eagerly (and sizes every
// arena), the second captures, later ones replay.  The key holds everything the captured launches depend on (index
// state, shapes, arena addresses, stream); any change starts over.  Capture problems of any kind fall back to eager
// launches for good.  Caller holds ctx->mu.
struct GraphCache {
    uint64_t key[12] = {};
    int state = 0;  // 0 empty, 1 key seen once, 2 graph ready, -1 disabled
    cudaGraphExec_t exec = nullptr;
    void reset() {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
        state = 0;
    }
};

template <class F>
int32_t graph_cached_run(vers_ctx* ctx, GraphCache& gc, const uint64_t (&key)[12], F&& fn) {
    static const bool off = getenv("VERS_NO_CALL_GRAPH") != nullptr;
    static const bool dbg = getenv("VERS_DEBUG_CALL_GRAPH") != nullptr;
    if (off || ctx->timing || gc.state < 0) return fn();
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return fn();  // the caller is capturing this call itself
    }
    if (gc.state == 0 || memcmp(gc.key, key, sizeof(gc.key)) != 0) {
        if (dbg && gc.state != 0) {
            fprintf(stderr, "[vers] call graph: key changed:");
            for (int i = 0; i < 12; ++i)
                if (gc.key[i] != key[i]) fprintf(stderr, " [%d] %llx -> %llx", i, (unsigned long long)gc.key[i], (unsigned long long)key[i]);
            fprintf(stderr, "\n");
        }
        gc.reset();
        memcpy(gc.key, key, sizeof(gc.key));
        gc.state = 1;
        return fn();
    }
    if (gc.state == 1) {
        // captured on a private stream (every launch of fn goes to ctx->stream, which is swapped for the duration), replayed
        // on the caller's
        if (!ctx->cap_stream && cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            gc.state = -1;
            return fn();
        }
        if (cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            if (dbg) fprintf(stderr, "[vers] call graph: begin capture failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            cudaGetLastError();
            gc.state = -1;
            return fn();
        }
        const uint64_t launches0 = ctx->launches;
        cudaStream_t real = ctx->stream;
        ctx->stream = ctx->cap_stream;
        const int32_t rc = fn();
        ctx->stream = real;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(ctx->cap_stream, &g);
        if (rc == VERS_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&gc.exec, g, 0);
        if (g) cudaGraphDestroy(g);
        if (rc != VERS_OK || e != cudaSuccess || !gc.exec) {  // nothing ran: do it the plain way, now and from now on
            if (dbg)
                fprintf(stderr, "[vers] call graph: capture failed (rc %d: %s; cuda: %s), staying eager\n", rc,
                        vers_last_error(), cudaGetErrorString(e));
            cudaGetLastError();
            gc.reset();
            gc.state = -1;
            ctx->launches = launches0;
            return fn();
        }
        gc.state = 2;
        if (dbg) fprintf(stderr, "[vers] call graph: captured %llu launches\n", (unsigned long long)(ctx->launches - launches0));
    }
    VERS_CUDA(cudaGraphLaunch(gc.exec, ctx->stream));
    return VERS_OK;
}

struct ScratchCarver {
    char* base;
    size_t off = 0;
    explicit ScratchCarver(void* b) : base((char*)b) {}
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T* p = (T*)(base + off);
        off += count * sizeof(T);
        return p;
    }
    template <typename T>
    void plan(size_t count) {
        off = (off + 255) & ~size_t(255);
        off += count * sizeof(T);
    }
};

// records family timing events around a launch when enabled
struct FamilyTimer {
    vers_ctx* ctx;
    int fam;
    bool on;
    FamilyTimer(vers_ctx* c, int f) : ctx(c), fam(f) {  // f < 0: untimed (a launch inside an enclosing timer)
        on = fam >= 0 && ctx->timing && ctx->ev_used[fam] < (uint64_t)vers_ctx::EV_RING;
        if (on) cudaEventRecord(ctx->ev0[fam][ctx->ev_used[fam]], ctx->stream);
    }
    ~FamilyTimer() {
        if (on) {
            cudaEventRecord(ctx->ev1[fam][ctx->ev_used[fam]], ctx->stream);
            ctx->ev_used[fam] += 1;
        }
        if (fam >= 0) ctx->fam_launches[fam] += 1;
    }
};

#define VERS_LAUNCH_CHECK(ctx)                                                                          \
    do {                                                                                                \
        (ctx)->launches += 1;                                                                           \
        cudaError_t _e = cudaGetLastError();                                                            \
        if (_e != cudaSuccess)                                                                          \
            return ::vers::fail(VERS_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                                    \
    } while (0)

// ---------------------------------------------------------------- device helpers
#if defined(__CUDACC__)

constexpr unsigned FULL_MASK = 0xffffffffu;

// total order used everywhere a stable sort by distance is restated: (distance, position/id)
__device__ __forceinline__ bool pair_less(float d0, uint32_t p0, float d1, uint32_t p1) {
    return (d0 < d1) || (d0 == d1 && p0 < p1);
}
__device__ __forceinline__ bool pair_less64(float d0, uint64_t p0, float d1, uint64_t p1) {
    return (d0 < d1) || (d0 == d1 && p0 < p1);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;  // src-size 0 => 16 zero bytes written
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

#endif  // __CUDACC__

}  // namespace vers
