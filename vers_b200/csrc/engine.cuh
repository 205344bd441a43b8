// engine.cuh — the exact-order pairwise tile engine and the warp top-k selection shared by every path.
//
// What it computes: for a tile of TA "rows" (database rows, list rows, data points) and TB "columns" (queries,
// centroids, hyperplanes) the value the reference computes for each pair, BIT FOR BIT:
//   OP_L2SQ : s = 0; for i in 0..dim: t = a[i]-b[i]; s = s + t*t      (indexes/base.rs:119-126)
//   OP_DOT  : s = 0; for i in 0..dim: s = s + a[i]*b[i]              (indexes/base.rs:91-93)
// i.e. rounded sub, rounded mul, rounded add, strictly left to right, no FMA (__fsub_rn/__fmul_rn/__fadd_rn are
// never contracted by nvcc).  The sequential chain is per PAIR; parallelism comes from the TAxTB independent
// pairs (each thread owns an MAxMB register micro-tile = MA*MB independent chains), so the kernel is bound by the
// fp32 pipe at 3 (L2) or 2 (dot) instructions per pair-dimension — see DESIGN.md for the roofline.
//
// Staging: both operands stream through shared memory in 32-float k-chunks, double buffered with cp.async
// (LDGSTS, 16 B per request, zero-fill for out-of-range rows/columns — adding (0-0)^2 or 0*0 leaves the running
// sum unchanged, which is what makes padding legal under the exact-order contract).  Rows are padded to
// LDS = 36 floats in shared memory so that the 8 lanes of each LDS.128 phase hit 8 distinct bank groups.
#pragma once
#include "common.cuh"

namespace vers {

enum { OP_L2SQ = 0, OP_DOT = 1, OP_DOT_FMA = 2 };  // OP_DOT_FMA: candidate pass only, any rounding

// source of tile rows: row r lives at base + (idx ? idx[r] : r) * ld
struct RowSrc {
    const float* base;
    const uint32_t* idx;  // optional gather
    uint32_t ld;          // floats, multiple of 4
    uint64_t n;           // number of rows addressable through this source
};

template <int TA_, int TB_, int MA_, int MB_, int STAGES_ = 2>
struct TileCfg {
    static constexpr int TA = TA_, TB = TB_, MA = MA_, MB = MB_;
    static constexpr int NTA = TA / MA, NTB = TB / MB, NT = NTA * NTB;
    static constexpr int KC = 32, LDS = KC + 4, STAGES = STAGES_;
    static constexpr int TILE_FLOATS = STAGES * (TA + TB) * LDS;
    static constexpr int NWARPS = NT / 32;
    static constexpr int LANES_PER_TB = NTA < 32 ? NTA : 32;    // lanes of a warp that share one tb
    static constexpr int TBS_PER_WARP = 32 / LANES_PER_TB;      // distinct tb values inside a warp
    static constexpr int NSPLIT = NTA / LANES_PER_TB;           // warps that share one tb (row halves)
    static constexpr int NLISTS = NWARPS * TBS_PER_WARP * MB;   // private top-k lists per CTA
    static constexpr int A_LOADS = TA * (KC / 4) / NT;          // cp.async per thread per chunk (A side)
    static constexpr int B_LOADS = (TB * (KC / 4) + NT - 1) / NT;
    static_assert(NT % 32 == 0, "whole warps");
    static_assert((TA * (KC / 4)) % NT == 0, "A tile divides evenly");
};
// many columns (>= 32): flat search with a big batch, centroid probe, k-means assign, hashing many planes
using WideCfg = TileCfg<128, 64, 8, 4>;
// few columns (<= 8): inverted-list scan (about nq*nprobe/nlist queries per list), small-batch GEMV-style scans
using NarrowCfg = TileCfg<256, 8, 4, 2>;
// candidate pass of the inverted-list scan: every thread owns 2 rows x all 8 queries of the group, so each row
// element is read from shared memory exactly once and the 8 query values are warp-wide broadcasts
using StreamCfg = TileCfg<256, 8, 1, 8, 2>;  // 2 CTAs/SM (16 warps): measured best of {512x2, 256x1} x {2,4 stages}

template <int OP>
__device__ __forceinline__ void pair_step(float& acc, float a, float b) {
    if (OP == OP_L2SQ) {
        float t = __fsub_rn(a, b);
        acc = __fadd_rn(acc, __fmul_rn(t, t));
    } else if (OP == OP_DOT) {
        acc = __fadd_rn(acc, __fmul_rn(a, b));
    } else {
        acc = __fmaf_rn(a, b, acc);  // 1 instruction per pair-dimension; NOT the reference's rounding
    }
}

// All threads of the CTA call this.  On return acc[i][j] holds the exact value for
// row (a0 + ta + i*NTA) x column (b0 + tb + j*NTB); out-of-range rows/columns hold garbage-free zeros-based sums
// that the caller must ignore.  Ends with a __syncthreads(): shared memory is free on return.
template <class Cfg, int OP>
__device__ __forceinline__ void tile_compute(float (&acc)[Cfg::MA][Cfg::MB], const RowSrc& A, uint64_t a0,
                                             const RowSrc& B, uint64_t b0, uint32_t ld, float* smem) {
    constexpr int TA = Cfg::TA, TB = Cfg::TB, MA = Cfg::MA, MB = Cfg::MB, NTA = Cfg::NTA, NTB = Cfg::NTB;
    constexpr int NT = Cfg::NT, KC = Cfg::KC, LDS = Cfg::LDS;
    const int tid = threadIdx.x;
    const int ta = tid % NTA, tb = tid / NTA;

    // per-thread source pointers for the cp.async slots this thread owns (fixed for the whole tile)
    const float* aptr[Cfg::A_LOADS];
#pragma unroll
    for (int i = 0; i < Cfg::A_LOADS; ++i) {
        int f = tid + i * NT;
        uint64_t r = a0 + (uint64_t)(f >> 3);
        aptr[i] = nullptr;
        if (r < A.n) {
            uint64_t row = A.idx ? (uint64_t)A.idx[r] : r;
            aptr[i] = A.base + row * (uint64_t)A.ld + (uint32_t)((f & 7) * 4);
        }
    }
    const float* bptr[Cfg::B_LOADS];
#pragma unroll
    for (int i = 0; i < Cfg::B_LOADS; ++i) {
        int f = tid + i * NT;
        bptr[i] = nullptr;
        if (f < TB * (KC / 4)) {
            uint64_t r = b0 + (uint64_t)(f >> 3);
            if (r < B.n) {
                uint64_t row = B.idx ? (uint64_t)B.idx[r] : r;
                bptr[i] = B.base + row * (uint64_t)B.ld + (uint32_t)((f & 7) * 4);
            }
        }
    }

#pragma unroll
    for (int i = 0; i < MA; ++i)
#pragma unroll
        for (int j = 0; j < MB; ++j) acc[i][j] = 0.0f;

    auto load_chunk = [&](int stage, uint32_t k0) {
        float* As = smem + stage * (TA + TB) * LDS;
        float* Bs = As + TA * LDS;
#pragma unroll
        for (int i = 0; i < Cfg::A_LOADS; ++i) {
            int f = tid + i * NT;
            bool valid = aptr[i] != nullptr && (k0 + (uint32_t)((f & 7) * 4)) < ld;
            cp_async16(As + (f >> 3) * LDS + (f & 7) * 4, valid ? (const void*)(aptr[i] + k0) : (const void*)A.base,
                       valid);
        }
#pragma unroll
        for (int i = 0; i < Cfg::B_LOADS; ++i) {
            int f = tid + i * NT;
            if (f < TB * (KC / 4)) {
                bool valid = bptr[i] != nullptr && (k0 + (uint32_t)((f & 7) * 4)) < ld;
                cp_async16(Bs + (f >> 3) * LDS + (f & 7) * 4,
                           valid ? (const void*)(bptr[i] + k0) : (const void*)B.base, valid);
            }
        }
        cp_async_commit();
    };

    // STAGES-deep cp.async ring: chunks c+1 .. c+STAGES-1 are in flight while chunk c is consumed, one
    // __syncthreads per chunk (the refill of a buffer is issued after the barrier that retires its last reader)
    constexpr int S = Cfg::STAGES;
    const uint32_t nchunks = (ld + KC - 1) / KC;
#pragma unroll
    for (int st = 0; st < S - 1; ++st) {
        if ((uint32_t)st < nchunks) load_chunk(st, (uint32_t)st * KC); else cp_async_commit();
    }
    for (uint32_t c = 0; c < nchunks; ++c) {
        cp_async_wait<S - 2>();
        __syncthreads();
        if (c + S - 1 < nchunks) load_chunk((int)((c + S - 1) % S), (c + S - 1) * KC); else cp_async_commit();
        const float* As = smem + (int)(c % S) * (TA + TB) * LDS + ta * LDS;
        const float* Bs = smem + (int)(c % S) * (TA + TB) * LDS + TA * LDS + tb * LDS;
        const int kmax = (int)min((uint32_t)KC, ld - c * KC);  // multiple of 4
        auto kstep = [&](int kk) {
            float4 a[MA], b[MB];
#pragma unroll
            for (int i = 0; i < MA; ++i) a[i] = *reinterpret_cast<const float4*>(As + i * NTA * LDS + kk);
#pragma unroll
            for (int j = 0; j < MB; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + j * NTB * LDS + kk);
#pragma unroll
            for (int i = 0; i < MA; ++i)
#pragma unroll
                for (int j = 0; j < MB; ++j) pair_step<OP>(acc[i][j], a[i].x, b[j].x);
#pragma unroll
            for (int i = 0; i < MA; ++i)
#pragma unroll
                for (int j = 0; j < MB; ++j) pair_step<OP>(acc[i][j], a[i].y, b[j].y);
#pragma unroll
            for (int i = 0; i < MA; ++i)
#pragma unroll
                for (int j = 0; j < MB; ++j) pair_step<OP>(acc[i][j], a[i].z, b[j].z);
#pragma unroll
            for (int i = 0; i < MA; ++i)
#pragma unroll
                for (int j = 0; j < MB; ++j) pair_step<OP>(acc[i][j], a[i].w, b[j].w);
        };
        if (kmax == KC) {  // full chunk: branch-free so the loads of step k+1 can be hoisted over the math of step k
#pragma unroll(Cfg::MA * Cfg::MB <= 16 ? 8 : 2)
            for (int kk = 0; kk < KC; kk += 4) kstep(kk);
        } else {
            for (int kk = 0; kk < kmax; kk += 4) kstep(kk);
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // shared memory is free on return
}

// ---------------------------------------------------------------- warp top-k over (distance, position)
// A list is k entries in shared memory sorted ascending by (d, p); the whole warp inserts one candidate.
template <typename P>
__device__ __forceinline__ bool entry_less(float d0, P p0, float d1, P p1) {
    return (d0 < d1) || (d0 == d1 && p0 < p1);
}

template <typename P>
__device__ __forceinline__ void warp_topk_insert(float* sd, P* sp, int k, float v, P p, int lane) {
    int pos = 0;
    for (int base = 0; base < k; base += 32) {
        int i = base + lane;
        bool lt = false;
        if (i < k) lt = entry_less<P>(sd[i], sp[i], v, p);
        pos += __popc(__ballot_sync(FULL_MASK, lt));
    }
    if (pos >= k) return;  // warp-uniform
    for (int base = ((k - 1) / 32) * 32; base >= 0 && base + 32 > pos; base -= 32) {
        int i = base + lane;
        bool mv = (i >= pos) && (i < k - 1);
        float e = 0.f;
        P ep = 0;
        if (mv) {
            e = sd[i];
            ep = sp[i];
        }
        __syncwarp();
        if (mv) {
            sd[i + 1] = e;
            sp[i + 1] = ep;
        }
        __syncwarp();
    }
    if (lane == 0) {
        sd[pos] = v;
        sp[pos] = p;
    }
    __syncwarp();
}

// Fold the register micro-tile of one computed tile into the per-(warp, column) private top-k lists.
// XF: 0 = value as is, 1 = cosine distance 1 - dot (indexes/base.rs:155),
//     2 = candidate key ||x||^2 - 2 x.q (rownorm[p] holds ||x||^2 of the row at position p; ||q||^2 is a per-query
//         constant and is added back by the certifier).
// position stored = (u32)(row + pos_add): monotone in id order inside one scan range.
template <class Cfg, int XF>
__device__ __forceinline__ void tile_select_topk(const float (&acc)[Cfg::MA][Cfg::MB], uint64_t a0, uint64_t r_end,
                                                 uint64_t b0, uint64_t nB, uint32_t k, uint32_t kpad, float* list_d,
                                                 uint32_t* list_p, uint64_t pos_add,
                                                 const float* __restrict__ rownorm = nullptr) {
    constexpr int MA = Cfg::MA, MB = Cfg::MB, NTA = Cfg::NTA, NTB = Cfg::NTB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ta = tid % NTA, tb = tid / NTA;
    const int tb_local = (Cfg::LANES_PER_TB == 32) ? 0 : lane / Cfg::LANES_PER_TB;
#pragma unroll
    for (int j = 0; j < MB; ++j) {
        const int slot = (warp * Cfg::TBS_PER_WARP + tb_local) * MB + j;
        float* sd = list_d + slot * kpad;
        uint32_t* sp = list_p + slot * kpad;
        const bool colvalid = (b0 + (uint64_t)(tb + j * NTB)) < nB;
#pragma unroll
        for (int i = 0; i < MA; ++i) {
            const uint64_t row = a0 + (uint64_t)(ta + i * NTA);
            float v = acc[i][j];
            if (XF == 1) v = __fsub_rn(1.0f, v);
            const uint32_t p = (uint32_t)(row + pos_add);
            bool live = colvalid && row < r_end;
            if (XF == 2) v = live ? __fmaf_rn(-2.0f, v, __ldg(rownorm + p)) : v;
            while (true) {
                float tv = sd[k - 1];
                uint32_t tp = sp[k - 1];
                bool pass = live && entry_less<uint32_t>(v, p, tv, tp);
                unsigned m = __ballot_sync(FULL_MASK, pass);
                if (!m) break;
                int src = __ffs(m) - 1;
                float bv = __shfl_sync(FULL_MASK, v, src);
                uint32_t bp = __shfl_sync(FULL_MASK, p, src);
                int bslot = __shfl_sync(FULL_MASK, slot, src);
                warp_topk_insert<uint32_t>(list_d + bslot * kpad, list_p + bslot * kpad, (int)k, bv, bp, lane);
                if (lane == src) live = false;
            }
        }
    }
}

template <class Cfg>
__device__ __forceinline__ void lists_init(float* list_d, uint32_t* list_p, uint32_t kpad) {
    for (int i = threadIdx.x; i < Cfg::NLISTS * (int)kpad; i += Cfg::NT) {
        list_d[i] = __int_as_float(0x7f800000);
        list_p[i] = 0xffffffffu;
    }
    __syncthreads();
}

// which column / split a private list slot belongs to (inverse of the mapping in tile_select_topk)
template <class Cfg>
__device__ __forceinline__ void slot_to_col(int slot, int& col, int& split) {
    int j = slot % Cfg::MB;
    int wt = slot / Cfg::MB;  // warp * TBS_PER_WARP + tb_local
    int warp = wt / Cfg::TBS_PER_WARP, tb_local = wt % Cfg::TBS_PER_WARP;
    int tb;
    if (Cfg::LANES_PER_TB == 32) {
        tb = warp / Cfg::NSPLIT;
        split = warp % Cfg::NSPLIT;
    } else {
        tb = warp * Cfg::TBS_PER_WARP + tb_local;
        split = 0;
    }
    col = tb + j * Cfg::NTB;
}

inline size_t scan_smem_bytes(int tile_floats, int nlists, uint32_t kpad) {
    return (size_t)tile_floats * 4 + (size_t)nlists * kpad * 8;
}

}  // namespace vers
