// scan.cuh — range scan + partial top-k + merge, declared for flat.cu / ivf.cu / lsh.cu.
#pragma once
#include "engine.cuh"

namespace vers {

// partial results: for query q, part t (t < nparts): k entries at ((q * nparts) + t) * k
struct FlatScanParams {
    RowSrc A;  // database rows
    RowSrc B;  // queries
    uint32_t ld;
    uint32_t k, kpad;
    uint64_t rows_per_chunk;
    uint32_t nparts;  // gridDim.x * NSPLIT
    float* part_d;
    uint32_t* part_p;
    const uint32_t* nB_dev;  // optional: the live query count is read on the device (blocks past it exit at once)
};

// Generic merge: one warp per query folds the entries [begin, end) of (d, p) into the final top-k by (d, id).
//   id = map ? map[p] : id_base + p;  entries with p == 0xffffffff are empty.
struct MergeParams {
    const float* part_d;
    const uint32_t* part_p;
    const uint64_t* seg;    // optional entry offsets: query q owns [seg[q*seg_stride], seg[(q+1)*seg_stride]) * seg_scale
    uint64_t seg_scale;     // multiplier applied to seg[] values
    uint32_t seg_stride;    // stride between consecutive queries' boundaries in seg[]
    uint64_t per_query;     // entries per query when seg == null
    const uint64_t* map;    // optional position -> id
    uint64_t id_base;
    uint32_t nq, k;
    uint64_t* out_ids;      // [nq][k]
    float* out_d;           // [nq][k]
    uint32_t* out_cnt;      // [nq] optional
    const uint32_t* qmask;  // optional [nq]: only queries with a non-zero mask are merged/written
    const uint32_t* nq_dev = nullptr;  // optional: live query count read on the device
    const uint32_t* skip_if_zero = nullptr;  // optional: no-op launch when this device word is 0
};

// scan rows [0, A.n) of A against nq queries (B), exact order; writes the global top-k per query.
// metric: VERS_METRIC_*; family: KernelFamily for timing.  Uses ctx scratch (caller holds ctx->mu).
int32_t scan_topk_dev(vers_ctx* ctx, const RowSrc& A, const RowSrc& B, uint32_t nq, uint32_t ld, uint32_t k,
                      uint32_t metric, const uint64_t* id_map, uint64_t id_base, uint64_t* d_ids, float* d_d,
                      uint32_t* d_cnt, int family);

struct ScanPlan {
    bool narrow;
    uint64_t nchunks, rows_per_chunk;
    uint32_t nparts;
    size_t entries;
    size_t bytes;  // scratch needed by scan_topk_run
};
ScanPlan scan_topk_plan(const vers_ctx* ctx, uint64_t nA, uint32_t nq, uint32_t k);
// same as scan_topk_dev but carves its partial buffers from `scratch` (>= pl.bytes, 256-byte aligned)
int32_t scan_topk_run(vers_ctx* ctx, const ScanPlan& pl, void* scratch, const RowSrc& A, const RowSrc& B, uint32_t nq,
                      uint32_t ld, uint32_t k, uint32_t metric, const uint64_t* id_map, uint64_t id_base,
                      uint64_t* d_ids, float* d_d, uint32_t* d_cnt, int family, const uint32_t* nq_dev = nullptr);

int32_t launch_merge(vers_ctx* ctx, const MergeParams& mp);

// ivf.cu: exhaustive L2 search of a batch through the tensor-core candidate path (top-M keys -> exact-order rerank ->
// certificate -> exact redo of uncertified queries).  Returns VERS_ERR_UNSUPPORTED (and *used_tc = false) when the
// batch is not eligible; the caller then runs the exact-order engine.  stats: >= 8 device counters.  norm must be
// padded with +inf up to a multiple of 64 rows.  flat_path 2 forces the list-scan style kernel (rows on the M side, the
// table re-streamed per 32 queries); otherwise eligible batches take tc_flat_kernel (queries resident in TMEM), which
// streams a tile-major tf32 image of the table: *row_tiles_io caches it (allocated on first use, owned by the caller).
int32_t flat_search_tc_plan_and_run(vers_ctx* ctx, const float* rows, uint64_t n, uint32_t ld, const float* norm,
                                    const uint32_t* nmax_bits, uint64_t id_base, unsigned long long* stats,
                                    const float* d_queries, uint32_t nq, uint32_t k, uint64_t* d_ids, float* d_d,
                                    uint32_t* d_cnt, bool* used_tc, int flat_path = 0, float** row_tiles_io = nullptr);
int32_t launch_rownorm(vers_ctx* ctx, const float* rows, uint32_t ld, uint64_t n, float* norm, uint32_t* nmax_bits);

// small-batch streaming scan (flat.cu): every exact l2sq distance of nq <= 8 queries to a row table, dense [NQ][n]
bool flat_stream_fits(uint32_t ld, uint32_t nq);
// smallest query batch that takes the tensor-core candidate path against one table (probe / exhaustive search)
inline uint32_t tc_min_batch() {
    static const uint32_t v = [] {
        const char* e = getenv("VERS_TC_MIN_NQ");
        const int x = e ? atoi(e) : 9;
        return (uint32_t)(x < 1 ? 1 : x);
    }();
    return v;
}
int32_t flat_stream_dense(vers_ctx* ctx, const float* rows, uint64_t n, uint32_t ld, const float* d_queries, uint32_t nq,
                          float* qpad, float* dense_out, int family);

// exclusive scan of n uint32 -> uint64 out[n+1] (single block; n up to a few million is fine)
int32_t launch_exclusive_scan(vers_ctx* ctx, const uint32_t* d_in, uint64_t n, uint64_t* d_out,
                              const uint32_t* skip_if_zero = nullptr);

}  // namespace vers
