// kmeans.cuh — device k-means state shared by kmeans.cu and ivf.cu
#pragma once
#include "engine.cuh"

struct vers_kmeans {
    vers_dataset* ds = nullptr;
    uint32_t C = 0;
    float* d_cents = nullptr;        // [C][ld] current centroids (pad columns zero)
    float* d_next = nullptr;         // [C][ld] candidate centroids
    uint32_t* d_assign = nullptr;    // [n] cluster of each local row
    uint32_t* d_sorted_rows = nullptr;  // [n] rows grouped by cluster, ascending row inside a cluster
    uint32_t* d_sorted_keys = nullptr;  // [n]
    uint32_t* d_iota = nullptr;         // [n] 0..n-1
    uint32_t* d_hist = nullptr;         // [C]
    uint64_t* d_off = nullptr;          // [C+1]
    float* d_sums = nullptr;            // [C][ld] used by fit()/update on one GPU
    uint64_t* d_counts = nullptr;       // [C]
    float* d_rowdist = nullptr;         // [n] cost scratch (allocated lazily)
    uint32_t* d_flag = nullptr;         // [1]
    uint32_t* d_bad = nullptr;          // [1] set by the exact-order assign when a row has no comparable distance (NaN)
    void* d_cub = nullptr;
    size_t cub_bytes = 0;
    bool csr_valid = false;  // d_sorted_rows/d_off describe the current d_assign
    // tensor-core candidate pass of assign (kmeans_tc.cuh)
    int mode = 0;                    // 0 = tensor-core candidates + certificate + exact redo (tf32-first kernel when
                                     // ld <= 128), 1 = exact order only, 2 = like 0 with the split-precision kernel
    float* d_row_norm = nullptr;     // [n] ||x||^2, recomputed when the rows change (ds->epoch)
    uint64_t norm_epoch = 0;         // ds->epoch d_row_norm was computed for
    float* d_cent_norm = nullptr;    // [C]
    float* d_cent_hi = nullptr;      // [C][ld] tf32 hi part of the centroids
    float* d_cent_lo = nullptr;      // [C][ld] tf32 lo part
    float* d_cent_tiles = nullptr;   // shared-memory image of the rounded centroid tiles (tc_assign1_kernel's B operand)
    uint32_t* d_ncmax = nullptr;     // [1] bits of max ||c||^2
    uint32_t* d_nxmax = nullptr;     // [2] bits of max ||row||^2 (the fp16 kernel's scale) | elements of the fp16
                                     //     centroid image that did not fit
    float f16_scale = 1.0f;          // power of two: fp16(x * scale) cannot overflow for any row (nor for any mean of rows)
    uint32_t* d_flagged = nullptr;   // [n] rows whose candidate argmin was not certified
    uint32_t* d_nflagged = nullptr;  // [1]
    uint32_t* d_exact = nullptr;     // [n] exact re-assignments of the flagged rows
    uint64_t last_flagged = 0;       // statistic of the most recent assign step
};

namespace vers {
// groups the local rows by cluster in ascending row order (stable): fills d_sorted_rows and d_off
int32_t kmeans_build_csr(vers_kmeans* km);
int32_t kmeans_assign_rows(vers_ctx* ctx, const RowSrc& rows, const float* d_cents, uint32_t C, uint32_t ld,
                           uint32_t* d_assign, int family = KF_ASSIGN, uint32_t* d_bad = nullptr);
}  // namespace vers
