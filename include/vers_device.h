/*
 * vers_device.h — C ABI of the B200-native device layer behind vers' `Index<N>` trait.
 *
 * The reference (ashrielbrian/vers) has no process or device boundary; this header IS the boundary the
 * north-star introduces: Rust host (rust/vers-cuda-sys) -> extern "C" -> CUDA (sm_100a).
 * Every entry point names the reference code it replaces (paths relative to /root/reference/vers/src).
 *
 * Conventions
 *   - plain pointers and sizes only; `usize` == uint64_t; rows arrive as `const float*` with a stride in floats
 *     (size_of::<Vector<N>>()/4: 320 for N=300, 768, 128 — indexes/base.rs:15 `#[repr(align(256))]`).
 *   - every function returns int32_t: 0 = VERS_OK, negative = error class; the message is in vers_last_error()
 *     (thread-local).  VERS_ERR_PANIC marks inputs on which the reference itself panics (unwrap / index OOB);
 *     the Rust shim turns any non-zero status into panic!, matching the reference's error behaviour.
 *   - results are structure-of-arrays (ids[], dists[], counts[]) because Vec<(usize, f32)> has no stable layout.
 *     Unused result slots are filled with id = UINT64_MAX, dist = +inf.
 *   - returned distances are SQUARED L2 (indexes/base.rs:119-126) unless metric = VERS_METRIC_COSINE
 *     (1 - dot, indexes/base.rs:155).
 *   - `_dev` variants take DEVICE pointers and enqueue on the context's stream without synchronising: they are
 *     what bench.py times with inputs resident in HBM, and what the multi-GPU driver calls between collectives.
 *     The un-suffixed variants take HOST pointers, copy, run, copy back and synchronise (the `e2e` path).
 *   - there is no CPU fallback: without a CUDA device every call fails with VERS_ERR_CUDA.
 */
#ifndef VERS_DEVICE_H
#define VERS_DEVICE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VERS_OK 0
#define VERS_ERR_ARG (-1)
#define VERS_ERR_CUDA (-2)
#define VERS_ERR_NOMEM (-3)
#define VERS_ERR_PANIC (-4)       /* the reference would panic on these inputs */
#define VERS_ERR_UNSUPPORTED (-5) /* valid in the reference, outside this build's limits (e.g. top_k > 128) */

#define VERS_METRIC_L2SQ 0u
#define VERS_METRIC_COSINE 1u

#define VERS_MAX_TOPK 128u

typedef struct vers_ctx vers_ctx;         /* one per GPU: device id, stream, scratch arena                     */
typedef struct vers_dataset vers_dataset; /* Vec<Vector<N>> resident in HBM, row-major, ld = round_up(dim, 4)  */
typedef struct vers_kmeans vers_kmeans;   /* k-means state (centroids, assignments) on one GPU's row shard     */
typedef struct vers_ivf vers_ivf;         /* IVFFlatIndex<N> (indexes/ivfflat.rs:9-15), list-major in HBM      */
typedef struct vers_lsh vers_lsh;         /* ANNIndex<N>     (indexes/lsh.rs:47-55), forest + rows in HBM      */

const char* vers_last_error(void);
int32_t vers_abi_version(void);

/* ---- context -------------------------------------------------------------------------------------------- */
int32_t vers_ctx_create(int32_t device, vers_ctx** out);
int32_t vers_ctx_destroy(vers_ctx* ctx);
/* borrow a cudaStream_t (e.g. torch's current stream) so that the caller's CUDA events bracket our kernels */
int32_t vers_ctx_set_stream(vers_ctx* ctx, void* cuda_stream);
int32_t vers_ctx_sync(vers_ctx* ctx);
/* number of kernels this library has launched on ctx since creation (bench.py's gpu_launches) */
int32_t vers_ctx_launch_count(vers_ctx* ctx, uint64_t* out);
/* device time of each kernel family, measured with CUDA event pairs recorded on ctx's stream around every launch
 * while timing is on (up to 512 launches per family per window) — bench.py's live roofline numbers.
 * which: 0 = list scan (exact order), 1 = flat scan, 2 = k-means assign, 3 = k-means sums, 4 = lsh hash, 5 = probe,
 * 6 = list scan candidate pass, 7 = exact rerank + certificate.
 * enable_timing(on) starts a new window; kernel_ms sums the window (synchronises on the recorded events). */
int32_t vers_ctx_enable_timing(vers_ctx* ctx, int32_t on);
int32_t vers_ctx_kernel_ms(vers_ctx* ctx, int32_t which, float* total_ms, uint64_t* timed_launches);

/* ---- datasets: Vec<Vector<N>> (indexes/base.rs:15-17) ----------------------------------------------------- */
/* id_base: global id of local row 0 (row shards of a multi-GPU index are contiguous blocks of the global rows) */
int32_t vers_dataset_upload(vers_ctx* ctx, const float* rows, uint64_t n, uint32_t dim, uint32_t stride_floats,
                            uint64_t id_base, vers_dataset** out);
/* counter-based synthetic rows generated on the device, see vers_synth.h; rows [row0, row0+n), id_base = row0 */
int32_t vers_dataset_synth(vers_ctx* ctx, uint64_t seed, uint64_t center_seed, uint32_t kind, uint32_t n_centers,
                           uint64_t row0, uint64_t n, uint32_t dim, int32_t normalize, vers_dataset** out);
int32_t vers_dataset_info(const vers_dataset* ds, uint64_t* n, uint32_t* dim, uint32_t* ld, uint64_t* id_base);
/* Vector::normalize on every row in place (indexes/base.rs:95-105; the loader applies it, utils.rs:7-66) */
int32_t vers_dataset_normalize(vers_dataset* ds);
int32_t vers_dataset_download(const vers_dataset* ds, uint64_t row0, uint64_t n, float* out, uint32_t stride_floats);
int32_t vers_dataset_device_ptr(const vers_dataset* ds, void** ptr);
/* A dataset handle over rows that already live in device memory ([n][round_up(dim,4)] fp32, pad columns zero).  The
 * caller keeps the memory alive for the life of the handle; nothing is copied (multi-GPU drivers hand the rows they
 * received from their peers straight to the index build). */
int32_t vers_dataset_wrap_device(vers_ctx* ctx, const float* d_rows, uint64_t n, uint32_t dim, uint64_t id_base,
                                 vers_dataset** out);
int32_t vers_dataset_free(vers_dataset* ds);

/* ---- exhaustive search: utils::search_exhaustive (utils.rs:68-82) ----------------------------------------- */
/* top_k nearest rows per query by (distance, id); ids are id_base + row. counts[q] = min(top_k, n). */
int32_t vers_flat_search(vers_dataset* ds, const float* queries, uint32_t nq, uint32_t q_stride_floats,
                         uint32_t top_k, uint32_t metric, uint64_t* ids, float* dists, uint32_t* counts);
/* d_queries: device, nq rows of ld floats (ld = dataset ld, zero padded); outputs device [nq][top_k] */
int32_t vers_flat_search_dev(vers_dataset* ds, const float* d_queries, uint32_t nq, uint32_t top_k, uint32_t metric,
                             uint64_t* d_ids, float* d_dists, uint32_t* d_counts);
/* Up to 8 queries: one exact-order streaming pass (HBM-bound for 1-2 queries).  Batches of >= 9 queries with
 * top_k <= 64 and the L2 metric take the two-stage path of the inverted-list scan over the whole dataset: tensor-core
 * candidate keys (TMA + tcgen05 kind::tf32; split hi/lo with one pass of the table per 32 queries below 96 queries, one
 * pass per 128 queries from a tile-major image above) -> exact-order rerank of the 32/64/128 best -> rounding-error
 * certificate -> exact-order redo of uncertified queries.  ids and distance bits are the reference's either way.  mode 1 forces the exact-order engine (tests, timing).  Stats of the last tensor-core
 * search: out[4] = uncertified queries, out[5] = rows re-ranked, out[6] = bit pattern of the largest observed
 * |candidate value - exact value|; all zero when the exact-order engine ran. */
int32_t vers_flat_set_mode(vers_dataset* ds, int32_t mode);
int32_t vers_flat_last_search_stats(const vers_dataset* ds, uint64_t out[8]);

/* ---- HNSW distance offload: Vector::cosine_similarity_simd (indexes/base.rs:158-223, the only
 *      distance hnsw.rs evaluates: hnsw.rs:146, :258, :273) and squared_euclidean_simd (base.rs:225-294) for a batch of
 *      (query, row) pairs.  out[i] = distance(queries[pair_query[i]], row with global id pair_row[i]) in the reference's
 *      SIMD association: 64-wide chunks summed by the ordered reduce_sum, then 4-wide chunks, then the scalar tail,
 *      chunk sums added in order; metric VERS_METRIC_COSINE returns 1 - dot like the reference.  The graph traversal
 *      stays on the host; it hands the neighbour ids of a step (or of many queries' steps) to this call.
 *      pair_query == NULL: every pair uses query 0.  A pair naming a row outside the dataset or a query >= nq returns
 *      VERS_ERR_PANIC (id_to_vec.get(..).unwrap(), hnsw.rs:133, :270); the _dev variant writes NaN for such pairs and
 *      counts them in *d_bad_count (which the caller zeroes). */
int32_t vers_pair_distances_simd(vers_dataset* ds, const float* queries, uint32_t nq, uint32_t q_stride_floats,
                                 const uint32_t* pair_query, const uint64_t* pair_row, uint64_t n_pairs, uint32_t metric,
                                 float* out);
int32_t vers_pair_distances_simd_dev(vers_dataset* ds, const float* d_queries, uint32_t nq, uint32_t q_stride_floats,
                                     const uint32_t* d_pair_query, const uint64_t* d_pair_row, uint64_t n_pairs,
                                     uint32_t metric, float* d_out, uint32_t* d_bad_count);

/* ---- k-means: IVFFlatIndex::{assign_to_clusters, update_centroids, build_kmeans, calculate_kmeans_cost}
 *      (indexes/ivfflat.rs:29-46, :47-71, :73-100, :138-149) ------------------------------------------------ */
int32_t vers_kmeans_create(vers_dataset* ds, uint32_t num_clusters, vers_kmeans** out);
int32_t vers_kmeans_free(vers_kmeans* km);
/* initialize_centroids (ivfflat.rs:18-27) with the random draws injected: centroid j = copy of LOCAL row
 * init_rows[j] (single GPU), or set explicitly from host memory (multi-GPU: rank owning the row broadcasts it) */
int32_t vers_kmeans_init_from_rows(vers_kmeans* km, const uint64_t* init_rows);
int32_t vers_kmeans_set_centroids(vers_kmeans* km, const float* centroids, uint32_t stride_floats);
int32_t vers_kmeans_get_centroids(vers_kmeans* km, float* centroids, uint32_t stride_floats);
int32_t vers_kmeans_get_assignments(vers_kmeans* km, uint64_t* assignments);
int32_t vers_kmeans_centroids_device_ptr(vers_kmeans* km, void** ptr, uint32_t* ld);
/* device pointer of the current assignments: uint32 [n], cluster of every local row (ivfflat.rs:29-46) */
int32_t vers_kmeans_assign_device_ptr(vers_kmeans* km, void** ptr);
/* assign_to_clusters over the local rows with the current centroids (first minimum wins) */
int32_t vers_kmeans_assign_step(vers_kmeans* km);
/* 0 (default): tensor-core candidate argmin (TMA + tcgen05 kind::tf32, M=128 x N=256 tiles) + a rounding-error
 * certificate on the gap between the two smallest values, uncertified rows re-assigned in exact order;
 * 1: exact order everywhere.  Both give the reference's assignments bit for bit.  For rows of <= 128 floats mode 0
 * runs the single-MMA kernel (one MMA per K step, rows resident in tensor memory, the four best candidates re-ranked
 * in exact order inside the kernel, certificate against the fifth key) with fp16 operands (kind::f16: the same 11-bit
 * significand as tf32 at twice the MMA rate; both operands scaled by one power of two so nothing overflows); mode 3 runs
 * the same kernel with kind::tf32; mode 2 forces the split-precision kernel. */
int32_t vers_kmeans_set_mode(vers_kmeans* km, int32_t mode);
/* rows of the most recent assign step whose candidate argmin was not certified (redone in exact order) */
int32_t vers_kmeans_last_assign_stats(vers_kmeans* km, uint64_t* uncertified_rows);
/* the Σ of update_centroids over the LOCAL rows in row order, continuing from the running sums in
 * d_sums_io [C][ld] f32 / d_counts_io [C] u64 (device; pass zeros on the first shard).  Chaining shards in row
 * order reproduces the reference's global left-to-right association exactly. */
int32_t vers_kmeans_sums_step_dev(vers_kmeans* km, float* d_sums_io, uint64_t* d_counts_io);
/* new = sums / count (zero vector when count == 0, ivfflat.rs:63-67); *changed = bit patterns differ from the
 * current centroids (ivfflat.rs:84-93); the new centroids are adopted only when they differ. */
int32_t vers_kmeans_finalize_step_dev(vers_kmeans* km, const float* d_sums, const uint64_t* d_counts,
                                      uint32_t* changed);
/* calculate_kmeans_cost over the local rows: acc = *cost_io; acc += l2sq(row, centroid[assign]) in row order */
int32_t vers_kmeans_cost_step(vers_kmeans* km, float* cost_io);
/* build_kmeans on one GPU: <= max_iterations of (assign, update, bitwise compare) + the final assign */
int32_t vers_kmeans_fit(vers_kmeans* km, uint32_t max_iterations, uint32_t* iterations_run);
/* host-pointer conveniences used by the parity tests (one step each, reference argument meaning) */
int32_t vers_kmeans_assign(vers_dataset* ds, const float* centroids, uint32_t num_clusters, uint32_t stride_floats,
                           uint64_t* assignments);
int32_t vers_kmeans_update(vers_dataset* ds, const uint64_t* assignments, uint32_t num_clusters,
                           float* centroids /* [C][dim] */, uint64_t* counts);

/* ---- IVFFlatIndex (indexes/ivfflat.rs) --------------------------------------------------------------------- */
/* build_index (ivfflat.rs:102-136): best of num_attempts k-means runs by strict `<` on cost; init_rows is
 * [num_attempts][num_clusters] local row numbers. */
int32_t vers_ivf_build_index(vers_dataset* ds, uint32_t num_clusters, uint32_t num_attempts,
                             uint32_t max_iterations, const uint64_t* init_rows, vers_ivf** out);
/* take centroids + assignments from a fitted k-means state (multi-GPU build) */
int32_t vers_ivf_from_kmeans(vers_kmeans* km, vers_ivf** out);
/* rebuild the device mirror from deserialised parts (after Index::load_index, base.rs:45-58);
 * assignments == NULL recomputes them on the device */
int32_t vers_ivf_from_parts(vers_dataset* ds, const float* centroids, uint32_t num_clusters, uint32_t stride_floats,
                            const uint64_t* assignments, vers_ivf** out);
/* The struct of ivfflat.rs:9-15 from parts that already live on the device: centroids [num_clusters][ld], the
 * cluster of every row of `ds` (uint32) and, optionally, the GLOBAL id of every row (ascending; default
 * id_base + row).  Multi-GPU drivers that shard by inverted list build each GPU's lists this way after exchanging
 * rows; inside a list rows stay in ascending id order like `ids[c]` (ivfflat.rs:123-127). */
int32_t vers_ivf_from_parts_dev(vers_dataset* ds, const float* d_centroids, uint32_t num_clusters,
                                const uint32_t* d_assignments, const uint64_t* d_row_ids, vers_ivf** out);
int32_t vers_ivf_free(vers_ivf* ivf);
int32_t vers_ivf_info(const vers_ivf* ivf, uint64_t* n, uint32_t* dim, uint32_t* num_clusters, float* best_cost,
                      uint32_t* best_attempt);
int32_t vers_ivf_get_centroids(const vers_ivf* ivf, float* centroids, uint32_t stride_floats);
int32_t vers_ivf_get_assignments(const vers_ivf* ivf, uint64_t* assignments);
int32_t vers_ivf_get_list_sizes(const vers_ivf* ivf, uint64_t* sizes);
/* ids[list] (ivfflat.rs:14) and, optionally, the rows of that list in the same order (rows may be NULL) */
int32_t vers_ivf_get_list(const vers_ivf* ivf, uint32_t list, uint64_t* ids, float* rows, uint32_t stride_floats);
/* work done by the most recent search on this index (device counters, synchronises):
 * out[0] = rows in the DISTINCT lists touched (the algorithmic HBM stream), out[1] = Σ over (query, list) pairs of
 * the list length (un-deduplicated rows x queries), out[2] = work items, out[3] = distinct lists touched,
 * out[4] = queries whose certificate failed and were redone by the exact-order scan, out[5] = candidates
 * re-ranked in exact order, out[6] = bit pattern (low 32 bits) of the largest observed |candidate value - exact
 * value| among the re-ranked rows (validates the certificate's error allowance), out[7] = the tensor-core centroid
 * probe: low 32 bits = queries whose probe certificate failed (redone by the exact-order engine), high 32 bits =
 * candidate centroids re-ranked in exact order (0 when the exact-order probe ran) */
int32_t vers_ivf_last_search_stats(const vers_ivf* ivf, uint64_t out[8]);
/* nprobe >= 1 searches: 0 (default) = tensor-core candidate pass (TMA + tcgen05 kind::tf32 on the fp32 rows) +
 * exact-order rerank of the candidates + a rounding-error certificate, uncertified queries redone in exact order;
 * (the default splits every operand into tf32 hi + lo parts: 3 MMAs per K step, fp32-grade candidate values);
 * 1 = exact order everywhere; 2 = like 0 with an fp32 FMA (SIMT) candidate pass; 3 = like 0 with plain TF32;
 * 4 = candidates from an fp16 copy of the inverted lists (tcgen05 kind::f16, half the HBM bytes per scanned row; the
 * copy is built by the first search that wants it: +dim*2 bytes per row of device memory, kept current by vers_ivf_add),
 * same exact fp32 rerank, certificate extended by the copy's measured rounding error (Cauchy-Schwarz).
 * Every mode returns the reference's ids and distance bits; the knob selects speed and memory, never results. */
int32_t vers_ivf_set_mode(vers_ivf* ivf, int32_t mode);
/* Index::search_approximate (ivfflat.rs:153-198) for a batch.  nprobe == 0: the reference's semantics (nearest
 * list, spill to the next list while fewer than top_k found, output = concatenated per-list prefixes).
 * nprobe >= 1 (extension, BASELINE config 4): global top_k by (distance, id) over the nprobe nearest lists.
 * A call shape (nq, top_k, nprobe) that repeats on an unchanged index replays its device work from a cached CUDA graph
 * (first call eager, second captured, later ones replayed; environment VERS_NO_CALL_GRAPH=1 disables it). */
int32_t vers_ivf_search(vers_ivf* ivf, const float* queries, uint32_t nq, uint32_t q_stride_floats, uint32_t top_k,
                        uint32_t nprobe, uint64_t* ids, float* dists, uint32_t* counts);
int32_t vers_ivf_search_dev(vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t top_k, uint32_t nprobe,
                            uint64_t* d_ids, float* d_dists, uint32_t* d_counts);
/* The two halves of a nprobe >= 1 search, for drivers that split the centroid probe of a batch over several GPUs
 * (the centroids are replicated, so any GPU can probe any query): probe_dev writes the nprobe nearest lists of each
 * query by (distance, centroid index), exact order (ivfflat.rs:155-161); search_probed_dev scans exactly those lists. */
int32_t vers_ivf_probe_dev(vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t nprobe, uint64_t* d_probe_ids);
int32_t vers_ivf_search_probed_dev(vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t top_k, uint32_t nprobe,
                                   const uint64_t* d_probe_ids, uint64_t* d_ids, float* d_dists, uint32_t* d_counts);
/* Index::add (ivfflat.rs:200-213): nearest centroid by first minimum; vec_id is IGNORED like the reference
 * (the id is assignments.len()); *assigned_id / *cluster report what was stored. */
int32_t vers_ivf_add(vers_ivf* ivf, const float* embedding, uint64_t vec_id, uint64_t* assigned_id,
                     uint32_t* cluster);
/* Index::add for a batch, in order (exactly n sequential adds: the centroids do not move on add): embedding i gets id
 * assignments.len() + i; one exact-order assign launch, at most one re-layout of the lists, one scatter.  A NaN
 * distance returns VERS_ERR_PANIC with nothing modified.  assigned_ids / clusters may be NULL. */
int32_t vers_ivf_add_batch(vers_ivf* ivf, const float* embeddings, uint64_t n, uint32_t stride_floats,
                           uint64_t* assigned_ids, uint32_t* clusters);

/* merge `parts` per-shard result lists per query into the global top_k by (distance, id): the step after the
 * NCCL all-gather of per-GPU top-k.  Part p's [nq][top_k] block starts at d_ids_all + p*part_stride_ids (u64
 * elements) and d_dists_all + p*part_stride_dists (floats); a stride of 0 means the dense nq*top_k. */
int32_t vers_topk_merge_dev(vers_ctx* ctx, const uint64_t* d_ids_all, const float* d_dists_all, uint32_t parts,
                            uint64_t part_stride_ids, uint64_t part_stride_dists, uint32_t nq, uint32_t top_k,
                            uint64_t* d_ids, float* d_dists, uint32_t* d_counts);

/* The same exchange as ONE kernel over NVLink peer memory instead of all-gather + merge (multi-GPU hosts, one
 * process per GPU): every rank creates an exchange buffer, the 64-byte CUDA IPC handles are swapped by the host
 * (any transport), vers_peer_connect maps the peers' buffers.  vers_peer_gather_merge_dev then stores this rank's
 * local top-k into every peer's buffer, raises a flag, waits for all ranks' flags and merges by (distance, id).
 * Every rank must call it once per batch, in the same order; slot_bytes >= nq * top_k * 12. */
typedef struct vers_peer vers_peer;
int32_t vers_peer_create(vers_ctx* ctx, uint32_t world, uint32_t rank, uint64_t slot_bytes, vers_peer** out,
                         uint8_t ipc_handle_out[64]);
int32_t vers_peer_connect(vers_peer* peer, const uint8_t* all_handles /* [world][64], rank order */);
int32_t vers_peer_gather_merge_dev(vers_peer* peer, const uint64_t* d_local_ids, const float* d_local_dists,
                                   uint32_t nq, uint32_t top_k, uint64_t* d_ids, float* d_dists, uint32_t* d_counts);
int32_t vers_peer_free(vers_peer* peer);

/* ---- multi-GPU: one process per GPU of an NVLink/NVSwitch box ---------------------------------
 * The reference parallelises with rayon inside one process (ivfflat.rs:31, lsh.rs:146,268); this layer shards the
 * same work over GPUs.  vers_comm = this process's membership in a group of `world` ranks.  Bootstrap is NCCL's:
 * rank 0 calls vers_comm_unique_id, the HOST ships those 128 bytes to the other processes (any transport it has),
 * every rank calls vers_comm_create (collective).  libnccl.so.2 is resolved at that point (dlopen), single-GPU users
 * never need it.  world == 1 is allowed (unique_id may be NULL) and makes every vers_sharded_* call the plain
 * single-GPU path.  All collective calls below must be made by every rank, in the same order.
 * vers_comm_create also establishes every connection the build and search will use (NCCL connects lazily on first
 * use, which costs seconds at 8 ranks) and maps the peers' exchange buffers (CUDA IPC). */
#define VERS_REDUCE_CHAINED 0   /* ordered: rank r continues rank r-1's running sums — bit-identical to the reference */
#define VERS_REDUCE_ALLREDUCE 1 /* ncclAllReduce of per-shard sums: fastest, association differs from the reference's */
typedef struct vers_comm vers_comm;
int32_t vers_comm_unique_id(uint8_t id_out[128]);
int32_t vers_comm_create(vers_ctx* ctx, uint32_t world, uint32_t rank, const uint8_t unique_id[128], vers_comm** out);
int32_t vers_comm_destroy(vers_comm* comm);
int32_t vers_comm_info(const vers_comm* comm, uint32_t* world, uint32_t* rank, double* last_exchange_seconds);
/* rendezvous of all ranks + stream synchronise; max over ranks of a host value (timing: "max over ranks") */
int32_t vers_comm_barrier(vers_comm* comm);
/* diagnostic: GPU globaltimer stamps (ns) of the most recent fused top-k exchange of a sharded search step on this rank:
 * [0] kernel entry, [1] own slice stored on every peer and flag raised, [2] every rank's flag seen, [3] merged */
int32_t vers_debug_peer_times(vers_comm* comm, uint64_t out_ns[4]);
int32_t vers_comm_max_f64(vers_comm* comm, double* value_io);
/* IVFFlatIndex::build_kmeans (ivfflat.rs:73-100) over contiguous row shards (km is over this rank's vers_dataset,
 * whose id_base is its first GLOBAL row).  init_rows_global: [num_clusters] GLOBAL row numbers (initialize_centroids,
 * ivfflat.rs:18-27, draws injected); the owner of a row contributes it.  assign is local; update_centroids
 * (ivfflat.rs:47-71) is reduced per `reduce`; the bitwise convergence test sees identical sums on every rank. */
int32_t vers_sharded_kmeans_fit(vers_comm* comm, vers_kmeans* km, const uint64_t* init_rows_global,
                                uint32_t max_iterations, int32_t reduce, uint32_t* iterations_run);
/* calculate_kmeans_cost (ivfflat.rs:138-149) folded in GLOBAL row order (rank r continues rank r-1's value) */
int32_t vers_sharded_kmeans_cost(vers_comm* comm, vers_kmeans* km, float* cost);
/* The index over the fitted k-means state, sharded BY INVERTED LIST: every row (with its global id and cluster) goes
 * to the rank that owns its list (largest list first onto the least-loaded rank) in one all-to-all over NVLink; a
 * rank's lists keep ascending-id order like ids[c] (ivfflat.rs:123-127); lists it does not own are empty there,
 * the centroid table is whole.  The seconds the all-to-all took are reported by vers_comm_info. */
int32_t vers_sharded_ivf_build(vers_comm* comm, vers_kmeans* km, vers_ivf** out);
/* the list -> owner rank table that build uses, from the GLOBAL list sizes (host arithmetic, needs no GPU) */
int32_t vers_sharded_list_owners(const uint64_t* list_sizes, uint32_t num_clusters, uint32_t world,
                                 uint32_t* owner_out);
/* Index::search_approximate (ivfflat.rs:153-198, nprobe >= 1 extension) for a batch over the sharded index: every
 * rank passes the SAME queries and receives the SAME global result.  Per batch: each rank probes 1/world of the
 * queries, the probe lists are all-gathered, each rank scans the lists it owns, the per-rank top-k are exchanged and
 * merged by (distance, id).  Both exchanges are stores into the peers' IPC-mapped buffers plus flags (no NCCL call,
 * no host synchronisation inside a step: the _dev variant is capturable in a CUDA graph); the second one is fused
 * with the merge in one kernel. */
int32_t vers_sharded_ivf_search(vers_comm* comm, vers_ivf* ivf, const float* queries, uint32_t nq,
                                uint32_t q_stride_floats, uint32_t top_k, uint32_t nprobe, uint64_t* ids, float* dists,
                                uint32_t* counts);
int32_t vers_sharded_ivf_search_dev(vers_comm* comm, vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t top_k,
                                    uint32_t nprobe, uint64_t* d_ids, float* d_dists, uint32_t* d_counts);
/* ANNIndex::search_approximate (lsh.rs:264-282) for a batch with the forest REPLICATED on every rank (tree construction
 * is a data-dependent recursion and does not shard; the reference parallelises the search over trees and queries,
 * lsh.rs:146,268): every rank passes the SAME queries, searches its 1/world slice of them on its replica, and the
 * slices are all-gathered (ncclAllGather of ids + distances + counts), so every rank returns the full result.
 * Hashing a row shard (vers_lsh_hash) needs no collective at all: rows are independent. */
int32_t vers_sharded_lsh_search(vers_comm* comm, vers_lsh* lsh, const float* queries, uint32_t nq,
                                uint32_t q_stride_floats, uint32_t top_k, uint64_t* ids, float* dists, uint32_t* counts);

/* ---- "LSH" random-hyperplane forest (indexes/lsh.rs) -------------------------------------------------------- */
/* Hyperplane::point_is_above (lsh.rs:27-29) for every row x every plane: bits[r*P + p] = dot(plane_p, row_r) +
 * consts[p] >= 0.0 */
int32_t vers_lsh_hash(vers_dataset* ds, const float* planes, uint32_t num_planes, uint32_t plane_stride_floats,
                      const float* consts, uint8_t* bits);
int32_t vers_lsh_hash_dev(vers_dataset* ds, const float* d_planes /* [P][ld] */, uint32_t num_planes,
                          const float* d_consts, uint8_t* d_bits);
/* ANNIndex::build_index (lsh.rs:132-161): dedup by bit pattern, num_trees trees, leaves of < max_size rows.
 * The two sample rows per node (choose_multiple(thread_rng, 2), lsh.rs:63-65) come from vers_lsh_pick_pair(seed…) */
int32_t vers_lsh_build_index(vers_ctx* ctx, const float* rows, uint64_t n, uint32_t dim, uint32_t stride_floats,
                             const uint64_t* vector_ids, uint32_t num_trees, uint32_t max_size, uint64_t seed,
                             vers_lsh** out);
int32_t vers_lsh_free(vers_lsh* lsh);
int32_t vers_lsh_info(const vers_lsh* lsh, uint64_t* num_values, uint32_t* num_trees, uint64_t* num_nodes);
/* preorder dump of one tree (node, above-subtree, below-subtree) for structural parity checks; pass NULLs to size */
int32_t vers_lsh_flatten(const vers_lsh* lsh, uint32_t tree, uint8_t* kind, uint32_t* leaf_len, float* planes,
                         float* consts, uint32_t* items, uint32_t* n_nodes, uint32_t* n_inner, uint64_t* n_items);
/* Index::search_approximate (lsh.rs:264-282) for a batch; ties by deduplicated row index */
int32_t vers_lsh_search(vers_lsh* lsh, const float* queries, uint32_t nq, uint32_t q_stride_floats, uint32_t top_k,
                        uint64_t* ids, float* dists, uint32_t* counts);
/* Index::add (lsh.rs:255-263).  vec_id is stored in the leaves as a ROW INDEX like the reference does (lsh.rs:247): an id
 * that is not one returns VERS_ERR_PANIC before anything is modified. */
int32_t vers_lsh_add(vers_lsh* lsh, const float* embedding, uint64_t vec_id);
/* `values` and `ids` of the struct (lsh.rs:47-55): the deduplicated rows in stored order, rows added later included —
 * what Index::save_index serialises.  Either pointer may be NULL; sizes from vers_lsh_info. */
int32_t vers_lsh_get_values(const vers_lsh* lsh, float* values, uint32_t stride_floats, uint64_t* ids);
/* The device forest from deserialised parts (Index::load_index, base.rs:45-58; struct layout lsh.rs:13-55).  values:
 * the stored (already deduplicated) rows; the trees arrive concatenated, each in the preorder of vers_lsh_flatten
 * (node, ABOVE subtree, BELOW subtree): tree_nodes[t] nodes for tree t, kind / leaf_len per node, planes [dim] and
 * consts per inner node, items per leaf.  seed only feeds the sample pairs of later leaf splits (Index::add). */
int32_t vers_lsh_from_parts(vers_ctx* ctx, const float* values, uint64_t n, uint32_t dim, uint32_t stride_floats,
                            const uint64_t* ids, uint32_t num_trees, uint32_t max_size, uint64_t seed,
                            const uint32_t* tree_nodes, const uint8_t* kind, const uint32_t* leaf_len,
                            const float* planes, const float* consts, const uint32_t* items, vers_lsh** out);

#ifdef __cplusplus
}
#endif
#endif /* VERS_DEVICE_H */
