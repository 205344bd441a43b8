//! Raw bindings of include/vers_device.h (one declaration per C entry point).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

macro_rules! opaque { ($($n:ident),*) => { $(#[repr(C)] pub struct $n { _p: [u8; 0] })* } }
opaque!(vers_ctx, vers_dataset, vers_kmeans, vers_ivf, vers_lsh);

pub const VERS_OK: i32 = 0;
pub const VERS_ERR_PANIC: i32 = -4;
pub const VERS_METRIC_L2SQ: u32 = 0;
pub const VERS_METRIC_COSINE: u32 = 1;

extern "C" {
    pub fn vers_last_error() -> *const c_char;
    pub fn vers_abi_version() -> i32;
    pub fn vers_ctx_create(device: i32, out: *mut *mut vers_ctx) -> i32;
    pub fn vers_ctx_destroy(ctx: *mut vers_ctx) -> i32;
    pub fn vers_ctx_set_stream(ctx: *mut vers_ctx, cuda_stream: *mut c_void) -> i32;
    pub fn vers_ctx_sync(ctx: *mut vers_ctx) -> i32;
    pub fn vers_dataset_upload(ctx: *mut vers_ctx, rows: *const f32, n: u64, dim: u32, stride_floats: u32, id_base: u64, out: *mut *mut vers_dataset) -> i32;
    pub fn vers_dataset_normalize(ds: *mut vers_dataset) -> i32;
    pub fn vers_dataset_download(ds: *const vers_dataset, row0: u64, n: u64, out: *mut f32, stride_floats: u32) -> i32;
    pub fn vers_dataset_wrap_device(ctx: *mut vers_ctx, d_rows: *const f32, n: u64, dim: u32, id_base: u64, out: *mut *mut vers_dataset) -> i32;
    pub fn vers_dataset_free(ds: *mut vers_dataset) -> i32;
    pub fn vers_flat_search(ds: *mut vers_dataset, queries: *const f32, nq: u32, q_stride_floats: u32, top_k: u32, metric: u32, ids: *mut u64, dists: *mut f32, counts: *mut u32) -> i32;
    pub fn vers_flat_set_mode(ds: *mut vers_dataset, mode: i32) -> i32;
    pub fn vers_flat_last_search_stats(ds: *const vers_dataset, out: *mut u64) -> i32;
    pub fn vers_kmeans_assign(ds: *mut vers_dataset, centroids: *const f32, num_clusters: u32, stride_floats: u32, assignments: *mut u64) -> i32;
    pub fn vers_kmeans_update(ds: *mut vers_dataset, assignments: *const u64, num_clusters: u32, centroids: *mut f32, counts: *mut u64) -> i32;
    pub fn vers_ivf_build_index(ds: *mut vers_dataset, num_clusters: u32, num_attempts: u32, max_iterations: u32, init_rows: *const u64, out: *mut *mut vers_ivf) -> i32;
    pub fn vers_ivf_from_parts(ds: *mut vers_dataset, centroids: *const f32, num_clusters: u32, stride_floats: u32, assignments: *const u64, out: *mut *mut vers_ivf) -> i32;
    pub fn vers_ivf_from_parts_dev(ds: *mut vers_dataset, d_centroids: *const f32, num_clusters: u32, d_assignments: *const u32, d_row_ids: *const u64, out: *mut *mut vers_ivf) -> i32;
    pub fn vers_ivf_set_mode(ivf: *mut vers_ivf, mode: i32) -> i32;
    pub fn vers_ivf_last_search_stats(ivf: *const vers_ivf, out: *mut u64) -> i32;
    pub fn vers_ivf_free(ivf: *mut vers_ivf) -> i32;
    pub fn vers_ivf_get_centroids(ivf: *const vers_ivf, centroids: *mut f32, stride_floats: u32) -> i32;
    pub fn vers_ivf_get_assignments(ivf: *const vers_ivf, assignments: *mut u64) -> i32;
    pub fn vers_ivf_search(ivf: *mut vers_ivf, queries: *const f32, nq: u32, q_stride_floats: u32, top_k: u32, nprobe: u32, ids: *mut u64, dists: *mut f32, counts: *mut u32) -> i32;
    pub fn vers_ivf_add(ivf: *mut vers_ivf, embedding: *const f32, vec_id: u64, assigned_id: *mut u64, cluster: *mut u32) -> i32;
    pub fn vers_lsh_hash(ds: *mut vers_dataset, planes: *const f32, num_planes: u32, plane_stride_floats: u32, consts: *const f32, bits: *mut u8) -> i32;
    pub fn vers_lsh_build_index(ctx: *mut vers_ctx, rows: *const f32, n: u64, dim: u32, stride_floats: u32, vector_ids: *const u64, num_trees: u32, max_size: u32, seed: u64, out: *mut *mut vers_lsh) -> i32;
    pub fn vers_lsh_search(lsh: *mut vers_lsh, queries: *const f32, nq: u32, q_stride_floats: u32, top_k: u32, ids: *mut u64, dists: *mut f32, counts: *mut u32) -> i32;
    pub fn vers_lsh_add(lsh: *mut vers_lsh, embedding: *const f32, vec_id: u64) -> i32;
    pub fn vers_lsh_free(lsh: *mut vers_lsh) -> i32;
}
