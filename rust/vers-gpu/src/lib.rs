//! Drop-in `Index<N>` implementations over libvers_b200 (replaces vers/src/indexes/ivfflat.rs:152-214 and
//! vers/src/indexes/lsh.rs:254-283).  Serialized fields and their order are the reference's, so the trait's default
//! bincode save_index/load_index (vers/src/indexes/base.rs:31-58) keep working and files stay interchangeable.
use rand::Rng;
use serde::{Deserialize, Serialize};
use std::ffi::CStr;
use std::ptr::null_mut;
use vers::{Index, Vector};
use vers_cuda_sys as sys;

fn check(rc: i32) {
    if rc != sys::VERS_OK {
        // the reference unwraps / indexes out of bounds on this path: panic with the library's message
        let msg = unsafe { CStr::from_ptr(sys::vers_last_error()) }.to_string_lossy().into_owned();
        panic!("vers_b200 error {rc}: {msg}");
    }
}

struct Device { ctx: *mut sys::vers_ctx, ds: *mut sys::vers_dataset, ivf: *mut sys::vers_ivf }
unsafe impl Send for Device {}
unsafe impl Sync for Device {} // every ABI entry point takes the context mutex
impl Drop for Device {
    fn drop(&mut self) { unsafe { sys::vers_ivf_free(self.ivf); sys::vers_dataset_free(self.ds); sys::vers_ctx_destroy(self.ctx); } }
}

#[derive(Serialize, Deserialize)]
pub struct GpuIVFFlatIndex<const N: usize> {
    num_centroids: usize,
    values: Vec<Vector<N>>,
    centroids: Vec<Vector<N>>,
    assignments: Vec<usize>,
    ids: Vec<Vec<usize>>,
    #[serde(skip)]
    dev: std::sync::OnceLock<Device>,
}

impl<const N: usize> GpuIVFFlatIndex<N> {
    const STRIDE: u32 = (std::mem::size_of::<Vector<N>>() / 4) as u32; // repr(align(256)): 320 floats for N = 300

    pub fn build_index(num_clusters: usize, num_attempts: usize, max_iterations: usize, vectors: &Vec<Vector<N>>) -> Self {
        let mut rng = rand::thread_rng(); // the draws of initialize_centroids (ivfflat.rs:18-27), made on the host
        let init: Vec<u64> = (0..num_attempts * num_clusters).map(|_| rng.gen_range(0..vectors.len()) as u64).collect();
        let mut d = Device { ctx: null_mut(), ds: null_mut(), ivf: null_mut() };
        unsafe {
            check(sys::vers_ctx_create(0, &mut d.ctx));
            check(sys::vers_dataset_upload(d.ctx, vectors.as_ptr() as *const f32, vectors.len() as u64, N as u32, Self::STRIDE, 0, &mut d.ds));
            check(sys::vers_ivf_build_index(d.ds, num_clusters as u32, num_attempts as u32, max_iterations as u32, init.as_ptr(), &mut d.ivf));
        }
        let mut centroids = vec![Vector([0.0f32; N]); num_clusters];
        let mut a64 = vec![0u64; vectors.len()];
        unsafe {
            check(sys::vers_ivf_get_centroids(d.ivf, centroids.as_mut_ptr() as *mut f32, Self::STRIDE));
            check(sys::vers_ivf_get_assignments(d.ivf, a64.as_mut_ptr()));
        }
        let assignments: Vec<usize> = a64.iter().map(|&c| c as usize).collect();
        let mut ids = vec![vec![]; num_clusters];
        assignments.iter().enumerate().for_each(|(r, c)| ids[*c].push(r)); // ivfflat.rs:123-127
        let dev = std::sync::OnceLock::new();
        let _ = dev.set(d);
        GpuIVFFlatIndex { num_centroids: num_clusters, values: vectors.clone(), centroids, assignments, ids, dev }
    }

    fn device(&self) -> &Device {
        self.dev.get_or_init(|| {
            // after load_index: rebuild the device mirror from the deserialized fields
            let mut d = Device { ctx: null_mut(), ds: null_mut(), ivf: null_mut() };
            let a64: Vec<u64> = self.assignments.iter().map(|&c| c as u64).collect();
            unsafe {
                check(sys::vers_ctx_create(0, &mut d.ctx));
                check(sys::vers_dataset_upload(d.ctx, self.values.as_ptr() as *const f32, self.values.len() as u64, N as u32, Self::STRIDE, 0, &mut d.ds));
                check(sys::vers_ivf_from_parts(d.ds, self.centroids.as_ptr() as *const f32, self.num_centroids as u32, Self::STRIDE, a64.as_ptr(), &mut d.ivf));
            }
            d
        })
    }
}

impl<const N: usize> Index<N> for GpuIVFFlatIndex<N> {
    fn search_approximate(&self, query: Vector<N>, top_k: usize) -> Vec<(usize, f32)> {
        let (mut ids, mut d, mut cnt) = (vec![0u64; top_k.max(1)], vec![0f32; top_k.max(1)], 0u32);
        // nprobe = 0: the reference's nearest-list-plus-spill semantics (ivfflat.rs:163-197)
        unsafe { check(sys::vers_ivf_search(self.device().ivf, query.0.as_ptr(), 1, N as u32, top_k as u32, 0, ids.as_mut_ptr(), d.as_mut_ptr(), &mut cnt)); }
        ids.into_iter().zip(d).take(cnt as usize).map(|(i, d)| (i as usize, d)).collect()
    }

    fn add(&mut self, embedding: Vector<N>, vec_id: usize) {
        let (mut id, mut cl) = (0u64, 0u32);
        unsafe { check(sys::vers_ivf_add(self.device().ivf, embedding.0.as_ptr(), vec_id as u64, &mut id, &mut cl)); }
        self.values.push(embedding);
        self.assignments.push(cl as usize);
        self.ids[cl as usize].push(id as usize); // == assignments.len() before the push, like ivfflat.rs:209-212
    }
}
