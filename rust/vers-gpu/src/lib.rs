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

// ================================================================================================ GpuANNIndex
// The reference's ANNIndex (vers/src/indexes/lsh.rs:47-55) with its private tree types mirrored field for field, so
// that serde/bincode produce and accept the reference's own files (variant 0 = Inner, 1 = Leaf; lsh.rs:13-44).
#[derive(Serialize, Deserialize)]
struct Hyperplane<const N: usize> { coefficients: Vector<N>, constant: f32 }
#[derive(Serialize, Deserialize)]
enum Node<const N: usize> { Inner(Box<InnerNode<N>>), Leaf(Box<LeafNode>) }
#[derive(Serialize, Deserialize)]
struct InnerNode<const N: usize> { hyperplane: Hyperplane<N>, left_node: Node<N>, right_node: Node<N> }
#[derive(Serialize, Deserialize)]
struct LeafNode(Vec<usize>);

struct LshDevice { ctx: *mut sys::vers_ctx, lsh: *mut sys::vers_lsh }
unsafe impl Send for LshDevice {}
unsafe impl Sync for LshDevice {}
impl Drop for LshDevice {
    fn drop(&mut self) { unsafe { sys::vers_lsh_free(self.lsh); sys::vers_ctx_destroy(self.ctx); } }
}

#[derive(Serialize, Deserialize)]
pub struct GpuANNIndex<const N: usize> {
    max_node_size: usize,
    trees: Vec<Node<N>>,
    values: Vec<Vector<N>>,
    ids: Vec<usize>,
    #[serde(skip)]
    dev: std::sync::OnceLock<LshDevice>,
}

/// the flattened preorder the device layer speaks (node, ABOVE = right_node, BELOW = left_node)
#[derive(Default)]
struct Flat { kind: Vec<u8>, leaf_len: Vec<u32>, planes: Vec<f32>, consts: Vec<f32>, items: Vec<u32> }

impl<const N: usize> GpuANNIndex<N> {
    const STRIDE: u32 = (std::mem::size_of::<Vector<N>>() / 4) as u32;

    fn flatten(node: &Node<N>, f: &mut Flat) {
        match node {
            Node::Leaf(l) => { f.kind.push(1); f.leaf_len.push(l.0.len() as u32); f.items.extend(l.0.iter().map(|&i| i as u32)); }
            Node::Inner(i) => {
                f.kind.push(0); f.leaf_len.push(0);
                f.planes.extend_from_slice(&i.hyperplane.coefficients.0); f.consts.push(i.hyperplane.constant);
                Self::flatten(&i.right_node, f); // above subtree first
                Self::flatten(&i.left_node, f);
            }
        }
    }

    fn unflatten(f: &Flat, at: &mut usize, plane: &mut usize, item: &mut usize) -> Node<N> {
        let i = *at; *at += 1;
        if f.kind[i] == 1 {
            let len = f.leaf_len[i] as usize;
            let v = f.items[*item..*item + len].iter().map(|&x| x as usize).collect(); *item += len;
            Node::Leaf(Box::new(LeafNode(v)))
        } else {
            let mut c = [0.0f32; N]; c.copy_from_slice(&f.planes[*plane * N..(*plane + 1) * N]);
            let constant = f.consts[*plane]; *plane += 1;
            let right_node = Self::unflatten(f, at, plane, item); // above was emitted first
            let left_node = Self::unflatten(f, at, plane, item);
            Node::Inner(Box::new(InnerNode { hyperplane: Hyperplane { coefficients: Vector(c), constant }, left_node, right_node }))
        }
    }

    /// `build_index(num_trees, max_size, &vectors, &vector_ids)` (lsh.rs:132-161): dedup, tree construction and
    /// hashing run on the device; the struct fields are read back so that save_index serialises the reference layout.
    pub fn build_index(num_trees: usize, max_size: usize, vectors: &Vec<Vector<N>>, vector_ids: &Vec<usize>) -> Self {
        let seed: u64 = rand::thread_rng().gen(); // stands in for the thread_rng draws of build_hyperplane (lsh.rs:63-65)
        let ids64: Vec<u64> = vector_ids.iter().map(|&i| i as u64).collect();
        let mut d = LshDevice { ctx: null_mut(), lsh: null_mut() };
        let (mut nv, mut nt, mut nn) = (0u64, 0u32, 0u64);
        unsafe {
            check(sys::vers_ctx_create(0, &mut d.ctx));
            check(sys::vers_lsh_build_index(d.ctx, vectors.as_ptr() as *const f32, vectors.len() as u64, N as u32, Self::STRIDE,
                                            ids64.as_ptr(), num_trees as u32, max_size as u32, seed, &mut d.lsh));
            check(sys::vers_lsh_info(d.lsh, &mut nv, &mut nt, &mut nn));
        }
        let mut values = vec![Vector([0.0f32; N]); nv as usize];
        let mut ids = vec![0u64; nv as usize];
        unsafe { check(sys::vers_lsh_get_values(d.lsh, values.as_mut_ptr() as *mut f32, Self::STRIDE, ids.as_mut_ptr())); }
        let trees = (0..nt).map(|t| Self::read_tree(d.lsh, t)).collect();
        let dev = std::sync::OnceLock::new();
        let _ = dev.set(d);
        GpuANNIndex { max_node_size: max_size, trees, values, ids: ids.into_iter().map(|i| i as usize).collect(), dev }
    }

    fn read_tree(lsh: *mut sys::vers_lsh, t: u32) -> Node<N> {
        let (mut nn, mut ni, mut nit) = (0u32, 0u32, 0u64);
        let mut f = Flat::default();
        unsafe {
            check(sys::vers_lsh_flatten(lsh, t, null_mut(), null_mut(), null_mut(), null_mut(), null_mut(), &mut nn, &mut ni, &mut nit));
            f.kind = vec![0; nn as usize]; f.leaf_len = vec![0; nn as usize];
            f.planes = vec![0.0; ni as usize * N + 1]; f.consts = vec![0.0; ni as usize + 1]; f.items = vec![0; nit as usize + 1];
            check(sys::vers_lsh_flatten(lsh, t, f.kind.as_mut_ptr(), f.leaf_len.as_mut_ptr(), f.planes.as_mut_ptr(),
                                        f.consts.as_mut_ptr(), f.items.as_mut_ptr(), &mut nn, &mut ni, &mut nit));
        }
        Self::unflatten(&f, &mut 0, &mut 0, &mut 0)
    }

    fn device(&self) -> &LshDevice {
        self.dev.get_or_init(|| {
            // after load_index: rebuild the device forest from the deserialized fields (vers_lsh_from_parts)
            let mut f = Flat::default();
            let mut tree_nodes = vec![];
            for t in &self.trees { let n0 = f.kind.len(); Self::flatten(t, &mut f); tree_nodes.push((f.kind.len() - n0) as u32); }
            let ids64: Vec<u64> = self.ids.iter().map(|&i| i as u64).collect();
            f.planes.push(0.0); f.consts.push(0.0); f.items.push(0); // never empty pointers
            let mut d = LshDevice { ctx: null_mut(), lsh: null_mut() };
            unsafe {
                check(sys::vers_ctx_create(0, &mut d.ctx));
                check(sys::vers_lsh_from_parts(d.ctx, self.values.as_ptr() as *const f32, self.values.len() as u64, N as u32,
                                               Self::STRIDE, ids64.as_ptr(), self.trees.len() as u32, self.max_node_size as u32,
                                               rand::thread_rng().gen(), tree_nodes.as_ptr(), f.kind.as_ptr(), f.leaf_len.as_ptr(),
                                               f.planes.as_ptr(), f.consts.as_ptr(), f.items.as_ptr(), &mut d.lsh));
            }
            d
        })
    }
}

impl<const N: usize> Index<N> for GpuANNIndex<N> {
    fn search_approximate(&self, query: Vector<N>, top_k: usize) -> Vec<(usize, f32)> {
        let (mut ids, mut d, mut cnt) = (vec![0u64; top_k.max(1)], vec![0f32; top_k.max(1)], 0u32);
        unsafe { check(sys::vers_lsh_search(self.device().lsh, query.0.as_ptr(), 1, N as u32, top_k as u32, ids.as_mut_ptr(), d.as_mut_ptr(), &mut cnt)); }
        ids.into_iter().zip(d).take(cnt as usize).map(|(i, d)| (i as usize, d)).collect()
    }

    fn add(&mut self, embedding: Vector<N>, vec_id: usize) {
        // lsh.rs:255-263: values.push, ids.push, insert into every tree (a leaf that overflows is rebuilt as a subtree)
        let lsh = self.device().lsh;
        unsafe { check(sys::vers_lsh_add(lsh, embedding.0.as_ptr(), vec_id as u64)); }
        self.values.push(embedding);
        self.ids.push(vec_id);
        // the trees changed on the device (appended leaf member or a split): refresh the serialisable mirror
        self.trees = (0..self.trees.len() as u32).map(|t| Self::read_tree(lsh, t)).collect();
    }
}

/// HNSW distance offload (hnsw.rs:146, 258, 273 call `Vector::cosine_similarity_simd`, base.rs:158-223): the
/// `id_to_vec` map of `HNSWIndex` resident on the device.  The graph build and traversal stay in `hnsw.rs`; where it
/// evaluates a node against a list of neighbours it can hand the ids of the whole list (or of many queries' lists) to
/// `cosine_similarity_simd_batch` and gets the reference's values bit for bit (same 64-wide / 4-wide / scalar chunking).
pub struct DeviceVectors<const N: usize> { ctx: *mut sys::vers_ctx, ds: *mut sys::vers_dataset }
unsafe impl<const N: usize> Send for DeviceVectors<N> {}
unsafe impl<const N: usize> Sync for DeviceVectors<N> {}
impl<const N: usize> Drop for DeviceVectors<N> {
    fn drop(&mut self) { unsafe { sys::vers_dataset_free(self.ds); sys::vers_ctx_destroy(self.ctx); } }
}

impl<const N: usize> DeviceVectors<N> {
    /// `vectors[i]` has id `first_id + i` (HNSWIndex::build_index numbers its nodes 0..n, hnsw.rs:447-452)
    pub fn new(vectors: &Vec<Vector<N>>, first_id: usize) -> Self {
        let mut d = DeviceVectors { ctx: null_mut(), ds: null_mut() };
        unsafe {
            check(sys::vers_ctx_create(0, &mut d.ctx));
            check(sys::vers_dataset_upload(d.ctx, vectors.as_ptr() as *const f32, vectors.len() as u64, N as u32,
                                           (std::mem::size_of::<Vector<N>>() / 4) as u32, first_id as u64, &mut d.ds));
        }
        d
    }

    /// `query.cosine_similarity_simd(&id_to_vec[id], true)` for every id; panics on a missing id like `.unwrap()`
    pub fn cosine_similarity_simd_batch(&self, query: &Vector<N>, neighbour_ids: &[usize]) -> Vec<f32> {
        let ids: Vec<u64> = neighbour_ids.iter().map(|&i| i as u64).collect();
        let mut out = vec![0f32; ids.len()];
        unsafe {
            check(sys::vers_pair_distances_simd(self.ds, query.0.as_ptr(), 1, N as u32, std::ptr::null(), ids.as_ptr(),
                                                ids.len() as u64, sys::VERS_METRIC_COSINE, out.as_mut_ptr()));
        }
        out
    }

    /// many (query, neighbour) pairs in one call: `pairs[i] = (index into queries, neighbour id)`
    pub fn cosine_similarity_simd_pairs(&self, queries: &[Vector<N>], pairs: &[(u32, usize)]) -> Vec<f32> {
        let pq: Vec<u32> = pairs.iter().map(|p| p.0).collect();
        let pr: Vec<u64> = pairs.iter().map(|p| p.1 as u64).collect();
        let mut out = vec![0f32; pairs.len()];
        unsafe {
            check(sys::vers_pair_distances_simd(self.ds, queries.as_ptr() as *const f32, queries.len() as u32,
                                                (std::mem::size_of::<Vector<N>>() / 4) as u32, pq.as_ptr(), pr.as_ptr(),
                                                pairs.len() as u64, sys::VERS_METRIC_COSINE, out.as_mut_ptr()));
        }
        out
    }
}
