// vers_index.hpp — C++ host-side mirror of the reference's index interface over the C ABI (include/vers_device.h).
//
// The reference is compiled (Rust) code; with no Rust toolchain in this image the host side above the ABI is C++
// with the reference's names and argument meaning:
//   vers::Vector<N>                         indexes/base.rs:15-17   (alignas(256) [f32; N])
//   vers::Index<N>                          indexes/base.rs:27-59   add / search_approximate / save_index / load_index
//   vers::IVFFlatIndex<N>::build_index      indexes/ivfflat.rs:102-136
//   vers::ANNIndex<N>::build_index          indexes/lsh.rs:132-161
//   vers::search_exhaustive                 utils.rs:68-82
//   vers::DeviceVectors<N>::cosine_similarity_simd   indexes/base.rs:158-223 (HNSW's distance, batched: hnsw.rs:146,258,273)
// Error behaviour: the reference panics; here every non-zero ABI status throws vers::Panic (std::runtime_error).
// save_index / load_index write/read the reference's bincode 1.3 layouts of IVFFlatIndex (ivfflat.rs:9-15) and
// ANNIndex (lsh.rs:13-55, the recursive Node enum included).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/vers_device.h"

namespace vers {

struct Panic : std::runtime_error {
    int32_t code;
    Panic(int32_t c, const std::string& m) : std::runtime_error("vers_b200 (" + std::to_string(c) + "): " + m), code(c) {}
};
inline void check(int32_t rc) {
    if (rc != VERS_OK) throw Panic(rc, vers_last_error());
}

template <size_t N>
struct alignas(256) Vector {
    float v[N];
    float& operator[](size_t i) { return v[i]; }
    const float& operator[](size_t i) const { return v[i]; }
};

struct Context {
    vers_ctx* h = nullptr;
    explicit Context(int device = 0) { check(vers_ctx_create(device, &h)); }
    ~Context() { vers_ctx_destroy(h); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
};

template <size_t N>
struct Index {
    virtual ~Index() = default;
    virtual void add(const Vector<N>& embedding, size_t vec_id) = 0;
    virtual std::vector<std::pair<size_t, float>> search_approximate(const Vector<N>& query, size_t top_k) const = 0;
    virtual void save_index(const std::string& file_path) const = 0;
};

template <size_t N>
constexpr uint32_t vector_stride() { return (uint32_t)(sizeof(Vector<N>) / sizeof(float)); }

// utils::search_exhaustive (utils.rs:68-82)
template <size_t N>
std::vector<std::pair<size_t, float>> search_exhaustive(Context& ctx, const std::vector<Vector<N>>& data,
                                                        const Vector<N>& query, size_t top_k) {
    vers_dataset* ds = nullptr;
    check(vers_dataset_upload(ctx.h, &data[0].v[0], data.size(), N, vector_stride<N>(), 0, &ds));
    std::vector<uint64_t> ids(top_k ? top_k : 1);
    std::vector<float> d(top_k ? top_k : 1);
    uint32_t cnt = 0;
    int32_t rc = vers_flat_search(ds, query.v, 1, N, (uint32_t)top_k, VERS_METRIC_L2SQ, ids.data(), d.data(), &cnt);
    vers_dataset_free(ds);
    check(rc);
    std::vector<std::pair<size_t, float>> out;
    for (uint32_t i = 0; i < cnt; ++i) out.emplace_back((size_t)ids[i], d[i]);
    return out;
}

// HNSW distance offload: the id -> vector map of hnsw.rs (id_to_vec) resident on the device; the CPU traversal hands it
// the neighbour ids of a step and gets Vector::cosine_similarity_simd(query, neighbour, true) for each, bit for bit
// (base.rs:158-223: 64-wide chunks, 4-wide chunks, scalar tail).  A missing id throws like id_to_vec.get(..).unwrap().
template <size_t N>
class DeviceVectors {
  public:
    DeviceVectors(Context& ctx, const std::vector<Vector<N>>& vectors, uint64_t first_id = 0) {
        check(vers_dataset_upload(ctx.h, &vectors[0].v[0], vectors.size(), N, vector_stride<N>(), first_id, &ds_));
    }
    ~DeviceVectors() { vers_dataset_free(ds_); }
    DeviceVectors(const DeviceVectors&) = delete;
    DeviceVectors& operator=(const DeviceVectors&) = delete;
    std::vector<float> cosine_similarity_simd(const Vector<N>& query, const std::vector<uint64_t>& neighbour_ids) const {
        std::vector<float> out(neighbour_ids.size());
        check(vers_pair_distances_simd(ds_, query.v, 1, N, nullptr, neighbour_ids.data(), neighbour_ids.size(),
                                       VERS_METRIC_COSINE, out.data()));
        return out;
    }
    std::vector<float> squared_euclidean_simd(const Vector<N>& query, const std::vector<uint64_t>& neighbour_ids) const {
        std::vector<float> out(neighbour_ids.size());
        check(vers_pair_distances_simd(ds_, query.v, 1, N, nullptr, neighbour_ids.data(), neighbour_ids.size(),
                                       VERS_METRIC_L2SQ, out.data()));
        return out;
    }

  private:
    vers_dataset* ds_ = nullptr;
};

template <size_t N>
class IVFFlatIndex : public Index<N> {
  public:
    // build_index(num_clusters, num_attempts, max_iterations, &vectors) (ivfflat.rs:102-107); init_rows are the
    // num_attempts*num_clusters row draws the reference takes from thread_rng (ivfflat.rs:18-27)
    static std::unique_ptr<IVFFlatIndex> build_index(Context& ctx, size_t num_clusters, size_t num_attempts,
                                                     size_t max_iterations, const std::vector<Vector<N>>& vectors,
                                                     const std::vector<uint64_t>& init_rows) {
        std::unique_ptr<IVFFlatIndex> ix(new IVFFlatIndex(ctx));
        ix->values_ = vectors;
        check(vers_dataset_upload(ctx.h, &vectors[0].v[0], vectors.size(), N, vector_stride<N>(), 0, &ix->ds_));
        check(vers_ivf_build_index(ix->ds_, (uint32_t)num_clusters, (uint32_t)num_attempts, (uint32_t)max_iterations,
                                   init_rows.data(), &ix->ivf_));
        ix->num_centroids_ = num_clusters;
        return ix;
    }
    // Index::load_index (base.rs:45-58) + device mirror rebuild
    static std::unique_ptr<IVFFlatIndex> load_index(Context& ctx, const std::string& file_path) {
        std::ifstream f(file_path, std::ios::binary);
        if (!f) throw std::runtime_error("cannot open " + file_path);
        auto rd64 = [&]() { uint64_t x = 0; f.read((char*)&x, 8); return x; };
        std::unique_ptr<IVFFlatIndex> ix(new IVFFlatIndex(ctx));
        ix->num_centroids_ = rd64();
        uint64_t nv = rd64();
        ix->values_.resize(nv);
        for (auto& v : ix->values_) f.read((char*)v.v, N * 4);
        uint64_t nc = rd64();
        std::vector<Vector<N>> cents(nc);
        for (auto& v : cents) f.read((char*)v.v, N * 4);
        uint64_t na = rd64();
        std::vector<uint64_t> assign(na);
        f.read((char*)assign.data(), na * 8);
        if (!f) throw std::runtime_error("Deserialization error: truncated " + file_path);
        check(vers_dataset_upload(ctx.h, &ix->values_[0].v[0], nv, N, vector_stride<N>(), 0, &ix->ds_));
        check(vers_ivf_from_parts(ix->ds_, &cents[0].v[0], (uint32_t)nc, vector_stride<N>(), assign.data(), &ix->ivf_));
        return ix;
    }
    ~IVFFlatIndex() override {
        vers_ivf_free(ivf_);
        vers_dataset_free(ds_);
    }
    void add(const Vector<N>& embedding, size_t vec_id) override {  // ivfflat.rs:200-213 (vec_id ignored there too)
        uint64_t id = 0;
        uint32_t cl = 0;
        check(vers_ivf_add(ivf_, embedding.v, vec_id, &id, &cl));
        values_.push_back(embedding);
    }
    std::vector<std::pair<size_t, float>> search_approximate(const Vector<N>& query, size_t top_k) const override {
        std::vector<uint64_t> ids(top_k ? top_k : 1);
        std::vector<float> d(top_k ? top_k : 1);
        uint32_t cnt = 0;
        check(vers_ivf_search(ivf_, query.v, 1, N, (uint32_t)top_k, 0, ids.data(), d.data(), &cnt));
        std::vector<std::pair<size_t, float>> out;
        for (uint32_t i = 0; i < cnt; ++i) out.emplace_back((size_t)ids[i], d[i]);
        return out;
    }
    // Extension the trait lacks (BASELINE config 4): a batch of queries, global top_k by (distance, id) over the nprobe
    // nearest lists of each query; nprobe == 0 keeps the reference's nearest-list-plus-spill semantics per query.
    std::vector<std::vector<std::pair<size_t, float>>> search_batch(const std::vector<Vector<N>>& queries, size_t top_k,
                                                                    size_t nprobe) const {
        const size_t nq = queries.size(), kk = top_k ? top_k : 1;
        std::vector<uint64_t> ids(nq * kk);
        std::vector<float> d(nq * kk);
        std::vector<uint32_t> cnt(nq ? nq : 1);
        if (nq)
            check(vers_ivf_search(ivf_, &queries[0].v[0], (uint32_t)nq, vector_stride<N>(), (uint32_t)top_k,
                                  (uint32_t)nprobe, ids.data(), d.data(), cnt.data()));
        std::vector<std::vector<std::pair<size_t, float>>> out(nq);
        for (size_t q = 0; q < nq; ++q)
            for (uint32_t i = 0; i < cnt[q]; ++i) out[q].emplace_back((size_t)ids[q * top_k + i], d[q * top_k + i]);
        return out;
    }
    std::vector<uint64_t> assignments() const {
        std::vector<uint64_t> a(values_.size());
        check(vers_ivf_get_assignments(ivf_, a.data()));
        return a;
    }
    std::vector<Vector<N>> centroids() const {
        std::vector<Vector<N>> c(num_centroids_);
        check(vers_ivf_get_centroids(ivf_, &c[0].v[0], vector_stride<N>()));
        return c;
    }
    // Index::save_index (base.rs:31-43): bincode 1.3, fields in declaration order (ivfflat.rs:9-15)
    void save_index(const std::string& file_path) const override {
        std::ofstream f(file_path, std::ios::binary);
        if (!f) throw std::runtime_error("cannot create " + file_path);
        auto wr64 = [&](uint64_t x) { f.write((const char*)&x, 8); };
        auto cents = centroids();
        auto assign = assignments();
        wr64(num_centroids_);
        wr64(values_.size());
        for (const auto& v : values_) f.write((const char*)v.v, N * 4);
        wr64(cents.size());
        for (const auto& v : cents) f.write((const char*)v.v, N * 4);
        wr64(assign.size());
        f.write((const char*)assign.data(), assign.size() * 8);
        std::vector<std::vector<uint64_t>> ids(num_centroids_);
        for (size_t r = 0; r < assign.size(); ++r) ids[assign[r]].push_back(r);  // ivfflat.rs:123-127
        wr64(ids.size());
        for (const auto& l : ids) {
            wr64(l.size());
            f.write((const char*)l.data(), l.size() * 8);
        }
    }
    size_t len() const { return values_.size(); }

  private:
    explicit IVFFlatIndex(Context& ctx) : ctx_(ctx) {}
    Context& ctx_;
    size_t num_centroids_ = 0;
    std::vector<Vector<N>> values_;
    vers_dataset* ds_ = nullptr;
    vers_ivf* ivf_ = nullptr;
};

template <size_t N>
class ANNIndex : public Index<N> {
  public:
    // build_index(num_trees, max_size, &vectors, &vector_ids) (lsh.rs:132-137); seed feeds the injected sample pairs
    static std::unique_ptr<ANNIndex> build_index(Context& ctx, size_t num_trees, size_t max_size,
                                                 const std::vector<Vector<N>>& vectors,
                                                 const std::vector<uint64_t>& vector_ids, uint64_t seed) {
        std::unique_ptr<ANNIndex> ix(new ANNIndex());
        check(vers_lsh_build_index(ctx.h, &vectors[0].v[0], vectors.size(), N, vector_stride<N>(), vector_ids.data(),
                                   (uint32_t)num_trees, (uint32_t)max_size, seed, &ix->lsh_));
        ix->max_node_size_ = max_size;
        return ix;
    }
    ~ANNIndex() override { vers_lsh_free(lsh_); }
    void add(const Vector<N>& embedding, size_t vec_id) override { check(vers_lsh_add(lsh_, embedding.v, vec_id)); }
    std::vector<std::pair<size_t, float>> search_approximate(const Vector<N>& query, size_t top_k) const override {
        std::vector<uint64_t> ids(top_k ? top_k : 1);
        std::vector<float> d(top_k ? top_k : 1);
        uint32_t cnt = 0;
        check(vers_lsh_search(lsh_, query.v, 1, N, (uint32_t)top_k, ids.data(), d.data(), &cnt));
        std::vector<std::pair<size_t, float>> out;
        for (uint32_t i = 0; i < cnt; ++i) out.emplace_back((size_t)ids[i], d[i]);
        return out;
    }
    // Index::save_index (base.rs:31-43): bincode 1.3 layout of ANNIndex (lsh.rs:13-55):
    //   max_node_size u64 | trees: u64 len, every tree as the recursive enum Node { Inner(Box<InnerNode>) = variant 0:
    //   coefficients N f32, constant f32, left_node, right_node | Leaf(Box<LeafNode(Vec<usize>)>) = variant 1: u64 len,
    //   len u64 } | values: u64 len, len * N f32 | ids: u64 len, len u64
    void save_index(const std::string& file_path) const override {
        uint64_t nv = 0, nn_all = 0;
        uint32_t nt = 0;
        check(vers_lsh_info(lsh_, &nv, &nt, &nn_all));
        std::ofstream f(file_path, std::ios::binary);
        if (!f) throw std::runtime_error("save_index: cannot open " + file_path);
        auto put64 = [&](uint64_t v) { f.write(reinterpret_cast<const char*>(&v), 8); };
        auto put32 = [&](uint32_t v) { f.write(reinterpret_cast<const char*>(&v), 4); };
        put64(max_node_size_);
        put64(nt);
        for (uint32_t t = 0; t < nt; ++t) {
            uint32_t nn = 0, ni = 0;
            uint64_t nit = 0;
            check(vers_lsh_flatten(lsh_, t, nullptr, nullptr, nullptr, nullptr, nullptr, &nn, &ni, &nit));
            std::vector<uint8_t> kind(nn);
            std::vector<uint32_t> leaf_len(nn), items(nit ? nit : 1);
            std::vector<float> planes((size_t)ni * N + 1), consts(ni + 1);
            check(vers_lsh_flatten(lsh_, t, kind.data(), leaf_len.data(), planes.data(), consts.data(), items.data(), &nn,
                                   &ni, &nit));
            // the flattened preorder is (node, ABOVE subtree, BELOW subtree); the file wants (node, left = below, right =
            // above): per node its plane / item offset and the end of its subtree, then an explicit-stack emit
            std::vector<uint32_t> plane_of(nn, 0), end(nn, 0);
            std::vector<uint64_t> item_off(nn, 0);
            uint32_t pi = 0;
            uint64_t io = 0;
            for (uint32_t i = 0; i < nn; ++i) {
                if (kind[i] == 0) plane_of[i] = pi++;
                else { item_off[i] = io; io += leaf_len[i]; }
            }
            std::vector<std::pair<uint32_t, int>> open_nodes;
            for (uint32_t i = 0; i < nn; ++i) {
                if (kind[i] == 0) { open_nodes.emplace_back(i, 2); continue; }
                end[i] = i + 1;
                while (!open_nodes.empty() && --open_nodes.back().second == 0) {
                    end[open_nodes.back().first] = i + 1;
                    open_nodes.pop_back();
                }
            }
            std::vector<uint32_t> todo{0};
            while (!todo.empty()) {
                const uint32_t i = todo.back();
                todo.pop_back();
                if (kind[i] == 1) {
                    put32(1);
                    put64(leaf_len[i]);
                    for (uint32_t e = 0; e < leaf_len[i]; ++e) put64(items[item_off[i] + e]);
                } else {
                    put32(0);
                    f.write(reinterpret_cast<const char*>(&planes[(size_t)plane_of[i] * N]), N * 4);
                    f.write(reinterpret_cast<const char*>(&consts[plane_of[i]]), 4);
                    todo.push_back(i + 1);       // above = right_node: emitted second
                    todo.push_back(end[i + 1]);  // below = left_node: emitted first
                }
            }
        }
        std::vector<float> values((size_t)nv * N + 1);
        std::vector<uint64_t> ids(nv + 1);
        check(vers_lsh_get_values(lsh_, values.data(), N, ids.data()));
        put64(nv);
        f.write(reinterpret_cast<const char*>(values.data()), (std::streamsize)((size_t)nv * N * 4));
        put64(nv);
        f.write(reinterpret_cast<const char*>(ids.data()), (std::streamsize)(nv * 8));
        if (!f) throw std::runtime_error("save_index: write failed");
    }
    // Index::load_index (base.rs:45-58) + the device forest (vers_lsh_from_parts); seed feeds later leaf splits
    static std::unique_ptr<ANNIndex> load_index(Context& ctx, const std::string& file_path, uint64_t seed = 4) {
        std::ifstream f(file_path, std::ios::binary);
        if (!f) throw std::runtime_error("load_index: cannot open " + file_path);
        auto get64 = [&]() { uint64_t v = 0; f.read(reinterpret_cast<char*>(&v), 8); return v; };
        auto get32 = [&]() { uint32_t v = 0; f.read(reinterpret_cast<char*>(&v), 4); return v; };
        std::unique_ptr<ANNIndex> ix(new ANNIndex());
        ix->max_node_size_ = get64();
        const uint64_t nt = get64();
        std::vector<uint32_t> tree_nodes, leaf_len, items;
        std::vector<uint8_t> kind;
        std::vector<float> planes, consts;
        // the file is (node, left = below, right = above); vers_lsh_from_parts wants (node, above, below): parse each
        // tree into node records first, then walk it in the other order
        struct Rec { uint8_t kind; uint32_t left, right; size_t plane, item0; uint32_t len; };
        for (uint64_t t = 0; t < nt; ++t) {
            std::vector<Rec> recs;
            std::vector<float> tp, tc;
            std::vector<uint32_t> ti;
            std::vector<std::pair<uint32_t, int>> open_nodes;  // (inner node, children attached so far)
            auto attach = [&](uint32_t child) {
                while (true) {
                    if (open_nodes.empty()) return;
                    auto& top = open_nodes.back();
                    if (top.second == 0) { recs[top.first].left = child; top.second = 1; return; }
                    recs[top.first].right = child;
                    open_nodes.pop_back();
                    return;
                }
            };
            do {
                const uint32_t variant = get32();
                if (!f) throw std::runtime_error("load_index: truncated tree");
                const uint32_t me = (uint32_t)recs.size();
                if (variant == 1) {
                    const uint64_t len = get64();
                    recs.push_back(Rec{1, 0, 0, 0, ti.size(), (uint32_t)len});
                    for (uint64_t e = 0; e < len; ++e) ti.push_back((uint32_t)get64());
                    if (me) attach(me);
                } else if (variant == 0) {
                    recs.push_back(Rec{0, 0, 0, tc.size(), 0, 0});
                    tp.resize(tp.size() + N);
                    f.read(reinterpret_cast<char*>(&tp[tp.size() - N]), N * 4);
                    float c = 0.f;
                    f.read(reinterpret_cast<char*>(&c), 4);
                    tc.push_back(c);
                    if (me) attach(me);
                    open_nodes.emplace_back(me, 0);
                } else {
                    throw std::runtime_error("load_index: bad Node variant (wrong N or not an ANNIndex file)");
                }
            } while (!open_nodes.empty());
            const size_t n0 = kind.size();
            std::vector<uint32_t> todo{0};
            while (!todo.empty()) {
                const Rec& r = recs[todo.back()];
                todo.pop_back();
                kind.push_back(r.kind);
                leaf_len.push_back(r.len);
                if (r.kind == 1) {
                    items.insert(items.end(), ti.begin() + (long)r.item0, ti.begin() + (long)(r.item0 + r.len));
                } else {
                    planes.insert(planes.end(), tp.begin() + (long)(r.plane * N), tp.begin() + (long)((r.plane + 1) * N));
                    consts.push_back(tc[r.plane]);
                    todo.push_back(r.left);   // below: visited second
                    todo.push_back(r.right);  // above: visited first
                }
            }
            tree_nodes.push_back((uint32_t)(kind.size() - n0));
        }
        const uint64_t nv = get64();
        std::vector<float> values((size_t)nv * N + 1);
        f.read(reinterpret_cast<char*>(values.data()), (std::streamsize)((size_t)nv * N * 4));
        const uint64_t ni = get64();
        std::vector<uint64_t> ids(ni + 1);
        f.read(reinterpret_cast<char*>(ids.data()), (std::streamsize)(ni * 8));
        if (!f || ni != nv) throw std::runtime_error("load_index: truncated file");
        planes.push_back(0.f), consts.push_back(0.f), items.push_back(0), kind.push_back(0), leaf_len.push_back(0);
        check(vers_lsh_from_parts(ctx.h, values.data(), nv, N, N, ids.data(), (uint32_t)nt, (uint32_t)ix->max_node_size_, seed,
                                  tree_nodes.data(), kind.data(), leaf_len.data(), planes.data(), consts.data(),
                                  items.data(), &ix->lsh_));
        return ix;
    }

  private:
    ANNIndex() = default;
    vers_lsh* lsh_ = nullptr;
    uint64_t max_node_size_ = 0;
};

}  // namespace vers
