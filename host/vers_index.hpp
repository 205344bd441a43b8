// vers_index.hpp — C++ host-side mirror of the reference's index interface over the C ABI (include/vers_device.h).
//
// The reference is compiled (Rust) code; with no Rust toolchain in this image the host side above the ABI is C++
// with the reference's names and argument meaning:
//   vers::Vector<N>                         indexes/base.rs:15-17   (alignas(256) [f32; N])
//   vers::Index<N>                          indexes/base.rs:27-59   add / search_approximate / save_index / load_index
//   vers::IVFFlatIndex<N>::build_index      indexes/ivfflat.rs:102-136
//   vers::ANNIndex<N>::build_index          indexes/lsh.rs:132-161
//   vers::search_exhaustive                 utils.rs:68-82
// Error behaviour: the reference panics; here every non-zero ABI status throws vers::Panic (std::runtime_error).
// save_index / load_index write/read the reference's bincode 1.3 layout of IVFFlatIndex (ivfflat.rs:9-15).
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
    void save_index(const std::string&) const override {
        throw std::runtime_error("ANNIndex::save_index: the recursive Node enum layout is not mirrored yet (SURVEY §8f)");
    }

  private:
    ANNIndex() = default;
    vers_lsh* lsh_ = nullptr;
};

}  // namespace vers
