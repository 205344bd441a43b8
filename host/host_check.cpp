// host_check.cpp — exercises host/vers_index.hpp on a GPU: build, search, add, save/load round trip.
// Built and run by tests/test_gpu_parity.py::test_cpp_host_mirror (prints "host_check ok").
#include <cstdio>
#include <cstdlib>
#include <iterator>
#include <string>

#include "vers_index.hpp"
#include "../include/vers_synth.h"

using namespace vers;
constexpr size_t N = 300;

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "/tmp/host_check_ivf.bin";
    const size_t n = 5000, C = 16, k = 10;
    static_assert(sizeof(Vector<N>) == 1280, "size_of::<Vector<300>>() in the reference (repr(align(256)))");
    std::vector<Vector<N>> rows(n);
    for (size_t r = 0; r < n; ++r)
        for (size_t c = 0; c < N; ++c) rows[r][c] = vers_synth_elem(1, 7, VERS_SYNTH_CLUSTERED, 16, r, (uint32_t)c, N);
    std::vector<uint64_t> init(C);
    for (size_t j = 0; j < C; ++j) init[j] = vers_synth_init_row(3, 0, (uint32_t)j, n);
    Context ctx(0);
    auto ix = IVFFlatIndex<N>::build_index(ctx, C, 1, 10, rows, init);
    auto exact = search_exhaustive(ctx, rows, rows[42], k);
    auto approx = ix->search_approximate(rows[42], k);
    if (exact.size() != k || approx.size() != k || exact[0].first != 42 || approx[0].first != 42 || approx[0].second != 0.0f) {
        std::printf("unexpected neighbours\n");
        return 1;
    }
    Vector<N> extra = rows[7];
    extra[0] += 0.125f;
    ix->add(extra, 999999);
    auto near = ix->search_approximate(extra, 2);
    if (near[0].first != n) {  // the reference ignores vec_id and stores assignments.len()
        std::printf("add: stored id %zu, expected %zu\n", near[0].first, n);
        return 1;
    }
    ix->save_index(path);
    auto ix2 = IVFFlatIndex<N>::load_index(ctx, path);
    auto again = ix2->search_approximate(rows[42], k);
    for (size_t i = 0; i < k; ++i)
        if (again[i] != approx[i]) {
            std::printf("save/load changed result %zu\n", i);
            return 1;
        }
    std::vector<uint64_t> ids(n);
    for (size_t i = 0; i < n; ++i) ids[i] = i;
    auto ann = ANNIndex<N>::build_index(ctx, 4, 100, rows, ids, 4);
    auto l = ann->search_approximate(rows[42], k);
    if (l.empty() || l[0].first != 42) {
        std::printf("lsh: nearest of a member is not itself\n");
        return 1;
    }
    // ANNIndex: add, save_index, load_index (the reference's run_test order) — same neighbours, byte-identical re-save
    Vector<N> extra2 = rows[9];
    extra2[1] += 0.25f;
    ann->add(extra2, n);
    const std::string apath = std::string(path) + ".ann";
    ann->save_index(apath);
    auto ann2 = ANNIndex<N>::load_index(ctx, apath, 4);
    auto l2 = ann2->search_approximate(rows[42], k);
    auto l1 = ann->search_approximate(rows[42], k);
    if (l1 != l2 || ann2->search_approximate(extra2, 1)[0].first != n) {
        std::printf("lsh: save/load changed the result\n");
        return 1;
    }
    ann2->save_index(apath + "2");
    {
        std::ifstream a(apath, std::ios::binary), b(apath + "2", std::ios::binary);
        std::string sa((std::istreambuf_iterator<char>(a)), std::istreambuf_iterator<char>());
        std::string sb((std::istreambuf_iterator<char>(b)), std::istreambuf_iterator<char>());
        if (sa.empty() || sa != sb) {
            std::printf("lsh: re-saved index differs\n");
            return 1;
        }
    }
    {   // HNSW distance offload == the SIMD-order cosine distance of base.rs:158-223, bit for bit
        DeviceVectors<N> dv(ctx, rows);
        std::vector<uint64_t> nb = {7, 42, 0, n - 1, 13};
        auto got = dv.cosine_similarity_simd(rows[5], nb);
        for (size_t j = 0; j < nb.size(); ++j) {
            const float* u = rows[5].v;
            const float* v = rows[nb[j]].v;
            volatile float res = 0.0f;  // volatile: keep every add separately rounded, in this order
            size_t i = 0;
            for (; i + 64 <= N; i += 64) {
                volatile float s = 0.0f;
                for (size_t e = 0; e < 64; ++e) { volatile float p = u[i + e] * v[i + e]; s = s + p; }
                res = res + s;
            }
            for (; i + 4 <= N; i += 4) {
                volatile float s = 0.0f;
                for (size_t e = 0; e < 4; ++e) { volatile float p = u[i + e] * v[i + e]; s = s + p; }
                res = res + s;
            }
            for (; i < N; ++i) { volatile float p = u[i] * v[i]; res = res + p; }
            const float want = 1.0f - res;
            if (std::memcmp(&want, &got[j], 4) != 0) {
                std::printf("hnsw distance %zu: %a vs %a\n", j, (double)got[j], (double)want);
                return 1;
            }
        }
    }
    bool threw = false;
    try {
        ix->search_approximate(rows[0], 200);  // > VERS_MAX_TOPK
    } catch (const Panic&) {
        threw = true;
    }
    if (!threw) return 1;
    std::printf("host_check ok\n");
    return 0;
}
