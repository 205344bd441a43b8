/*
 * vers_synth.h — the specification of the synthetic inputs used by the parity tests and bench.py.
 *
 * The reference (ashrielbrian/vers) ships no dataset and no seeded RNG (rand::thread_rng everywhere:
 * vers/src/indexes/ivfflat.rs:19, vers/src/indexes/lsh.rs:64), so parity is only definable conditional on
 * injected inputs.  This header defines those inputs as pure counter-based functions so that the CPU oracle
 * (oracle/) and the device library (vers_b200/csrc/) regenerate bit-identical rows without any bulk copy:
 *
 *   unit(seed, i)        = ((splitmix64(splitmix64(seed) + i) >> 40) - 2^23) / 2^23     in [-1, 1), exact in fp32
 *   kind 0 "uniform"     : x[r][c] = unit(seed, r*dim + c)
 *   kind 1 "clustered"   : x[r][c] = unit(center_seed, center(r)*dim + c) + 0.25f * unit(seed, r*dim + c)
 *                          center(r) = splitmix64(center_seed ^ 0x5bd1e995 ^ (r * 0x9E3779B97F4A7C15)) % n_centers
 *                          (one IEEE add; 0.25f*u is exact, so an fma gives the same bits)
 *   k-means init rows    : init_row(seed, attempt, j, n) = splitmix64(seed + attempt*2^32 + j) % n  (WITH replacement,
 *                          like gen_range in ivfflat.rs:23)
 *   LSH sample pair      : vers_lsh_pick_pair — two distinct member positions keyed by (seed, tree, node path hash)
 *                          (stands in for choose_multiple(thread_rng, 2) at lsh.rs:63-65)
 *
 * Everything here is integer arithmetic plus one exact int->float conversion: no libm, no rounding ambiguity.
 */
#ifndef VERS_SYNTH_H
#define VERS_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define VERS_HD __host__ __device__ __forceinline__
#else
#define VERS_HD static inline
#endif

#define VERS_SYNTH_UNIFORM 0u
#define VERS_SYNTH_CLUSTERED 1u

VERS_HD uint64_t vers_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* seed is pre-mixed once by the caller: pass vers_splitmix64(seed) as mixed_seed */
VERS_HD float vers_synth_unit_mixed(uint64_t mixed_seed, uint64_t i) {
    uint64_t h = vers_splitmix64(mixed_seed + i);
    int32_t v = (int32_t)(h >> 40) - 8388608; /* [-2^23, 2^23) */
    return (float)v * (1.0f / 8388608.0f);
}

VERS_HD float vers_synth_unit(uint64_t seed, uint64_t i) {
    return vers_synth_unit_mixed(vers_splitmix64(seed), i);
}

VERS_HD uint64_t vers_synth_center_of(uint64_t center_seed, uint64_t r, uint32_t n_centers) {
    return vers_splitmix64(center_seed ^ 0x5bd1e995ull ^ (r * 0x9E3779B97F4A7C15ull)) % (uint64_t)n_centers;
}

/* one element of the synthetic matrix (before normalisation) */
VERS_HD float vers_synth_elem(uint64_t seed, uint64_t center_seed, uint32_t kind, uint32_t n_centers,
                              uint64_t r, uint32_t c, uint32_t dim) {
    float u = vers_synth_unit(seed, r * (uint64_t)dim + c);
    if (kind == VERS_SYNTH_CLUSTERED) {
        uint64_t ctr = vers_synth_center_of(center_seed, r, n_centers);
        float m = vers_synth_unit(center_seed, ctr * (uint64_t)dim + c);
#if defined(__CUDA_ARCH__)
        return __fadd_rn(m, __fmul_rn(0.25f, u));
#else
        float t = 0.25f * u; /* exact */
        return m + t;
#endif
    }
    return u;
}

/* k-means initial centroid = a data row, drawn with replacement (ivfflat.rs:18-27) */
VERS_HD uint64_t vers_synth_init_row(uint64_t seed, uint32_t attempt, uint32_t j, uint64_t n_rows) {
    return vers_splitmix64(vers_splitmix64(seed) + ((uint64_t)attempt << 32) + j) % n_rows;
}

/* LSH: node path hash. root = vers_lsh_root_hash(tree); child = vers_lsh_child_hash(parent, above) */
VERS_HD uint64_t vers_lsh_root_hash(uint64_t seed, uint32_t tree) {
    return vers_splitmix64(vers_splitmix64(seed) ^ (0xA24BAED4963EE407ull * (uint64_t)(tree + 1)));
}
VERS_HD uint64_t vers_lsh_child_hash(uint64_t parent, uint32_t above) {
    return vers_splitmix64(parent * 2ull + (uint64_t)(above ? 1u : 0u) + 0x632BE59BD9B4E019ull);
}
/* two distinct positions in [0, len), len >= 2 (stands in for choose_multiple(rng, 2), lsh.rs:63-67) */
VERS_HD void vers_lsh_pick_pair(uint64_t node_hash, uint64_t len, uint64_t* pos_a, uint64_t* pos_b) {
    uint64_t h1 = vers_splitmix64(node_hash ^ 0x1234567ull);
    uint64_t h2 = vers_splitmix64(node_hash ^ 0x89ABCDEFull);
    uint64_t a = h1 % len;
    uint64_t b = h2 % (len - 1);
    if (b >= a) b += 1;
    *pos_a = a;
    *pos_b = b;
}

#endif /* VERS_SYNTH_H */
