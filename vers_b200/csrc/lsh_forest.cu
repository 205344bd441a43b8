// lsh_forest.cu — ANNIndex (indexes/lsh.rs:47-283): the random-hyperplane forest, built and searched on the GPU.
//
// build_index (lsh.rs:132-161)  dedup by bit pattern on the device (64-bit row hash + stable sort + bit compare inside
//   equal-hash runs, lsh.rs:113-130), then a LEVEL-SYNCHRONOUS
//   build of all trees at once: every node still to be split gets its plane from two sampled members
//   (vers_lsh_pick_pair stands in for choose_multiple(thread_rng, 2), lsh.rs:63-65), every member of every such node
//   is hashed in one launch with the exact-order dot engine (Hyperplane::point_is_above, lsh.rs:27-29), and a stable
//   partition (prefix sum of the hash bits) puts `below` before `above` keeping member order (lsh.rs:85-91).
//   Nodes with fewer than max_size members become leaves (lsh.rs:97-98) and move to fixed-capacity leaf slots.
// search_approximate (lsh.rs:264-282)  one warp per (query, tree) replays tree_result (lsh.rs:163-216) with an
//   explicit stack, including the reference's quirk that a backtracking node returns only the backup side's count;
//   hash bits and leaf distances use the reference's left-to-right fp32 arithmetic; then one warp per query unions the
//   trees' candidates (the DashSet), re-computes exact distances and takes the top-k by (distance, row index).
// add (lsh.rs:255-263, insert :218-251)  descend every tree on the device, append to the leaf or split it with the
//   same level-synchronous builder.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <unordered_map>

#include "scan.cuh"

namespace vers {
int32_t upload_queries(vers_ctx* ctx, const float* q, uint32_t nq, uint32_t stride, uint32_t dim, uint32_t ld,
                       float** d_q);

struct HostTree {
    std::vector<uint8_t> kind;  // 0 inner, 1 leaf
    std::vector<uint32_t> left, right, plane, slot, leaf_len;
    std::vector<uint64_t> hash;
    uint32_t add_node() {
        kind.push_back(1);
        left.push_back(0);
        right.push_back(0);
        plane.push_back(0);
        slot.push_back(0);
        leaf_len.push_back(0);
        hash.push_back(0);
        return (uint32_t)kind.size() - 1;
    }
};
}  // namespace vers

struct vers_lsh {
    vers_ctx* ctx = nullptr;
    uint32_t dim = 0, ld = 0, max_size = 0, num_trees = 0, slot_cap = 0;
    uint64_t seed = 0;
    uint64_t n = 0, cap = 0;        // deduplicated rows / capacity
    float* d_values = nullptr;      // [cap][ld]
    std::vector<uint64_t> ids;      // caller ids of the deduplicated rows
    std::vector<vers::HostTree> trees;
    float* d_planes = nullptr;      // [cap_planes][ld]
    float* d_consts = nullptr;
    uint32_t n_planes = 0, cap_planes = 0;
    uint32_t* d_leaf_items = nullptr;  // [cap_slots][slot_cap]
    uint32_t n_slots = 0, cap_slots = 0;
    // device mirror of the node arrays, all trees concatenated
    uint8_t* d_kind = nullptr;
    uint32_t *d_left = nullptr, *d_right = nullptr, *d_plane = nullptr, *d_slot = nullptr, *d_leaf_len = nullptr;
    uint32_t* d_tree_base = nullptr;  // [T]
    uint64_t* d_ids = nullptr;
    uint64_t nodes_cap = 0, ids_cap = 0;
    bool nodes_dirty = true, ids_dirty = true;
};

namespace vers {

constexpr int LSH_STACK = 160;      // frames per (query, tree) traversal
constexpr uint32_t HASH_CHUNK = 256;  // members per hash work item == NarrowCfg::TA

// ---------------------------------------------------------------- build kernels
struct SplitNode {
    uint64_t start;  // into the member buffer
    uint32_t len, pa, pb, plane;
};

// lsh.rs:71-73: coef = v[b] - v[a]; mid = (v[a] + v[b]) / 2; constant = -dot(coef, mid) (left to right, no FMA)
__global__ void __launch_bounds__(128) make_planes_kernel(const float* __restrict__ values, uint32_t dim, uint32_t ld,
                                                         const uint32_t* __restrict__ members,
                                                         const SplitNode* __restrict__ nodes, float* planes,
                                                         float* consts) {
    const SplitNode nd = nodes[blockIdx.x];
    const float* va = values + (uint64_t)members[nd.start + nd.pa] * ld;
    const float* vb = values + (uint64_t)members[nd.start + nd.pb] * ld;
    float* coef = planes + (uint64_t)nd.plane * ld;
    for (uint32_t i = threadIdx.x; i < ld; i += blockDim.x) coef[i] = i < dim ? __fsub_rn(vb[i], va[i]) : 0.0f;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (uint32_t i = 0; i < dim; ++i) {
            float m = __fdiv_rn(__fadd_rn(va[i], vb[i]), 2.0f);
            s = __fadd_rn(s, __fmul_rn(coef[i], m));
        }
        consts[nd.plane] = -s;
    }
}

struct HashItem {
    uint64_t start;  // first member of this chunk
    uint32_t count, plane;
};

// hash bit of every member of every node being split (exact-order dot engine, one plane per item)
__global__ void __launch_bounds__(NarrowCfg::NT, 2)
    hash_members_kernel(const float* __restrict__ values, uint32_t ld, const uint32_t* __restrict__ members,
                        const HashItem* __restrict__ items, const float* __restrict__ planes,
                        const float* __restrict__ consts, uint8_t* __restrict__ bits) {
    using Cfg = NarrowCfg;
    extern __shared__ __align__(16) float smem[];
    const HashItem it = items[blockIdx.x];
    RowSrc A{values, members + it.start, ld, it.count};
    RowSrc B{planes + (uint64_t)it.plane * ld, nullptr, ld, 1};
    float acc[Cfg::MA][Cfg::MB];
    tile_compute<Cfg, OP_DOT>(acc, A, 0, B, 0, ld, smem);
    const int ta = threadIdx.x % Cfg::NTA, tb = threadIdx.x / Cfg::NTA;
    if (tb == 0) {  // column 0 = tb 0, j 0
        const float k = consts[it.plane];
#pragma unroll
        for (int i = 0; i < Cfg::MA; ++i) {
            uint32_t r = ta + i * Cfg::NTA;
            if (r < it.count) bits[it.start + r] = (__fadd_rn(acc[i][0], k) >= 0.0f) ? 1 : 0;
        }
    }
}

__global__ void count_above_kernel(const uint32_t* __restrict__ S, const SplitNode* __restrict__ nodes, uint32_t n_nodes,
                                   uint32_t* __restrict__ n_above) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_nodes) n_above[j] = S[nodes[j].start + nodes[j].len] - S[nodes[j].start];
}

struct ScatterItem {
    uint64_t start;       // first member of this chunk
    uint64_t node_start;  // first member of the node
    uint32_t count, node_len;
};

// stable partition inside each node: [below | above], member order preserved on both sides (lsh.rs:82-91)
__global__ void scatter_members_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                       const uint8_t* __restrict__ bits, const uint32_t* __restrict__ S,
                                       const ScatterItem* __restrict__ items) {
    const ScatterItem it = items[blockIdx.x];
    const uint32_t node_above = S[it.node_start + it.node_len] - S[it.node_start];
    for (uint32_t r = threadIdx.x; r < it.count; r += blockDim.x) {
        uint64_t i = it.start + r;
        uint32_t rank_above = S[i] - S[it.node_start];
        uint64_t local = i - it.node_start;
        uint64_t dst = bits[i] ? it.node_start + (it.node_len - node_above) + rank_above
                               : it.node_start + (local - rank_above);
        out[dst] = in[i];
    }
}

struct LeafItem {
    uint64_t start;
    uint32_t len, slot;
};
__global__ void emit_leaves_kernel(const uint32_t* __restrict__ members, const LeafItem* __restrict__ items,
                                   uint32_t slot_cap, uint32_t* __restrict__ leaf_items) {
    const LeafItem it = items[blockIdx.x];
    for (uint32_t r = threadIdx.x; r < it.len; r += blockDim.x)
        leaf_items[(uint64_t)it.slot * slot_cap + r] = members[it.start + r];
}

// ---- deduplicate (lsh.rs:113-130) on the device
// 64-bit hash of a row's bit pattern (one warp per row: per-lane chains over its columns, combined with the lane index)
__global__ void dedup_hash_kernel(const float* __restrict__ rows, uint64_t n, uint32_t dim, uint32_t ld,
                                  unsigned long long* __restrict__ hash, uint32_t* __restrict__ row_id) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n; r += warps) {
        const uint32_t* x = reinterpret_cast<const uint32_t*>(rows + r * ld);
        uint64_t h = 0x243F6A8885A308D3ull + (uint64_t)lane;
        for (uint32_t c = lane; c < dim; c += 32) h = vers_splitmix64(h ^ x[c]);
        for (int o = 16; o; o >>= 1) h = vers_splitmix64(h + 31 * __shfl_xor_sync(FULL_MASK, h, o));
        if (lane == 0) {
            hash[r] = h;
            row_id[r] = (uint32_t)r;
        }
    }
}
// sorted position p: keep[row] = 1 unless an earlier row of the same hash run has identical bits (one warp per position)
__global__ void dedup_flag_kernel(const float* __restrict__ rows, uint64_t n, uint32_t dim, uint32_t ld,
                                  const unsigned long long* __restrict__ hs, const uint32_t* __restrict__ rs,
                                  uint32_t* __restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t p = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); p < n; p += warps) {
        const uint32_t row = rs[p];
        const uint32_t* x = reinterpret_cast<const uint32_t*>(rows + (uint64_t)row * ld);
        bool dup = false;
        for (uint64_t q = p; q > 0 && hs[q - 1] == hs[p] && !dup; --q) {
            const uint32_t* y = reinterpret_cast<const uint32_t*>(rows + (uint64_t)rs[q - 1] * ld);
            bool same = true;
            for (uint32_t c = lane; c < dim; c += 32) same = same && x[c] == y[c];
            dup = __all_sync(FULL_MASK, same);  // the stable sort put the earlier rows of the run first
        }
        if (lane == 0) keep[row] = dup ? 0u : 1u;
    }
}
__global__ void dedup_gather_kernel(const float* __restrict__ rows, uint64_t n, uint32_t ld,
                                    const uint32_t* __restrict__ keep, const uint32_t* __restrict__ at,
                                    float* __restrict__ out) {
    const uint32_t ld4 = ld >> 2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * ld4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / ld4;
        const uint32_t c = (uint32_t)(i - r * ld4);
        if (keep[r]) reinterpret_cast<float4*>(out)[(uint64_t)at[r] * ld4 + c] = reinterpret_cast<const float4*>(rows)[i];
    }
}

__global__ void iota_trees_kernel(uint32_t* p, uint64_t n, uint32_t trees) {
    uint64_t total = n * trees;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = (uint32_t)(i % n);
}

// ---------------------------------------------------------------- search kernels
struct ForestDev {
    const uint8_t* kind;
    const uint32_t *left, *right, *plane, *slot, *leaf_len, *tree_base;
    const float *planes, *consts, *values;
    const uint32_t* leaf_items;
    uint32_t ld, dim, slot_cap, num_trees;
};

// The reference's left-to-right dot (lsh.rs:27-29, base.rs:91-93).  Only the ADD chain is sequential by contract: the
// products coef[i] * q[i] are formed by the whole warp (coalesced 16-byte loads of the plane, separately rounded
// multiplies) into shared memory `prod` [ld], then lane 0 folds them in order.  With prod == null lane 0 does everything
// itself (one thread fetching the plane float by float: 25x slower on a plane that is not cached).
__device__ __forceinline__ bool plane_bit(const float* __restrict__ coef, float constant, const float* qs, uint32_t dim,
                                          int lane, float* prod = nullptr, uint32_t ld = 0) {
    int bit = 0;
    if (prod) {
        const float4* c4 = reinterpret_cast<const float4*>(coef);
        const float4* q4 = reinterpret_cast<const float4*>(qs);
        for (uint32_t i = lane; i < (ld >> 2); i += 32) {
            const float4 c = __ldg(c4 + i), q = q4[i];
            reinterpret_cast<float4*>(prod)[i] =
                make_float4(__fmul_rn(c.x, q.x), __fmul_rn(c.y, q.y), __fmul_rn(c.z, q.z), __fmul_rn(c.w, q.w));
        }
        __syncwarp();
        if (lane == 0) {
            float s = 0.0f;
            uint32_t i = 0;
            for (; i + 4 <= dim; i += 4) {
                const float4 p = *reinterpret_cast<const float4*>(prod + i);
                s = __fadd_rn(s, p.x);
                s = __fadd_rn(s, p.y);
                s = __fadd_rn(s, p.z);
                s = __fadd_rn(s, p.w);
            }
            for (; i < dim; ++i) s = __fadd_rn(s, prod[i]);
            bit = __fadd_rn(s, constant) >= 0.0f;
        }
        bit = __shfl_sync(FULL_MASK, bit, 0);
        __syncwarp();  // prod is rewritten at the next level
        return bit != 0;
    }
    if (lane == 0) {
        float s = 0.0f;
        for (uint32_t i = 0; i < dim; ++i) s = __fadd_rn(s, __fmul_rn(__ldg(coef + i), qs[i]));
        bit = __fadd_rn(s, constant) >= 0.0f;
    }
    return __shfl_sync(FULL_MASK, bit, 0) != 0;
}

__device__ __forceinline__ float exact_l2sq_row(const float* __restrict__ row, const float* qs, uint32_t ld) {
    float s = 0.0f;
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (uint32_t i = 0; i < ld; i += 4) {
        float4 a = __ldg(r4 + (i >> 2));
        float t;
        t = __fsub_rn(a.x, qs[i]); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(a.y, qs[i + 1]); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(a.z, qs[i + 2]); s = __fadd_rn(s, __fmul_rn(t, t));
        t = __fsub_rn(a.w, qs[i + 3]); s = __fadd_rn(s, __fmul_rn(t, t));
    }
    return s;
}

// The same exact-order distances for (up to) 32 rows at once, lane = row: the rows are scattered over the table, so a
// lane walking its own row through global memory pays a dependent DRAM round trip every few elements.  Here the warp
// stages the rows chunk by chunk (64 dimensions; cp.async, 16 bytes per lane and copy, two rows per instruction, the next
// chunk in flight while the current one is consumed) into a padded shared-memory tile that lane i then walks
// sequentially — the summation order per row is unchanged.  stage: [2][32][FR_LDS] floats of this warp.
template <int FR_KCH>
struct FrCfg {
    static constexpr int LDS = FR_KCH + 4;
    static constexpr size_t STAGE_BYTES = (size_t)2 * 32 * LDS * 4;
};
template <int FR_KCH>
__device__ __forceinline__ float staged_l2sq_32(const float* __restrict__ table, uint32_t ld, uint32_t my_row, bool live,
                                                const float* qs, float* stage, int lane) {
    constexpr int FR_LDS = FrCfg<FR_KCH>::LDS, LPR = FR_KCH / 4, RPI = 32 / LPR;  // lanes per row chunk, rows per copy
    const int half = lane / LPR, sub = lane % LPR;
    const uint32_t nch = (ld + FR_KCH - 1) / FR_KCH;
    auto issue = [&](uint32_t c) {
        const uint32_t col = c * FR_KCH + sub * 4;
        float* tb = stage + (size_t)(c & 1u) * 32 * FR_LDS;
#pragma unroll
        for (int i = 0; i < 32 / RPI; ++i) {
            const uint32_t r = __shfl_sync(FULL_MASK, my_row, RPI * i + half);
            const bool ok = __shfl_sync(FULL_MASK, (int)live, RPI * i + half) != 0 && col < ld;
            cp_async16(tb + (RPI * i + half) * FR_LDS + sub * 4, ok ? (const void*)(table + (uint64_t)r * ld + col) : (const void*)table, ok);
        }
        cp_async_commit();
    };
    float s = 0.0f;
    issue(0);
    for (uint32_t c = 0; c < nch; ++c) {
        const uint32_t k0 = c * FR_KCH, kn = min((uint32_t)FR_KCH, ld - k0);
        if (c + 1 < nch) {
            issue(c + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float* tl = stage + ((size_t)(c & 1u) * 32 + lane) * FR_LDS;
        if (live) {
#pragma unroll 4
            for (uint32_t i = 0; i < kn; i += 4) {
                const float4 a = *reinterpret_cast<const float4*>(tl + i);
                const float4 b = *reinterpret_cast<const float4*>(qs + k0 + i);
                float t;
                t = __fsub_rn(a.x, b.x); s = __fadd_rn(s, __fmul_rn(t, t));
                t = __fsub_rn(a.y, b.y); s = __fadd_rn(s, __fmul_rn(t, t));
                t = __fsub_rn(a.z, b.z); s = __fadd_rn(s, __fmul_rn(t, t));
                t = __fsub_rn(a.w, b.w); s = __fadd_rn(s, __fmul_rn(t, t));
            }
        }
        __syncwarp();
    }
    return s;
}

// one warp per (query, tree): tree_result (lsh.rs:163-216) with an explicit stack
constexpr int FT_WPB = 2;  // warps per block: the warps are independent and bound by latency, small blocks pack the SM
template <int KCH>
__global__ void __launch_bounds__(FT_WPB * 32)
    forest_traverse_kernel(ForestDev f, const float* __restrict__ queries, uint32_t nq, uint32_t top_k,
                           uint32_t cand_cap, uint32_t* __restrict__ cand, uint32_t* __restrict__ cand_cnt,
                           uint32_t* overflow) {
    extern __shared__ __align__(16) unsigned char fsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t w = (uint64_t)blockIdx.x * FT_WPB + warp;
    if (w >= (uint64_t)nq * f.num_trees) return;
    const uint32_t q = (uint32_t)(w / f.num_trees), t = (uint32_t)(w % f.num_trees);
    // shared memory per warp: row staging of the leaf scans (between leaves: the plane products) | query [ld] | stack
    // nodes [LSH_STACK] | stack n | stack stage+backup | top-n list
    constexpr size_t FR_STAGE_BYTES = FrCfg<KCH>::STAGE_BYTES;
    const size_t per_warp = FR_STAGE_BYTES + (size_t)f.ld * 4 + LSH_STACK * 12 + (size_t)top_k * 8;
    unsigned char* base = fsm + warp * ((per_warp + 15) & ~size_t(15));
    float* stage = reinterpret_cast<float*>(base);
    float* qs = reinterpret_cast<float*>(base + FR_STAGE_BYTES);
    float* prod = (size_t)f.ld * 4 <= FR_STAGE_BYTES ? stage : nullptr;  // very wide rows: lane 0 walks the plane alone
    uint32_t* st_node = reinterpret_cast<uint32_t*>(base + FR_STAGE_BYTES + (size_t)f.ld * 4);
    int32_t* st_n = reinterpret_cast<int32_t*>(st_node + LSH_STACK);
    uint32_t* st_aux = reinterpret_cast<uint32_t*>(st_n + LSH_STACK);  // bit31 = waiting for backup, low bits = backup node
    float* ld_ = reinterpret_cast<float*>(st_aux + LSH_STACK);
    uint32_t* lp = reinterpret_cast<uint32_t*>(ld_ + top_k);
    for (uint32_t i = lane; i < f.ld; i += 32) qs[i] = queries[(uint64_t)q * f.ld + i];
    __syncwarp();
    const uint32_t nb = f.tree_base[t];
    uint32_t* out = cand + ((uint64_t)q * f.num_trees + t) * cand_cap;
    uint32_t n_out = 0;
    int sp = 0;
    int32_t ret = 0;
    if (lane == 0) {
        st_node[0] = 0;
        st_n[0] = (int32_t)top_k;
        st_aux[0] = 0xffffffffu;  // not yet visited
    }
    __syncwarp();
    sp = 1;
    bool over = false;
    while (sp > 0) {
        const uint32_t node = st_node[sp - 1];
        const int32_t n = st_n[sp - 1];
        const uint32_t aux = st_aux[sp - 1];
        const uint32_t g = nb + node;
        if (f.kind[g] == 1) {
            const uint32_t len = f.leaf_len[g];
            const uint32_t* items = f.leaf_items + (uint64_t)f.slot[g] * f.slot_cap;
            if ((int64_t)len < (int64_t)n) {  // take every member (lsh.rs:174-180)
                for (uint32_t i = lane; i < len; i += 32) {
                    if (n_out + i < cand_cap) out[n_out + i] = items[i]; else over = true;
                }
                n_out += len;
                ret = (int32_t)len;
            } else {  // the n closest members of the leaf, stable in leaf order (lsh.rs:185-198)
                for (int32_t e = lane; e < n; e += 32) {
                    ld_[e] = __int_as_float(0x7f800000);
                    lp[e] = 0xffffffffu;
                }
                __syncwarp();
                for (uint32_t i0 = 0; i0 < len; i0 += 32) {
                    uint32_t i = i0 + lane;
                    bool live = i < len;
                    const float d = staged_l2sq_32<KCH>(f.values, f.ld, live ? items[i] : 0u, live, qs, stage, lane);
                    while (n > 0) {
                        bool pass = live && entry_less<uint32_t>(d, i, ld_[n - 1], lp[n - 1]);
                        unsigned m = __ballot_sync(FULL_MASK, pass);
                        if (!m) break;
                        int src = __ffs(m) - 1;
                        float bd = __shfl_sync(FULL_MASK, d, src);
                        uint32_t bi = __shfl_sync(FULL_MASK, i, src);
                        warp_topk_insert<uint32_t>(ld_, lp, n, bd, bi, lane);
                        if (lane == src) live = false;
                    }
                }
                for (int32_t e = lane; e < n; e += 32) {
                    if (n_out + e < cand_cap) out[n_out + e] = items[lp[e]]; else over = true;
                }
                n_out += (uint32_t)n;
                ret = n;
            }
            --sp;
            __syncwarp();
            continue;
        }
        if (aux == 0xffffffffu) {  // first visit: hash, descend into the main side
            bool above = plane_bit(f.planes + (uint64_t)f.plane[g] * f.ld, f.consts[f.plane[g]], qs, f.dim, lane, prod, f.ld);
            uint32_t main_n = above ? f.right[g] : f.left[g];
            uint32_t back_n = above ? f.left[g] : f.right[g];
            if (sp >= LSH_STACK) {
                over = true;
                break;
            }
            if (lane == 0) {
                st_aux[sp - 1] = back_n & 0x7fffffffu;
                st_node[sp] = main_n;
                st_n[sp] = n;
                st_aux[sp] = 0xffffffffu;
            }
            __syncwarp();
            ++sp;
            continue;
        }
        if (!(aux & 0x80000000u)) {  // main side returned `ret`
            if (ret < n) {            // lsh.rs:210-211: search the backup side for n - k MORE (its count is returned)
                if (lane == 0) {
                    st_aux[sp - 1] = aux | 0x80000000u;
                    st_node[sp] = aux;
                    st_n[sp] = n - ret;
                    st_aux[sp] = 0xffffffffu;
                }
                __syncwarp();
                ++sp;
                continue;
            }
            --sp;  // ret stays k
            continue;
        }
        --sp;  // backup side returned: propagate ITS count (the reference drops k here)
    }
    if (lane == 0) {
        cand_cnt[(uint64_t)q * f.num_trees + t] = n_out < cand_cap ? n_out : cand_cap;
    }
    over = __any_sync(FULL_MASK, over) || n_out > cand_cap;
    if (over && lane == 0) atomicOr(overflow, 1u);
}

// one warp per query: union of the trees' candidates, exact distances, top-k by (distance, row index) -> ids[idx]
template <int KCH>
__global__ void __launch_bounds__(128)
    forest_rerank_kernel(ForestDev f, const float* __restrict__ queries, uint32_t nq, uint32_t top_k, uint32_t cand_cap,
                         const uint32_t* __restrict__ cand, const uint32_t* __restrict__ cand_cnt,
                         const uint64_t* __restrict__ ids, uint64_t* out_ids, float* out_d, uint32_t* out_cnt) {
    extern __shared__ __align__(16) unsigned char rsm3[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * 4 + warp;
    if (q >= nq) return;
    constexpr size_t FR_STAGE_BYTES = FrCfg<KCH>::STAGE_BYTES;
    const size_t per_warp = FR_STAGE_BYTES + (size_t)f.ld * 4 + (size_t)top_k * 8;
    unsigned char* base = rsm3 + warp * ((per_warp + 15) & ~size_t(15));
    float* stage = reinterpret_cast<float*>(base);
    float* qs = reinterpret_cast<float*>(base + FR_STAGE_BYTES);
    float* sd = reinterpret_cast<float*>(base + FR_STAGE_BYTES + (size_t)f.ld * 4);
    uint32_t* sp = reinterpret_cast<uint32_t*>(sd + top_k);
    for (uint32_t i = lane; i < f.ld; i += 32) qs[i] = queries[(uint64_t)q * f.ld + i];
    for (uint32_t e = lane; e < top_k; e += 32) {
        sd[e] = __int_as_float(0x7f800000);
        sp[e] = 0xffffffffu;
    }
    __syncwarp();
    const int k = (int)top_k;
    // the trees' candidate lists back to back: 32 candidates per round (a tree contributes ~top_k of them, so a round
    // per tree left two thirds of the lanes idle).  Order does not matter: top-k by the total order (distance, row),
    // duplicates dropped by row.
    uint32_t total = 0;
    for (uint32_t t = 0; t < f.num_trees; ++t) total += cand_cnt[(uint64_t)q * f.num_trees + t];
    {
        for (uint32_t i0 = 0; i0 < total; i0 += 32) {
            uint32_t i = i0 + lane;
            bool live = i < total;
            uint32_t idx = 0xffffffffu;
            if (live) {
                uint32_t rest = i, t = 0;
                for (;; ++t) {
                    const uint32_t c = cand_cnt[(uint64_t)q * f.num_trees + t];
                    if (rest < c) break;
                    rest -= c;
                }
                idx = cand[((uint64_t)q * f.num_trees + t) * cand_cap + rest];
            }
            const float d = staged_l2sq_32<KCH>(f.values, f.ld, live ? idx : 0u, live, qs, stage, lane);
            while (true) {
                bool pass = live && entry_less<uint32_t>(d, idx, sd[k - 1], sp[k - 1]);
                unsigned m = __ballot_sync(FULL_MASK, pass);
                if (!m) break;
                int src = __ffs(m) - 1;
                float bd = __shfl_sync(FULL_MASK, d, src);
                uint32_t bi = __shfl_sync(FULL_MASK, idx, src);
                // the DashSet: a row reached through several trees counts once
                bool dup = false;
                for (int e0 = 0; e0 < k; e0 += 32) {
                    int e = e0 + lane;
                    dup |= __any_sync(FULL_MASK, e < k && sp[e] == bi);
                }
                if (!dup) warp_topk_insert<uint32_t>(sd, sp, k, bd, bi, lane);
                if (lane == src) live = false;
            }
        }
    }
    uint32_t cntv = 0;
    for (uint32_t e0 = 0; e0 < top_k; e0 += 32) {
        uint32_t e = e0 + lane;
        bool have = false;
        if (e < top_k) {
            have = sp[e] != 0xffffffffu;
            out_ids[(uint64_t)q * top_k + e] = have ? ids[sp[e]] : 0xffffffffffffffffull;
            out_d[(uint64_t)q * top_k + e] = sd[e];
        }
        cntv += __popc(__ballot_sync(FULL_MASK, have));
    }
    if (out_cnt && lane == 0) out_cnt[q] = cntv;
}

// one warp per tree: the leaf a new row lands in (insert, lsh.rs:225-236)
__global__ void forest_descend_kernel(ForestDev f, const float* __restrict__ row, uint32_t* __restrict__ leaf_node) {
    extern __shared__ __align__(16) float dsm[];  // the row [ld], then the plane products [ld]
    const int lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x;
    for (uint32_t i = lane; i < f.ld; i += 32) dsm[i] = row[i];
    __syncwarp();
    const uint32_t nb = f.tree_base[t];
    uint32_t node = 0;
    while (f.kind[nb + node] == 0) {
        uint32_t g = nb + node;
        bool above = plane_bit(f.planes + (uint64_t)f.plane[g] * f.ld, f.consts[f.plane[g]], dsm, f.dim, lane, dsm + f.ld, f.ld);
        node = above ? f.right[g] : f.left[g];
    }
    if (lane == 0) leaf_node[t] = node;
}

// ---------------------------------------------------------------- host side
template <typename T>
static int32_t grow_device(T** p, uint64_t* cap, uint64_t need, uint64_t keep, cudaStream_t s) {
    if (need <= *cap) return VERS_OK;
    uint64_t ncap = std::max<uint64_t>(need, *cap * 2 + 64);
    T* np_ = nullptr;
    VERS_CUDA(cudaMalloc(&np_, ncap * sizeof(T)));
    if (*p && keep) VERS_CUDA(cudaMemcpyAsync(np_, *p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
    VERS_CUDA(cudaStreamSynchronize(s));
    if (*p) cudaFree(*p);
    *p = np_;
    *cap = ncap;
    return VERS_OK;
}

struct BuildRoot {
    uint32_t tree, node;
    uint64_t start;
    uint32_t len;
    uint64_t hash;
};

template <typename T>
static int32_t upload_vec(const std::vector<T>& v, T** d_buf, uint64_t* cap, cudaStream_t s) {
    VERS_TRY(grow_device(d_buf, cap, v.size() ? v.size() : 1, 0, s));
    if (!v.empty()) VERS_CUDA(cudaMemcpyAsync(*d_buf, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return VERS_OK;
}

// level-synchronous construction of the subtrees rooted at `roots`, whose members are the segments of d_mem_a
static int32_t build_subtrees(vers_lsh* L, std::vector<BuildRoot> frontier, uint32_t* d_mem_a, uint32_t* d_mem_b,
                              uint64_t total) {
    vers_ctx* ctx = L->ctx;
    cudaStream_t s = ctx->stream;
    uint8_t* d_bits = nullptr;
    uint32_t* d_S = nullptr;
    void* d_cub = nullptr;
    size_t cub_bytes = 0;
    SplitNode* d_nodes = nullptr;
    HashItem* d_hitems = nullptr;
    ScatterItem* d_sitems = nullptr;
    LeafItem* d_litems = nullptr;
    uint32_t* d_above = nullptr;
    uint64_t c_nodes = 0, c_h = 0, c_s = 0, c_l = 0, c_above = 0;
    int32_t rc = VERS_OK;
    auto cleanup = [&]() {
        cudaFree(d_bits);
        cudaFree(d_S);
        cudaFree(d_cub);
        cudaFree(d_nodes);
        cudaFree(d_hitems);
        cudaFree(d_sitems);
        cudaFree(d_litems);
        cudaFree(d_above);
    };
#define LB_TRY(x)            \
    do {                     \
        rc = (x);            \
        if (rc != VERS_OK) { \
            cleanup();       \
            return rc;       \
        }                    \
    } while (0)
#define LB_CUDA(x)                                                                                    \
    do {                                                                                              \
        cudaError_t _e = (x);                                                                         \
        if (_e != cudaSuccess) {                                                                      \
            cleanup();                                                                                \
            return fail(VERS_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                             \
    } while (0)
    LB_CUDA(cudaMalloc(&d_bits, total + 1));
    LB_CUDA(cudaMalloc(&d_S, (total + 1) * 4));
    LB_CUDA(cudaMemsetAsync(d_bits, 0, total + 1, s));
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, d_bits, d_S, (int64_t)(total + 1), s);
    LB_CUDA(cudaMalloc(&d_cub, cub_bytes ? cub_bytes : 1));
    uint32_t* cur = d_mem_a;
    uint32_t* nxt = d_mem_b;
    uint32_t depth = 0;
    while (!frontier.empty()) {
        if (++depth > 4096) {
            cleanup();
            return fail(VERS_ERR_PANIC, "lsh build: recursion deeper than 4096 (the reference would overflow its stack)");
        }
        std::vector<SplitNode> split;
        std::vector<BuildRoot> split_roots;
        std::vector<LeafItem> leaves;
        for (const BuildRoot& r : frontier) {
            HostTree& T = L->trees[r.tree];
            T.hash[r.node] = r.hash;
            if (r.len < L->max_size) {  // lsh.rs:97-98
                T.kind[r.node] = 1;
                T.slot[r.node] = L->n_slots++;
                T.leaf_len[r.node] = r.len;
                leaves.push_back(LeafItem{r.start, r.len, T.slot[r.node]});
            } else {
                if (r.len < 2) {
                    cleanup();
                    return fail(VERS_ERR_PANIC, "lsh build: a node of %u rows cannot be split (samples[1] out of bounds, lsh.rs:67)", r.len);
                }
                uint64_t pa, pb;
                vers_lsh_pick_pair(r.hash, r.len, &pa, &pb);
                T.kind[r.node] = 0;
                T.plane[r.node] = L->n_planes;
                split.push_back(SplitNode{r.start, r.len, (uint32_t)pa, (uint32_t)pb, L->n_planes});
                L->n_planes++;
                split_roots.push_back(r);
            }
        }
        if (!leaves.empty()) {
            uint64_t cs = L->cap_slots;
            uint64_t need = (uint64_t)L->n_slots * L->slot_cap, have = (uint64_t)L->cap_slots * L->slot_cap;
            (void)cs;
            if (need > have) {
                uint64_t capw = have;
                LB_TRY(grow_device(&L->d_leaf_items, &capw, need, have, s));
                L->cap_slots = (uint32_t)(capw / L->slot_cap);
            }
            LB_TRY(upload_vec(leaves, &d_litems, &c_l, s));
            emit_leaves_kernel<<<(unsigned)leaves.size(), 128, 0, s>>>(cur, d_litems, L->slot_cap, L->d_leaf_items);
            ctx->launches += 1;
            LB_CUDA(cudaGetLastError());
        }
        if (split.empty()) break;
        {  // plane pool capacity
            uint64_t capp = (uint64_t)L->cap_planes * L->ld, capc = L->cap_planes;
            uint64_t keepn = (uint64_t)(L->n_planes - split.size());
            if (L->n_planes > L->cap_planes) {
                uint64_t want = std::max<uint64_t>(L->n_planes, (uint64_t)L->cap_planes * 2 + 1024);
                LB_TRY(grow_device(&L->d_planes, &capp, want * L->ld, keepn * L->ld, s));
                LB_TRY(grow_device(&L->d_consts, &capc, want, keepn, s));
                L->cap_planes = (uint32_t)std::min<uint64_t>(capp / L->ld, capc);
            }
        }
        std::vector<HashItem> hitems;
        std::vector<ScatterItem> sitems;
        for (const SplitNode& nd : split) {
            for (uint32_t off = 0; off < nd.len; off += HASH_CHUNK) {
                uint32_t cnt = std::min<uint32_t>(HASH_CHUNK, nd.len - off);
                hitems.push_back(HashItem{nd.start + off, cnt, nd.plane});
                sitems.push_back(ScatterItem{nd.start + off, nd.start, cnt, nd.len});
            }
        }
        LB_TRY(upload_vec(split, &d_nodes, &c_nodes, s));
        LB_TRY(upload_vec(hitems, &d_hitems, &c_h, s));
        LB_TRY(upload_vec(sitems, &d_sitems, &c_s, s));
        LB_TRY(grow_device(&d_above, &c_above, split.size(), 0, s));
        make_planes_kernel<<<(unsigned)split.size(), 128, 0, s>>>(L->d_values, L->dim, L->ld, cur, d_nodes, L->d_planes,
                                                                 L->d_consts);
        ctx->launches += 1;
        LB_CUDA(cudaGetLastError());
        LB_CUDA(cudaMemsetAsync(d_bits, 0, total + 1, s));
        {
            auto kern = hash_members_kernel;
            size_t smem = (size_t)NarrowCfg::TILE_FLOATS * 4;
            LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FamilyTimer ft(ctx, KF_LSH_HASH);
            kern<<<(unsigned)hitems.size(), NarrowCfg::NT, smem, s>>>(L->d_values, L->ld, cur, d_hitems, L->d_planes,
                                                                     L->d_consts, d_bits);
            ctx->launches += 1;
            LB_CUDA(cudaGetLastError());
        }
        size_t cb = cub_bytes;
        LB_CUDA(cub::DeviceScan::ExclusiveSum(d_cub, cb, d_bits, d_S, (int64_t)(total + 1), s));
        ctx->launches += 1;
        count_above_kernel<<<(unsigned)ceil_div(split.size(), 256), 256, 0, s>>>(d_S, d_nodes, (uint32_t)split.size(),
                                                                                d_above);
        scatter_members_kernel<<<(unsigned)sitems.size(), 256, 0, s>>>(cur, nxt, d_bits, d_S, d_sitems);
        ctx->launches += 2;
        LB_CUDA(cudaGetLastError());
        std::vector<uint32_t> above(split.size());
        LB_CUDA(cudaMemcpyAsync(above.data(), d_above, above.size() * 4, cudaMemcpyDeviceToHost, s));
        LB_CUDA(cudaStreamSynchronize(s));
        std::vector<BuildRoot> next;
        next.reserve(split.size() * 2);
        for (size_t j = 0; j < split.size(); ++j) {
            const BuildRoot& r = split_roots[j];
            HostTree& T = L->trees[r.tree];
            uint32_t na = above[j], nbel = r.len - na;
            if (na == 0 || nbel == 0) {
                cleanup();
                return fail(VERS_ERR_PANIC, "lsh build: degenerate split (all %u rows on one side); the reference would "
                                            "recurse forever", r.len);
            }
            uint32_t nr = T.add_node();  // above -> right_node (lsh.rs:108), created first like the oracle
            uint32_t nl = T.add_node();  // below -> left_node  (lsh.rs:107)
            T.right[r.node] = nr;
            T.left[r.node] = nl;
            next.push_back(BuildRoot{r.tree, nr, r.start + nbel, na, vers_lsh_child_hash(r.hash, 1)});
            next.push_back(BuildRoot{r.tree, nl, r.start, nbel, vers_lsh_child_hash(r.hash, 0)});
        }
        frontier.swap(next);
        std::swap(cur, nxt);
    }
    cleanup();
    L->nodes_dirty = true;
    return VERS_OK;
#undef LB_TRY
#undef LB_CUDA
}

static int32_t sync_device_mirror(vers_lsh* L) {
    vers_ctx* ctx = L->ctx;
    cudaStream_t s = ctx->stream;
    if (L->nodes_dirty) {
        uint64_t total = 0;
        std::vector<uint32_t> base(L->num_trees);
        for (uint32_t t = 0; t < L->num_trees; ++t) {
            base[t] = (uint32_t)total;
            total += L->trees[t].kind.size();
        }
        if (total > L->nodes_cap) {
            cudaFree(L->d_kind);
            cudaFree(L->d_left);
            cudaFree(L->d_right);
            cudaFree(L->d_plane);
            cudaFree(L->d_slot);
            cudaFree(L->d_leaf_len);
            uint64_t c = total + total / 2 + 64;
            VERS_CUDA(cudaMalloc(&L->d_kind, c));
            VERS_CUDA(cudaMalloc(&L->d_left, c * 4));
            VERS_CUDA(cudaMalloc(&L->d_right, c * 4));
            VERS_CUDA(cudaMalloc(&L->d_plane, c * 4));
            VERS_CUDA(cudaMalloc(&L->d_slot, c * 4));
            VERS_CUDA(cudaMalloc(&L->d_leaf_len, c * 4));
            L->nodes_cap = c;
        }
        if (!L->d_tree_base) VERS_CUDA(cudaMalloc(&L->d_tree_base, (size_t)std::max(1u, L->num_trees) * 4));
        for (uint32_t t = 0; t < L->num_trees; ++t) {
            const HostTree& T = L->trees[t];
            size_t nn = T.kind.size();
            VERS_CUDA(cudaMemcpyAsync(L->d_kind + base[t], T.kind.data(), nn, cudaMemcpyHostToDevice, s));
            VERS_CUDA(cudaMemcpyAsync(L->d_left + base[t], T.left.data(), nn * 4, cudaMemcpyHostToDevice, s));
            VERS_CUDA(cudaMemcpyAsync(L->d_right + base[t], T.right.data(), nn * 4, cudaMemcpyHostToDevice, s));
            VERS_CUDA(cudaMemcpyAsync(L->d_plane + base[t], T.plane.data(), nn * 4, cudaMemcpyHostToDevice, s));
            VERS_CUDA(cudaMemcpyAsync(L->d_slot + base[t], T.slot.data(), nn * 4, cudaMemcpyHostToDevice, s));
            VERS_CUDA(cudaMemcpyAsync(L->d_leaf_len + base[t], T.leaf_len.data(), nn * 4, cudaMemcpyHostToDevice, s));
        }
        VERS_CUDA(cudaMemcpyAsync(L->d_tree_base, base.data(), (size_t)L->num_trees * 4, cudaMemcpyHostToDevice, s));
        VERS_CUDA(cudaStreamSynchronize(s));
        L->nodes_dirty = false;
    }
    if (L->ids_dirty) {
        if (L->ids.size() > L->ids_cap) {
            cudaFree(L->d_ids);
            uint64_t c = L->ids.size() + L->ids.size() / 2 + 64;
            VERS_CUDA(cudaMalloc(&L->d_ids, c * 8));
            L->ids_cap = c;
        }
        if (!L->ids.empty())
            VERS_CUDA(cudaMemcpyAsync(L->d_ids, L->ids.data(), L->ids.size() * 8, cudaMemcpyHostToDevice, s));
        VERS_CUDA(cudaStreamSynchronize(s));
        L->ids_dirty = false;
    }
    return VERS_OK;
}

static ForestDev forest_dev(const vers_lsh* L) {
    ForestDev f;
    f.kind = L->d_kind;
    f.left = L->d_left;
    f.right = L->d_right;
    f.plane = L->d_plane;
    f.slot = L->d_slot;
    f.leaf_len = L->d_leaf_len;
    f.tree_base = L->d_tree_base;
    f.planes = L->d_planes;
    f.consts = L->d_consts;
    f.values = L->d_values;
    f.leaf_items = L->d_leaf_items;
    f.ld = L->ld;
    f.dim = L->dim;
    f.slot_cap = L->slot_cap;
    f.num_trees = L->num_trees;
    return f;
}

}  // namespace vers

using namespace vers;

extern "C" int32_t vers_lsh_free(vers_lsh* L) {
    if (!L) return VERS_OK;
    cudaSetDevice(L->ctx->device);
    cudaStreamSynchronize(L->ctx->stream);
    cudaFree(L->d_values);
    cudaFree(L->d_planes);
    cudaFree(L->d_consts);
    cudaFree(L->d_leaf_items);
    cudaFree(L->d_kind);
    cudaFree(L->d_left);
    cudaFree(L->d_right);
    cudaFree(L->d_plane);
    cudaFree(L->d_slot);
    cudaFree(L->d_leaf_len);
    cudaFree(L->d_tree_base);
    cudaFree(L->d_ids);
    delete L;
    return VERS_OK;
}

extern "C" int32_t vers_lsh_build_index(vers_ctx* ctx, const float* rows, uint64_t n, uint32_t dim,
                                        uint32_t stride_floats, const uint64_t* vector_ids, uint32_t num_trees,
                                        uint32_t max_size, uint64_t seed, vers_lsh** out) {
    if (!ctx || !out || (!rows && n)) return fail(VERS_ERR_ARG, "lsh_build_index: null argument");
    *out = nullptr;
    if (dim == 0 || stride_floats < dim) return fail(VERS_ERR_ARG, "lsh_build_index: bad dim/stride");
    if (n >= 0xffffffffull) return fail(VERS_ERR_UNSUPPORTED, "lsh_build_index: more than 2^32-2 rows");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    vers_lsh* L = new vers_lsh();
    L->ctx = ctx;
    L->dim = dim;
    L->ld = round_up(dim, 4);
    L->max_size = max_size;
    L->num_trees = num_trees;
    L->slot_cap = max_size + 1;
    L->seed = seed;
    // deduplicate (lsh.rs:113-130) ON THE DEVICE: keep the first row of every distinct bit pattern, in row order.
    // 64-bit hash of every row's bit pattern -> stable radix sort of (hash, row) -> a row is a duplicate iff an
    // EARLIER row of its equal-hash run has the same bits (runs are one row long unless rows repeat or hashes collide;
    // the backward walk stops at the first bit-identical row) -> exclusive scan of the keep flags -> gather.
    int32_t rc = VERS_OK;
    uint32_t* d_mem_a = nullptr;
    uint32_t* d_mem_b = nullptr;
    auto bail = [&](int32_t code) {
        cudaFree(d_mem_a);
        cudaFree(d_mem_b);
        vers_lsh_free(L);
        return code;
    };
    {
        cudaStream_t s = ctx->stream;
        float* d_all = nullptr;
        unsigned long long *d_h = nullptr, *d_hs = nullptr;
        uint32_t *d_r = nullptr, *d_rs = nullptr, *d_keep = nullptr, *d_at = nullptr;
        void* d_cub = nullptr;
        std::vector<uint32_t> keep(n);
        auto drop = [&]() {
            cudaFree(d_all), cudaFree(d_h), cudaFree(d_hs), cudaFree(d_r), cudaFree(d_rs), cudaFree(d_keep), cudaFree(d_at);
            cudaFree(d_cub);
        };
        const size_t n1 = n ? n : 1;
        cudaError_t e = cudaMalloc(&d_all, n1 * L->ld * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d_h, n1 * 8);
        if (e == cudaSuccess) e = cudaMalloc(&d_hs, n1 * 8);
        if (e == cudaSuccess) e = cudaMalloc(&d_r, n1 * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d_rs, n1 * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d_keep, (n1 + 1) * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d_at, (n1 + 1) * 4);
        uint64_t kept = 0;
        if (e == cudaSuccess && n) {
            if (L->ld != dim) e = cudaMemsetAsync(d_all, 0, n * L->ld * 4, s);
            if (e == cudaSuccess)
                e = cudaMemcpy2DAsync(d_all, (size_t)L->ld * 4, rows, (size_t)stride_floats * 4, (size_t)dim * 4, n,
                                      cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess) {
                dedup_hash_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(d_all, n, dim, L->ld, d_h, d_r);
                size_t need = 0, need2 = 0;
                cub::DeviceRadixSort::SortPairs(nullptr, need, d_h, d_hs, d_r, d_rs, (int64_t)n, 0, 64, s);
                cub::DeviceScan::ExclusiveSum(nullptr, need2, d_keep, d_at, (int64_t)(n + 1), s);
                e = cudaMalloc(&d_cub, std::max(need, need2));
                if (e == cudaSuccess) {
                    size_t nb = std::max(need, need2);
                    e = cub::DeviceRadixSort::SortPairs(d_cub, nb, d_h, d_hs, d_r, d_rs, (int64_t)n, 0, 64, s);
                    if (e == cudaSuccess) e = cudaMemsetAsync(d_keep, 0, (n + 1) * 4, s);
                    if (e == cudaSuccess) {
                        dedup_flag_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(d_all, n, dim, L->ld, d_hs, d_rs, d_keep);
                        nb = std::max(need, need2);
                        e = cub::DeviceScan::ExclusiveSum(d_cub, nb, d_keep, d_at, (int64_t)(n + 1), s);
                    }
                    ctx->launches += 4;
                }
            }
            uint32_t total_kept = 0;
            if (e == cudaSuccess) e = cudaMemcpyAsync(keep.data(), d_keep, n * 4, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(&total_kept, d_at + n, 4, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            kept = total_kept;
        }
        if (e == cudaSuccess) {
            L->n = kept;
            L->cap = L->n + L->n / 8 + 64;
            e = cudaMalloc(&L->d_values, (size_t)L->cap * L->ld * 4);
            if (e == cudaSuccess && n) {
                dedup_gather_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(d_all, n, L->ld, d_keep, d_at, L->d_values);
                ctx->launches += 1;
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            }
        }
        drop();
        if (e != cudaSuccess)
            return bail(fail(e == cudaErrorMemoryAllocation ? VERS_ERR_NOMEM : VERS_ERR_CUDA, "lsh_build_index: %s",
                             cudaGetErrorString(e)));
        L->ids.reserve(kept);
        for (uint64_t r = 0; r < n; ++r)
            if (keep[r]) L->ids.push_back(vector_ids ? vector_ids[r] : r);
    }
    L->trees.resize(num_trees);
    const uint64_t total = (uint64_t)L->n * num_trees;
    std::vector<BuildRoot> roots;
    for (uint32_t t = 0; t < num_trees; ++t) {
        uint32_t root = L->trees[t].add_node();
        roots.push_back(BuildRoot{t, root, (uint64_t)t * L->n, (uint32_t)L->n, vers_lsh_root_hash(seed, t)});
    }
    if (cudaMalloc(&d_mem_a, (total ? total : 1) * 4) != cudaSuccess ||
        cudaMalloc(&d_mem_b, (total ? total : 1) * 4) != cudaSuccess)
        return bail(fail(VERS_ERR_NOMEM, "lsh_build_index: cudaMalloc members"));
    if (total) {
        iota_trees_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_mem_a, L->n, num_trees);
        ctx->launches += 1;
    }
    rc = build_subtrees(L, roots, d_mem_a, d_mem_b, total);
    if (rc != VERS_OK) return bail(rc);
    cudaFree(d_mem_a);
    cudaFree(d_mem_b);
    *out = L;
    return VERS_OK;
}

extern "C" int32_t vers_lsh_info(const vers_lsh* L, uint64_t* num_values, uint32_t* num_trees, uint64_t* num_nodes) {
    if (!L) return fail(VERS_ERR_ARG, "lsh_info: null");
    if (num_values) *num_values = L->n;
    if (num_trees) *num_trees = L->num_trees;
    if (num_nodes) {
        uint64_t t = 0;
        for (const HostTree& T : L->trees) t += T.kind.size();
        *num_nodes = t;
    }
    return VERS_OK;
}

extern "C" int32_t vers_lsh_flatten(const vers_lsh* L, uint32_t tree, uint8_t* kind, uint32_t* leaf_len, float* planes,
                                    float* consts, uint32_t* items, uint32_t* n_nodes, uint32_t* n_inner,
                                    uint64_t* n_items) {
    if (!L || tree >= L->num_trees || !n_nodes || !n_inner || !n_items) return fail(VERS_ERR_ARG, "lsh_flatten: bad argument");
    const HostTree& T = L->trees[tree];
    VERS_CUDA(cudaSetDevice(L->ctx->device));
    cudaStream_t s = L->ctx->stream;
    // preorder: node, above subtree (right), below subtree (left)
    std::vector<uint32_t> stack{0};
    uint32_t nn = 0, ni = 0;
    uint64_t nit = 0;
    while (!stack.empty()) {
        uint32_t node = stack.back();
        stack.pop_back();
        if (kind) kind[nn] = T.kind[node];
        if (T.kind[node] == 1) {
            if (leaf_len) leaf_len[nn] = T.leaf_len[node];
            if (items && T.leaf_len[node])
                VERS_CUDA(cudaMemcpyAsync(items + nit, L->d_leaf_items + (uint64_t)T.slot[node] * L->slot_cap,
                                          (size_t)T.leaf_len[node] * 4, cudaMemcpyDeviceToHost, s));
            nit += T.leaf_len[node];
        } else {
            if (leaf_len) leaf_len[nn] = 0;
            if (planes)
                VERS_CUDA(cudaMemcpyAsync(planes + (uint64_t)ni * L->dim, L->d_planes + (uint64_t)T.plane[node] * L->ld,
                                          (size_t)L->dim * 4, cudaMemcpyDeviceToHost, s));
            if (consts)
                VERS_CUDA(cudaMemcpyAsync(consts + ni, L->d_consts + T.plane[node], 4, cudaMemcpyDeviceToHost, s));
            ++ni;
            stack.push_back(T.left[node]);   // popped second
            stack.push_back(T.right[node]);  // popped first: above subtree comes first
        }
        ++nn;
    }
    VERS_CUDA(cudaStreamSynchronize(s));
    *n_nodes = nn;
    *n_inner = ni;
    *n_items = nit;
    return VERS_OK;
}

extern "C" int32_t vers_lsh_search(vers_lsh* L, const float* queries, uint32_t nq, uint32_t q_stride_floats,
                                   uint32_t top_k, uint64_t* ids, float* dists, uint32_t* counts) {
    if (!L || (!queries && nq) || (!ids && nq && top_k) || (!dists && nq && top_k))
        return fail(VERS_ERR_ARG, "lsh_search: null argument");
    if (q_stride_floats < L->dim) return fail(VERS_ERR_ARG, "lsh_search: query stride < dim");
    if (top_k > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "top_k %u > %u", top_k, VERS_MAX_TOPK);
    if (nq == 0) return VERS_OK;
    if (top_k == 0 || L->num_trees == 0) {
        if (counts) memset(counts, 0, sizeof(uint32_t) * nq);
        return VERS_OK;
    }
    vers_ctx* ctx = L->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    VERS_TRY(sync_device_mirror(L));
    // every leaf is visited at most once per traversal, so L->n candidates per tree always suffice; start small and
    // retry larger in the rare case the reference's cascading backtracking (lsh.rs:210-213) visits many leaves
    uint64_t cand_cap = 24ull * top_k + 2ull * L->slot_cap + 64;
    const size_t nk = (size_t)nq * top_k, npt = (size_t)nq * L->num_trees;
    ForestDev f = forest_dev(L);
    for (;;) {
        if (cand_cap > L->n) cand_cap = std::max<uint64_t>(L->n, 1);
        ScratchCarver plan(nullptr);
        plan.plan<float>((size_t)nq * L->ld);
        plan.plan<uint64_t>(nk);
        plan.plan<float>(nk);
        plan.plan<uint32_t>(nq);
        plan.plan<uint32_t>(npt * cand_cap);
        plan.plan<uint32_t>(npt);
        plan.plan<uint32_t>(4);
        VERS_TRY(scratch_reserve(ctx, plan.off + 256));
        ScratchCarver sc(ctx->scratch);
        float* d_q = sc.take<float>((size_t)nq * L->ld);
        uint64_t* d_ids = sc.take<uint64_t>(nk);
        float* d_d = sc.take<float>(nk);
        uint32_t* d_c = sc.take<uint32_t>(nq);
        uint32_t* d_cand = sc.take<uint32_t>(npt * cand_cap);
        uint32_t* d_cnt = sc.take<uint32_t>(npt);
        uint32_t* d_over = sc.take<uint32_t>(4);
        if (L->ld != L->dim) VERS_CUDA(cudaMemsetAsync(d_q, 0, (size_t)nq * L->ld * 4, ctx->stream));
        VERS_CUDA(cudaMemcpy2DAsync(d_q, (size_t)L->ld * 4, queries, (size_t)q_stride_floats * 4, (size_t)L->dim * 4, nq,
                                    cudaMemcpyHostToDevice, ctx->stream));
        VERS_CUDA(cudaMemsetAsync(d_over, 0, 16, ctx->stream));
        {
            // 32 dimensions per staged chunk: 8.7 KB of staging per warp keeps ~18 of these latency-bound warps on an SM
            // (measured on C3, 1000-query batch: 64-dimension chunks 1.51 ms per batch, 32: 1.09 ms, with 2-warp blocks and
            // the plane products sharing the staging buffer 1.01 ms; 16-dimension chunks the same)
            size_t per_warp = (FrCfg<32>::STAGE_BYTES + (size_t)L->ld * 4 + LSH_STACK * 12 + (size_t)top_k * 8 + 15) & ~size_t(15);
            size_t smem = per_warp * FT_WPB;
            auto kern = forest_traverse_kernel<32>;
            VERS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)ceil_div(npt, FT_WPB), FT_WPB * 32, smem, ctx->stream>>>(f, d_q, nq, top_k, (uint32_t)cand_cap,
                                                                                  d_cand, d_cnt, d_over);
            VERS_LAUNCH_CHECK(ctx);
        }
        uint32_t over = 0;
        VERS_CUDA(cudaMemcpyAsync(&over, d_over, 4, cudaMemcpyDeviceToHost, ctx->stream));
        VERS_CUDA(cudaStreamSynchronize(ctx->stream));
        if (over) {
            if (cand_cap >= L->n)
                return fail(VERS_ERR_UNSUPPORTED, "lsh_search: a traversal needs more than %d stack frames", LSH_STACK);
            cand_cap *= 8;
            continue;
        }
        {
            size_t per_warp = (FrCfg<64>::STAGE_BYTES + (size_t)L->ld * 4 + (size_t)top_k * 8 + 15) & ~size_t(15);
            size_t smem = per_warp * 4;
            VERS_CUDA(cudaFuncSetAttribute(forest_rerank_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            forest_rerank_kernel<64><<<(unsigned)ceil_div(nq, 4), 128, smem, ctx->stream>>>(
                f, d_q, nq, top_k, (uint32_t)cand_cap, d_cand, d_cnt, L->d_ids, d_ids, d_d, d_c);
            VERS_LAUNCH_CHECK(ctx);
        }
        VERS_CUDA(cudaMemcpyAsync(ids, d_ids, nk * 8, cudaMemcpyDeviceToHost, ctx->stream));
        VERS_CUDA(cudaMemcpyAsync(dists, d_d, nk * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (counts) VERS_CUDA(cudaMemcpyAsync(counts, d_c, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
        VERS_CUDA(cudaStreamSynchronize(ctx->stream));
        return VERS_OK;
    }
}

extern "C" int32_t vers_lsh_add(vers_lsh* L, const float* embedding, uint64_t vec_id) {
    if (!L || !embedding) return fail(VERS_ERR_ARG, "lsh_add: null argument");
    vers_ctx* ctx = L->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    // insert (lsh.rs:218-251) stores vec_id ITSELF in the leaf as a row index (quirk kept).  The reference would only
    // panic later (values[vec_id] out of bounds during a split or a search); we refuse up front — BEFORE anything is
    // mutated — rather than store an index that a kernel would dereference out of bounds.
    if (L->num_trees && (vec_id >= 0xffffffffull || vec_id >= L->n + 1))
        return fail(VERS_ERR_PANIC, "lsh_add: vec_id %llu is not a row index (lsh.rs:247 stores it as one)",
                    (unsigned long long)vec_id);
    // values.push(embedding); ids.push(vec_id)  (lsh.rs:257-258)
    if (L->n + 1 > L->cap) {
        uint64_t capw = L->cap * L->ld;
        VERS_TRY(grow_device(&L->d_values, &capw, (L->cap * 2 + 64) * L->ld, L->n * L->ld, s));
        L->cap = capw / L->ld;
    }
    std::vector<float> row(L->ld, 0.0f);
    memcpy(row.data(), embedding, (size_t)L->dim * 4);
    VERS_CUDA(cudaMemcpyAsync(L->d_values + L->n * L->ld, row.data(), (size_t)L->ld * 4, cudaMemcpyHostToDevice, s));
    const uint64_t new_row = L->n;
    L->n += 1;
    L->ids.push_back(vec_id);
    // the device mirrors follow in place (one id, one leaf length per tree): re-uploading ten megabytes of node arrays
    // and ids after every add was most of a 1.6 ms call
    if (!L->ids_dirty && L->d_ids && L->ids.size() <= L->ids_cap)
        VERS_CUDA(cudaMemcpyAsync(L->d_ids + (L->ids.size() - 1), &L->ids.back(), 8, cudaMemcpyHostToDevice, s));
    else
        L->ids_dirty = true;
    if (L->num_trees == 0) return VERS_OK;
    VERS_TRY(sync_device_mirror(L));
    const uint32_t member = (uint32_t)vec_id;
    VERS_TRY(io_reserve(ctx, (size_t)L->num_trees * 4 + 256));
    uint32_t* d_leaf = reinterpret_cast<uint32_t*>(ctx->io);
    ForestDev f = forest_dev(L);
    forest_descend_kernel<<<L->num_trees, 32, (size_t)L->ld * 8, s>>>(f, L->d_values + new_row * L->ld, d_leaf);
    ctx->launches += 1;
    std::vector<uint32_t> leaf(L->num_trees);
    cudaError_t e = cudaMemcpyAsync(leaf.data(), d_leaf, (size_t)L->num_trees * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(VERS_ERR_CUDA, "lsh_add: %s", cudaGetErrorString(e));
    std::vector<uint32_t> new_len(L->num_trees, 0);  // staging of the in-place leaf-length updates (alive until the sync)
    uint64_t tree_base = 0;
    const bool mirror_clean = !L->nodes_dirty;
    std::vector<BuildRoot> splits;
    std::vector<uint32_t> split_members;
    for (uint32_t t = 0; t < L->num_trees; ++t) {
        HostTree& T = L->trees[t];
        uint32_t node = leaf[t];
        uint32_t len = T.leaf_len[node];
        if (len + 1 > L->max_size) {  // lsh.rs:240-245: rebuild this leaf (+ the new id) as a subtree
            std::vector<uint32_t> mem(len + 1);
            if (len)
                VERS_CUDA(cudaMemcpy(mem.data(), L->d_leaf_items + (uint64_t)T.slot[node] * L->slot_cap, (size_t)len * 4,
                                     cudaMemcpyDeviceToHost));
            mem[len] = member;
            for (uint32_t m : mem)
                if (m >= L->n) return fail(VERS_ERR_PANIC, "lsh_add: leaf member %u is not a row (vec_id used as an index, lsh.rs:247)", m);
            splits.push_back(BuildRoot{t, node, (uint64_t)split_members.size(), len + 1, T.hash[node]});
            split_members.insert(split_members.end(), mem.begin(), mem.end());
        } else {
            VERS_CUDA(cudaMemcpyAsync(L->d_leaf_items + (uint64_t)T.slot[node] * L->slot_cap + len, &member, 4,
                                      cudaMemcpyHostToDevice, s));
            T.leaf_len[node] = len + 1;
            new_len[t] = len + 1;
            if (mirror_clean)
                VERS_CUDA(cudaMemcpyAsync(L->d_leaf_len + tree_base + node, &new_len[t], 4, cudaMemcpyHostToDevice, s));
            else
                L->nodes_dirty = true;
        }
        tree_base += T.kind.size();
    }
    VERS_CUDA(cudaStreamSynchronize(s));
    if (!splits.empty()) {
        uint32_t *d_a = nullptr, *d_b = nullptr;
        size_t tot = split_members.size();
        VERS_CUDA(cudaMalloc(&d_a, tot * 4));
        if (cudaMalloc(&d_b, tot * 4) != cudaSuccess) {
            cudaFree(d_a);
            return fail(VERS_ERR_NOMEM, "lsh_add: cudaMalloc");
        }
        e = cudaMemcpyAsync(d_a, split_members.data(), tot * 4, cudaMemcpyHostToDevice, s);
        int32_t rc = e == cudaSuccess ? build_subtrees(L, splits, d_a, d_b, tot)
                                      : fail(VERS_ERR_CUDA, "lsh_add: %s", cudaGetErrorString(e));
        cudaFree(d_a);
        cudaFree(d_b);
        if (rc != VERS_OK) return rc;
    }
    return VERS_OK;
}

// values / ids of the index (the deduplicated rows in their stored order: what Index::save_index serialises,
// lsh.rs:47-55), including rows added after build_index
extern "C" int32_t vers_lsh_get_values(const vers_lsh* L, float* values, uint32_t stride_floats, uint64_t* ids) {
    if (!L) return fail(VERS_ERR_ARG, "lsh_get_values: null");
    if (values && stride_floats < L->dim) return fail(VERS_ERR_ARG, "lsh_get_values: stride < dim");
    VERS_CUDA(cudaSetDevice(L->ctx->device));
    if (values && L->n) {
        VERS_CUDA(cudaMemcpy2DAsync(values, (size_t)stride_floats * 4, L->d_values, (size_t)L->ld * 4, (size_t)L->dim * 4,
                                    L->n, cudaMemcpyDeviceToHost, L->ctx->stream));
        VERS_CUDA(cudaStreamSynchronize(L->ctx->stream));
    }
    if (ids) memcpy(ids, L->ids.data(), L->ids.size() * 8);
    return VERS_OK;
}

// The forest from deserialised parts (after Index::load_index, base.rs:45-58): `values` are the stored (already
// deduplicated) rows, every tree arrives in the preorder of vers_lsh_flatten (node, ABOVE subtree, BELOW subtree), the
// trees concatenated: kind / leaf_len per node, planes / consts per inner node, items per leaf.
extern "C" int32_t vers_lsh_from_parts(vers_ctx* ctx, const float* values, uint64_t n, uint32_t dim,
                                       uint32_t stride_floats, const uint64_t* ids, uint32_t num_trees, uint32_t max_size,
                                       uint64_t seed, const uint32_t* tree_nodes, const uint8_t* kind,
                                       const uint32_t* leaf_len, const float* planes, const float* consts,
                                       const uint32_t* items, vers_lsh** out) {
    if (!ctx || !out || (!values && n) || (num_trees && (!tree_nodes || !kind || !leaf_len)))
        return fail(VERS_ERR_ARG, "lsh_from_parts: null argument");
    *out = nullptr;
    if (dim == 0 || stride_floats < dim) return fail(VERS_ERR_ARG, "lsh_from_parts: bad dim/stride");
    if (n >= 0xffffffffull) return fail(VERS_ERR_UNSUPPORTED, "lsh_from_parts: more than 2^32-2 rows");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    vers_lsh* L = new vers_lsh();
    L->ctx = ctx;
    L->dim = dim;
    L->ld = round_up(dim, 4);
    L->max_size = max_size;
    L->num_trees = num_trees;
    L->slot_cap = max_size + 1;
    L->seed = seed;
    L->n = n;
    L->cap = n + n / 8 + 64;
    L->ids.resize(n);
    for (uint64_t r = 0; r < n; ++r) L->ids[r] = ids ? ids[r] : r;
    L->trees.resize(num_trees);
    // host: node arrays, planes and leaf slots from the preorder walk
    std::vector<float> h_planes, h_consts;
    std::vector<uint32_t> h_items;
    uint64_t node_at = 0, plane_at = 0, item_at = 0;
    auto bail = [&](int32_t code) {
        vers_lsh_free(L);
        return code;
    };
    for (uint32_t t = 0; t < num_trees; ++t) {
        HostTree& T = L->trees[t];
        const uint64_t node_end = node_at + tree_nodes[t];
        struct Todo {
            uint32_t node;
            uint64_t hash;
        };
        std::vector<Todo> stack;
        stack.push_back(Todo{T.add_node(), vers_lsh_root_hash(seed, t)});
        while (!stack.empty()) {
            const Todo cur = stack.back();
            stack.pop_back();
            if (node_at >= node_end) return bail(fail(VERS_ERR_ARG, "lsh_from_parts: tree %u is truncated", t));
            const uint64_t i = node_at++;
            T.hash[cur.node] = cur.hash;
            if (kind[i] == 1) {
                const uint32_t len = leaf_len[i];
                if (len > L->slot_cap)
                    return bail(fail(VERS_ERR_UNSUPPORTED, "lsh_from_parts: a leaf of %u rows exceeds max_size + 1 = %u", len,
                                     L->slot_cap));
                if (len && !items) return bail(fail(VERS_ERR_ARG, "lsh_from_parts: null items"));
                T.kind[cur.node] = 1;
                T.leaf_len[cur.node] = len;
                T.slot[cur.node] = L->n_slots++;
                const size_t at = h_items.size();
                h_items.resize(at + L->slot_cap, 0u);
                for (uint32_t e = 0; e < len; ++e) {
                    if (items[item_at + e] >= n)
                        return bail(fail(VERS_ERR_PANIC, "lsh_from_parts: leaf member %u is not a row (index out of bounds)",
                                         items[item_at + e]));
                    h_items[at + e] = items[item_at + e];
                }
                item_at += len;
            } else if (kind[i] == 0) {
                if (!planes || !consts) return bail(fail(VERS_ERR_ARG, "lsh_from_parts: null planes"));
                T.kind[cur.node] = 0;
                T.plane[cur.node] = L->n_planes++;
                const size_t at = h_planes.size();
                h_planes.resize(at + L->ld, 0.0f);
                memcpy(h_planes.data() + at, planes + plane_at * dim, (size_t)dim * 4);
                h_consts.push_back(consts[plane_at]);
                ++plane_at;
                const uint32_t nr = T.add_node();  // above -> right_node (lsh.rs:108)
                const uint32_t nl = T.add_node();  // below -> left_node  (lsh.rs:107)
                T.right[cur.node] = nr;
                T.left[cur.node] = nl;
                stack.push_back(Todo{nl, vers_lsh_child_hash(cur.hash, 0)});  // popped second: the below subtree
                stack.push_back(Todo{nr, vers_lsh_child_hash(cur.hash, 1)});  // popped first: the above subtree
            } else {
                return bail(fail(VERS_ERR_ARG, "lsh_from_parts: bad node kind %u", (unsigned)kind[i]));
            }
        }
        if (node_at != node_end) return bail(fail(VERS_ERR_ARG, "lsh_from_parts: tree %u has trailing nodes", t));
    }
    // device: rows, planes, leaf slots (with the same slack build_index leaves for later adds)
    L->cap_planes = L->n_planes + L->n_planes / 2 + 64;
    L->cap_slots = L->n_slots + L->n_slots / 2 + 64;
    cudaError_t e = cudaMalloc(&L->d_values, (size_t)L->cap * L->ld * 4);
    if (e == cudaSuccess) e = cudaMalloc(&L->d_planes, (size_t)L->cap_planes * L->ld * 4);
    if (e == cudaSuccess) e = cudaMalloc(&L->d_consts, (size_t)L->cap_planes * 4);
    if (e == cudaSuccess) e = cudaMalloc(&L->d_leaf_items, (size_t)L->cap_slots * L->slot_cap * 4);
    if (e == cudaSuccess && n) {
        e = cudaMemsetAsync(L->d_values, 0, (size_t)n * L->ld * 4, s);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(L->d_values, (size_t)L->ld * 4, values, (size_t)stride_floats * 4, (size_t)dim * 4, n,
                                  cudaMemcpyHostToDevice, s);
    }
    if (e == cudaSuccess && L->n_planes) {
        e = cudaMemcpyAsync(L->d_planes, h_planes.data(), h_planes.size() * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(L->d_consts, h_consts.data(), h_consts.size() * 4, cudaMemcpyHostToDevice, s);
    }
    if (e == cudaSuccess && L->n_slots)
        e = cudaMemcpyAsync(L->d_leaf_items, h_items.data(), h_items.size() * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
        return bail(fail(e == cudaErrorMemoryAllocation ? VERS_ERR_NOMEM : VERS_ERR_CUDA, "lsh_from_parts: %s",
                         cudaGetErrorString(e)));
    L->nodes_dirty = true;
    L->ids_dirty = true;
    *out = L;
    return VERS_OK;
}
