// peer.cuh — exchange kernels over NVLink peer memory (CUDA IPC mapped buffers, plain remote stores + flags), shared
// by peer.cu (the stand-alone vers_peer_* API) and comm.cu (the sharded index behind vers_comm / vers_sharded_*).
//
// One EXCHANGE REGION in every rank's buffer:  data [2 parities][world][slot_bytes] | flags [2][world] u32 |
// ctl { done_pub, done_con, step } u32.  Rank r's contribution to step s lives in slot [s & 1][r] of EVERY rank's
// region (written by r with remote stores); flags[s & 1][r] = s on a rank says "r's slot of step s is complete here".
//
// Parity = step & 1.  A slot of parity p is rewritten at step s + 2; the writer only gets there after it consumed
// step s + 1, which needs every peer's step s + 1 flag, which a peer raises only after (stream order) its own
// consumption of step s — all its reads of parity p — has finished: no reader can still be in a slot that is being
// overwritten.
//
// The step number lives in DEVICE memory (ctl.step) when the caller wants the launches to be capturable in a CUDA
// graph: every block reads it on entry, the last block to leave the consuming kernel increments it.  Replaying the
// graph then advances the protocol exactly like eager launches do.
//
// No kernel here waits on another block of its OWN grid: a waiting block only depends on kernels of OTHER GPUs that
// never wait before publishing (publish kernels) or on its own grid's blocks that have all been counted in before the
// flag goes up — and those grids are sized to be fully resident (see peer_resident_blocks).
#pragma once
#include "engine.cuh"

namespace vers {

constexpr int PG_WARPS = 4;

struct PeerRegion {
    char* const* peer_base;  // [world] device-visible base pointers of every rank's buffer (own entry included)
    uint64_t data_off;       // region offset inside a buffer
    uint64_t slot_bytes;
    uint64_t flags_off;
    uint64_t ctl_off;        // done_pub, done_con, step
    uint32_t world, rank;
};

__device__ __forceinline__ unsigned long long pg_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void peer_wait_flags(const PeerRegion& g, uint32_t step) {  // one thread
    const char* mine = g.peer_base[g.rank];
    const uint32_t parity = step & 1u;
    const unsigned long long t0 = pg_now_ns();
    for (uint32_t r = 0; r < g.world; ++r) {
        const volatile uint32_t* f =
            reinterpret_cast<const volatile uint32_t*>(mine + g.flags_off) + (uint64_t)parity * g.world + r;
        while ((int32_t)(*f - step) < 0) {
            if (pg_now_ns() - t0 > 20000000000ull) __trap();  // 20 s: a dead peer traps instead of hanging the GPU
        }
    }
    __threadfence_system();
}

// after the whole grid has stored (every block calls this once, all threads): the last block raises this rank's flag
// of `step` on every rank
__device__ __forceinline__ void peer_block_published(const PeerRegion& g, uint32_t step) {
    __threadfence_system();  // this block's remote stores are visible system-wide before it reports in
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t* ctl = reinterpret_cast<uint32_t*>(g.peer_base[g.rank] + g.ctl_off);
        const uint32_t prev = atomicAdd(&ctl[0], 1u);
        if (prev == gridDim.x - 1) {
            ctl[0] = 0;  // self-cleaning for the next step
            __threadfence_system();
            const uint32_t parity = step & 1u;
            for (uint32_t r = 0; r < g.world; ++r) {
                volatile uint32_t* f =
                    reinterpret_cast<volatile uint32_t*>(g.peer_base[r] + g.flags_off) + (uint64_t)parity * g.world + g.rank;
                *f = step;
            }
        }
    }
}

// the last block to leave the consuming kernel advances the device-resident step
__device__ __forceinline__ void peer_block_consumed(const PeerRegion& g) {
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t* ctl = reinterpret_cast<uint32_t*>(g.peer_base[g.rank] + g.ctl_off);
        const uint32_t prev = atomicAdd(&ctl[1], 1u);
        if (prev == gridDim.x - 1) {
            ctl[1] = 0;
            ctl[2] += 1;
        }
    }
}

__device__ __forceinline__ uint32_t peer_step(const PeerRegion& g, uint32_t host_step) {
    if (host_step) return host_step;
    return *reinterpret_cast<const volatile uint32_t*>(g.peer_base[g.rank] + g.ctl_off + 8);
}

// ---- all-gather, part 1: my `bytes` (multiple of 16) into my slot on every rank
static __global__ void __launch_bounds__(256) peer_publish_kernel(PeerRegion g, const uint4* __restrict__ src, uint64_t bytes,
                                                           uint32_t host_step) {
    const uint32_t step = peer_step(g, host_step);
    const uint64_t my_slot = g.data_off + ((uint64_t)(step & 1u) * g.world + g.rank) * g.slot_bytes;
    const uint64_t n16 = bytes >> 4;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        for (uint32_t r = 0; r < g.world; ++r) reinterpret_cast<uint4*>(g.peer_base[r] + my_slot)[i] = v;
    }
    peer_block_published(g, step);
}

// ---- all-gather, part 2: wait for every rank's flag, copy the first `elems` 8-byte words of each of the world slots
// to dst [world][elems] (compact: the slots themselves are padded to 16 bytes)
static __global__ void __launch_bounds__(256) peer_wait_copy_kernel(PeerRegion g, unsigned long long* __restrict__ dst,
                                                                    uint64_t elems, uint32_t host_step) {
    const uint32_t step = peer_step(g, host_step);
    if (threadIdx.x == 0) peer_wait_flags(g, step);
    __syncthreads();
    const char* mine = g.peer_base[g.rank] + g.data_off + (uint64_t)(step & 1u) * g.world * g.slot_bytes;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems * g.world;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / elems);
        const uint64_t j = i - (uint64_t)r * elems;
        dst[i] = __ldcg(reinterpret_cast<const unsigned long long*>(mine + (uint64_t)r * g.slot_bytes) + j);
    }
    if (!host_step) peer_block_consumed(g);
}

// diagnostic: globaltimer stamps of block 0 of the most recent peer_gather_merge_kernel launch of this translation unit
// (entry, own slice published, every rank's flag seen, merged) — vers_debug_peer_times
static __device__ unsigned long long g_peer_dbg[4];

// ---- exchange + merge of the per-GPU top-k as ONE kernel: every warp stores its queries' local top-k into this
// rank's slot on every rank (ids [nq][k] then distances [nq][k]), the last block raises the flag, then each warp waits
// for all ranks' flags and merges world x k entries per query by (distance, id) — the same total order as
// merge_ids_kernel, so the result does not depend on arrival order.  The grid must be fully resident.
static __global__ void __launch_bounds__(PG_WARPS * 32)
    peer_gather_merge_kernel(PeerRegion g, uint32_t host_step, const uint64_t* __restrict__ loc_ids,
                             const float* __restrict__ loc_d, uint32_t nq, uint32_t k, uint64_t* out_ids, float* out_d,
                             uint32_t* out_cnt) {
    extern __shared__ __align__(16) unsigned char pgsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t step = peer_step(g, host_step);
    const uint32_t parity = step & 1u;
    const uint64_t nk = (uint64_t)nq * k;
    const uint64_t my_slot = g.data_off + ((uint64_t)parity * g.world + g.rank) * g.slot_bytes;
    const uint32_t q0 = blockIdx.x * PG_WARPS + warp, qstride = gridDim.x * PG_WARPS;
    const bool dbg = blockIdx.x == 0 && threadIdx.x == 0;
    if (dbg) g_peer_dbg[0] = pg_now_ns();

    // 1. publish
    for (uint32_t q = q0; q < nq; q += qstride) {
        for (uint32_t r = 0; r < g.world; ++r) {
            char* base = g.peer_base[r] + my_slot;
            uint64_t* pi = reinterpret_cast<uint64_t*>(base) + (uint64_t)q * k;
            float* pd = reinterpret_cast<float*>(base + nk * 8) + (uint64_t)q * k;
            for (uint32_t e = lane; e < k; e += 32) {
                pi[e] = loc_ids[(uint64_t)q * k + e];
                pd[e] = loc_d[(uint64_t)q * k + e];
            }
        }
    }
    peer_block_published(g, step);
    if (dbg) g_peer_dbg[1] = pg_now_ns();

    // 2. wait for every rank's flag of this step
    if (threadIdx.x == 0) peer_wait_flags(g, step);
    __syncthreads();
    if (dbg) g_peer_dbg[2] = pg_now_ns();

    // 3. merge world x k entries per query by (distance, id)
    const char* mine = g.peer_base[g.rank] + g.data_off + (uint64_t)parity * g.world * g.slot_bytes;
    uint64_t* sp = reinterpret_cast<uint64_t*>(pgsm) + (size_t)warp * k;
    float* sd = reinterpret_cast<float*>(pgsm + (size_t)PG_WARPS * k * 8) + (size_t)warp * k;
    const uint32_t total = g.world * k;
    for (uint32_t q = q0; q < nq; q += qstride) {
        for (uint32_t e = lane; e < k; e += 32) {
            sd[e] = __int_as_float(0x7f800000);
            sp[e] = 0xffffffffffffffffull;
        }
        __syncwarp();
        for (uint32_t e0 = 0; e0 < total; e0 += 32) {
            const uint32_t e = e0 + lane;
            float v = 0.f;
            uint64_t id = 0xffffffffffffffffull;
            if (e < total) {
                const uint32_t r = e / k, j = e % k;
                const char* base = mine + (uint64_t)r * g.slot_bytes;
                id = __ldcg(reinterpret_cast<const uint64_t*>(base) + (uint64_t)q * k + j);
                v = __ldcg(reinterpret_cast<const float*>(base + nk * 8) + (uint64_t)q * k + j);
            }
            bool live = id != 0xffffffffffffffffull;
            while (true) {
                bool pass = live && entry_less<uint64_t>(v, id, sd[k - 1], sp[k - 1]);
                unsigned m = __ballot_sync(FULL_MASK, pass);
                if (!m) break;
                int src = __ffs(m) - 1;
                float bv = __shfl_sync(FULL_MASK, v, src);
                uint64_t bid = __shfl_sync(FULL_MASK, id, src);
                warp_topk_insert<uint64_t>(sd, sp, (int)k, bv, bid, lane);
                if (lane == src) live = false;
            }
        }
        uint32_t cnt = 0;
        for (uint32_t e0 = 0; e0 < k; e0 += 32) {
            const uint32_t e = e0 + lane;
            bool have = false;
            if (e < k) {
                out_ids[(uint64_t)q * k + e] = sp[e];
                out_d[(uint64_t)q * k + e] = sd[e];
                have = sp[e] != 0xffffffffffffffffull;
            }
            cnt += __popc(__ballot_sync(FULL_MASK, have));
        }
        if (out_cnt && lane == 0) out_cnt[q] = cnt;
        __syncwarp();
    }
    if (dbg) g_peer_dbg[3] = pg_now_ns();
    if (!host_step) peer_block_consumed(g);
}

// blocks of peer_gather_merge_kernel that can be resident at once with `smem` dynamic bytes per block
inline int32_t peer_resident_blocks(vers_ctx* ctx, size_t smem, unsigned* out) {
    int per_sm = 0;
    VERS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer_gather_merge_kernel, PG_WARPS * 32, smem));
    if (per_sm < 1) return fail(VERS_ERR_UNSUPPORTED, "peer exchange: the merge kernel does not fit an SM");
    *out = (unsigned)per_sm * (unsigned)ctx->sm_count;
    return VERS_OK;
}

}  // namespace vers
