// peer.cu — the exchange step of a sharded search as ONE kernel over NVLink peer memory.
//
// After the list scan + exact rerank every GPU holds its local top-k per query.  The NCCL formulation is
// all-gather (nq x k x 12 B per rank) -> merge kernel: two launches and a collective whose cost at these sizes is pure
// latency.  Here one kernel does both: every warp owns a query, stores that query's local top-k straight into the
// slot this rank owns in EVERY peer's exchange buffer (peer pointers opened with CUDA IPC: plain stores over
// NVLink/NVSwitch), the last block to finish raises this rank's flag on every peer, then each warp waits for the
// flags of all ranks and merges the world x k entries of its query by (distance, id) — the same total order as
// merge_ids_kernel, so the result does not depend on arrival order.
//
// Buffer of one rank (cudaMalloc, IPC-exported): [2 parities][world][slot_bytes] | flags [2][world] u32 | done u32.
// Parity = step & 1.  A slot of parity p is rewritten at step s + 2; the writer can only get there after it has
// merged step s + 1, which needs every peer's step s + 1 flag, which a peer raises only after its own step s kernel
// (all its reads of parity p) has finished: no reader can still be in a slot that is being overwritten.
#include <algorithm>

#include "engine.cuh"

struct vers_peer {
    vers_ctx* ctx = nullptr;
    uint32_t world = 0, rank = 0;
    uint64_t slot_bytes = 0;
    char* d_buf = nullptr;         // this rank's exchange buffer
    char** d_peer_base = nullptr;  // [world] device-visible base pointers (own entry = d_buf)
    std::vector<char*> opened;     // peer mappings to close
    uint32_t step = 0;
    size_t flags_off = 0, done_off = 0, total = 0;
};

namespace vers {

constexpr int PG_WARPS = 4;

__device__ __forceinline__ unsigned long long pg_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(PG_WARPS * 32)
    peer_gather_merge_kernel(char* const* __restrict__ peer_base, uint32_t world, uint32_t rank, uint64_t slot_bytes,
                             uint64_t flags_off, uint64_t done_off, uint32_t step, const uint64_t* __restrict__ loc_ids,
                             const float* __restrict__ loc_d, uint32_t nq, uint32_t k, uint64_t* out_ids, float* out_d,
                             uint32_t* out_cnt) {
    extern __shared__ __align__(16) unsigned char pgsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * PG_WARPS + warp;
    const uint32_t parity = step & 1u;
    const uint64_t nk = (uint64_t)nq * k;
    const uint64_t my_slot = ((uint64_t)parity * world + rank) * slot_bytes;

    // 1. publish: this query's local top-k into my slot on every rank (ids [nq][k] then distances [nq][k])
    if (q < nq) {
        for (uint32_t r = 0; r < world; ++r) {
            char* base = peer_base[r] + my_slot;
            uint64_t* pi = reinterpret_cast<uint64_t*>(base) + (uint64_t)q * k;
            float* pd = reinterpret_cast<float*>(base + nk * 8) + (uint64_t)q * k;
            for (uint32_t e = lane; e < k; e += 32) {
                pi[e] = loc_ids[(uint64_t)q * k + e];
                pd[e] = loc_d[(uint64_t)q * k + e];
            }
        }
    }
    __threadfence_system();  // my stores are visible system-wide before the block reports in
    __syncthreads();
    char* mine = peer_base[rank];
    if (threadIdx.x == 0) {
        uint32_t* done = reinterpret_cast<uint32_t*>(mine + done_off);
        const uint32_t prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {  // last block of this rank: everything is published, raise my flag everywhere
            *done = 0;                // self-cleaning for the next step
            __threadfence_system();
            for (uint32_t r = 0; r < world; ++r) {
                volatile uint32_t* f =
                    reinterpret_cast<volatile uint32_t*>(peer_base[r] + flags_off) + (uint64_t)parity * world + rank;
                *f = step;
            }
        }
    }
    if (q >= nq) return;

    // 2. wait for every rank's flag of this step (bounded: a dead peer traps instead of hanging the GPU)
    if (lane == 0) {
        const unsigned long long t0 = pg_now_ns();
        for (uint32_t r = 0; r < world; ++r) {
            volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(mine + flags_off) + (uint64_t)parity * world + r;
            while ((int32_t)(*f - step) < 0) {
                if (pg_now_ns() - t0 > 20000000000ull) __trap();  // 20 s
            }
        }
        __threadfence_system();
    }
    __syncwarp();

    // 3. merge world x k entries of this query by (distance, id)
    uint64_t* sp = reinterpret_cast<uint64_t*>(pgsm) + (size_t)warp * k;
    float* sd = reinterpret_cast<float*>(pgsm + (size_t)PG_WARPS * k * 8) + (size_t)warp * k;
    for (uint32_t e = lane; e < k; e += 32) {
        sd[e] = __int_as_float(0x7f800000);
        sp[e] = 0xffffffffffffffffull;
    }
    __syncwarp();
    const uint32_t total = world * k;
    for (uint32_t e0 = 0; e0 < total; e0 += 32) {
        const uint32_t e = e0 + lane;
        float v = 0.f;
        uint64_t id = 0xffffffffffffffffull;
        if (e < total) {
            const uint32_t r = e / k, j = e % k;
            const char* base = mine + ((uint64_t)parity * world + r) * slot_bytes;
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
}

}  // namespace vers

using namespace vers;

extern "C" int32_t vers_peer_create(vers_ctx* ctx, uint32_t world, uint32_t rank, uint64_t slot_bytes, vers_peer** out,
                                    uint8_t ipc_handle_out[64]) {
    if (!ctx || !out || !ipc_handle_out || world == 0 || rank >= world || slot_bytes == 0)
        return fail(VERS_ERR_ARG, "peer_create: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    *out = nullptr;
    VERS_CUDA(cudaSetDevice(ctx->device));
    vers_peer* p = new vers_peer();
    p->ctx = ctx;
    p->world = world;
    p->rank = rank;
    p->slot_bytes = (slot_bytes + 255) & ~uint64_t(255);
    p->flags_off = (size_t)2 * world * p->slot_bytes;
    p->done_off = p->flags_off + (size_t)2 * world * 4;
    p->total = p->done_off + 256;
    cudaError_t e = cudaMalloc(&p->d_buf, p->total);
    if (e == cudaSuccess) e = cudaMemset(p->d_buf, 0, p->total);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_peer_base, sizeof(char*) * world);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->d_buf);
    if (e != cudaSuccess) {
        cudaFree(p->d_buf);
        cudaFree(p->d_peer_base);
        delete p;
        return fail(VERS_ERR_CUDA, "peer_create: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle_out, &h, 64);
    *out = p;
    return VERS_OK;
}

extern "C" int32_t vers_peer_connect(vers_peer* p, const uint8_t* all_handles) {
    if (!p || !all_handles) return fail(VERS_ERR_ARG, "peer_connect: null");
    VERS_CUDA(cudaSetDevice(p->ctx->device));
    std::vector<char*> base(p->world, nullptr);
    for (uint32_t r = 0; r < p->world; ++r) {
        if (r == p->rank) {
            base[r] = p->d_buf;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)r * 64, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(VERS_ERR_CUDA, "peer_connect: rank %u: %s", r, cudaGetErrorString(e));
        base[r] = (char*)ptr;
        p->opened.push_back((char*)ptr);
    }
    VERS_CUDA(cudaMemcpy(p->d_peer_base, base.data(), sizeof(char*) * p->world, cudaMemcpyHostToDevice));
    return VERS_OK;
}

extern "C" int32_t vers_peer_gather_merge_dev(vers_peer* p, const uint64_t* d_local_ids, const float* d_local_dists,
                                              uint32_t nq, uint32_t top_k, uint64_t* d_ids, float* d_dists,
                                              uint32_t* d_counts) {
    if (!p || !d_local_ids || !d_local_dists || !d_ids || !d_dists) return fail(VERS_ERR_ARG, "peer_gather_merge: null");
    if (top_k > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "top_k %u > %u", top_k, VERS_MAX_TOPK);
    if (nq == 0 || top_k == 0) return VERS_OK;
    if ((uint64_t)nq * top_k * 12 > p->slot_bytes)
        return fail(VERS_ERR_ARG, "peer_gather_merge: %u x %u entries exceed the slot of %llu bytes", nq, top_k,
                    (unsigned long long)p->slot_bytes);
    vers_ctx* ctx = p->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    const unsigned grid = (unsigned)ceil_div(nq, PG_WARPS);
    if (grid > (unsigned)ctx->sm_count * 8)  // every block must be resident while it waits for the peers
        return fail(VERS_ERR_UNSUPPORTED, "peer_gather_merge: batch of %u queries is too large for one resident grid", nq);
    p->step += 1;
    peer_gather_merge_kernel<<<grid, PG_WARPS * 32, (size_t)PG_WARPS * top_k * 12, ctx->stream>>>(
        p->d_peer_base, p->world, p->rank, p->slot_bytes, p->flags_off, p->done_off, p->step, d_local_ids,
        d_local_dists, nq, top_k, d_ids, d_dists, d_counts);
    VERS_LAUNCH_CHECK(ctx);
    return VERS_OK;
}

extern "C" int32_t vers_peer_free(vers_peer* p) {
    if (!p) return VERS_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    for (char* m : p->opened) cudaIpcCloseMemHandle(m);
    cudaFree(p->d_peer_base);
    cudaFree(p->d_buf);
    delete p;
    return VERS_OK;
}
