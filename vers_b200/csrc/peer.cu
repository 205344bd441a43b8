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

#include "peer.cuh"

struct vers_peer {
    vers_ctx* ctx = nullptr;
    uint32_t world = 0, rank = 0;
    uint64_t slot_bytes = 0;
    char* d_buf = nullptr;         // this rank's exchange buffer
    char** d_peer_base = nullptr;  // [world] device-visible base pointers (own entry = d_buf)
    std::vector<char*> opened;     // peer mappings to close
    uint32_t step = 0;             // host-managed step (the vers_comm path keeps it in device memory instead)
    size_t flags_off = 0, ctl_off = 0, total = 0;
    unsigned resident = 0;         // blocks of the exchange kernel that fit the GPU at once
};

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
    p->ctl_off = (p->flags_off + (size_t)2 * world * 4 + 255) & ~size_t(255);
    p->total = p->ctl_off + 256;
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
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    VERS_CUDA(cudaSetDevice(ctx->device));
    // every block must be resident while it waits for the peers: the grid is bounded by the occupancy of this kernel
    // and the warps loop over the queries
    const size_t smem = (size_t)PG_WARPS * top_k * 12;
    if (!p->resident) VERS_TRY(peer_resident_blocks(ctx, (size_t)PG_WARPS * VERS_MAX_TOPK * 12, &p->resident));
    const unsigned grid = std::min<unsigned>((unsigned)ceil_div(nq, PG_WARPS), p->resident);
    PeerRegion g;
    g.peer_base = p->d_peer_base;
    g.data_off = 0;
    g.slot_bytes = p->slot_bytes;
    g.flags_off = p->flags_off;
    g.ctl_off = p->ctl_off;
    g.world = p->world;
    g.rank = p->rank;
    peer_gather_merge_kernel<<<grid, PG_WARPS * 32, smem, ctx->stream>>>(g, p->step + 1, d_local_ids, d_local_dists, nq,
                                                                        top_k, d_ids, d_dists, d_counts);
    VERS_LAUNCH_CHECK(ctx);
    p->step += 1;  // only once the launch is known to be queued: a failed launch must not desynchronise the ranks
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
