// comm.cu — the multi-GPU index behind the C ABI: one process per GPU, `vers_comm` = this rank's
// membership in a group of `world` GPUs of one NVLink/NVSwitch box.
//
//   bootstrap : NCCL (ncclCommInitRank with a 128-byte unique id the HOST ships between its processes by whatever
//               transport it has — the Rust host owns that, the Python driver uses torch.distributed's store).  libnccl
//               is resolved with dlopen at the first vers_comm_create, so single-GPU users never need it.
//   k-means   : assign is local (rows independent).  update_centroids (ivfflat.rs:47-71) sums rows in GLOBAL row order:
//               reduce = CHAINED passes the running (sums, counts) from rank r-1 to rank r (ncclSend/ncclRecv), every
//               rank continues the left-to-right sum over its own rows, the last rank broadcasts: the association is
//               the reference's, so centroids/assignments stay bit-identical to the CPU reference at any GPU count.
//               reduce = ALLREDUCE is the plain ncclAllReduce of per-shard sums (fast, association != reference's).
//   build     : rows -> the rank that owns their inverted list (largest list first onto the least-loaded rank), one
//               all-to-all over NVLink (grouped ncclSend/ncclRecv), then vers_ivf_from_parts_dev on what arrived.
//   search    : per batch (a) every rank probes 1/world of the queries (centroids are replicated), (b) the probe
//               lists are all-gathered, (c) every rank scans the lists it owns, (d) the per-rank top-k are exchanged
//               and merged by (distance, id).  (b) and (d) do NOT go through NCCL: they are stores into the peers'
//               IPC-mapped exchange buffers + flags (peer.cuh), (d) fused with the merge in ONE kernel — a step has no
//               host synchronisation and no library collective in it, so it is capturable in a CUDA graph.
// Every NCCL call and every kernel is enqueued on the context's stream: ordering needs no cross-stream events.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cub/device/device_radix_sort.cuh>
#include <numeric>

#include "kmeans.cuh"
#include "peer.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

int32_t nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return VERS_OK;
    // by SONAME: a process that already loaded an NCCL (e.g. the copy bundled with torch) gets that one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return vers::fail(VERS_ERR_UNSUPPORTED, "vers_comm: libnccl.so.2 not found (%s)", dlerror());
    NcclApi a;
    a.handle = h;
    bool ok = true;
    auto sym = [&](const char* name) {
        void* p = dlsym(h, name);
        if (!p) ok = false;
        return p;
    };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.Broadcast = (decltype(a.Broadcast))sym("ncclBroadcast");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    if (!ok) return vers::fail(VERS_ERR_UNSUPPORTED, "vers_comm: libnccl.so.2 lacks a required symbol");
    g_nccl = a;
    return VERS_OK;
}

#define VERS_NCCL(expr)                                                                                        \
    do {                                                                                                       \
        ncclResult_t _r = (expr);                                                                              \
        if (_r != ncclSuccess)                                                                                 \
            return ::vers::fail(VERS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), __FILE__, \
                                __LINE__);                                                                     \
    } while (0)

}  // namespace

struct vers_comm {
    vers_ctx* ctx = nullptr;
    uint32_t world = 1, rank = 0;
    ncclComm_t nccl = nullptr;
    // exchange buffer of this rank (two regions: probe lists, top-k), IPC-mapped by every peer
    char* d_buf = nullptr;
    size_t buf_bytes = 0;
    char** d_peer_base = nullptr;
    std::vector<char*> opened;
    uint64_t probe_slot = 0, topk_slot = 0;
    vers::PeerRegion probe_rg{}, topk_rg{};
    // per-step device scratch (grow-only)
    uint64_t* d_probe_local = nullptr;
    uint64_t* d_probe_all = nullptr;
    size_t probe_cap = 0;  // bytes of d_probe_local; d_probe_all = world x that
    uint64_t* d_loc_ids = nullptr;
    float* d_loc_d = nullptr;
    uint32_t* d_loc_cnt = nullptr;
    size_t loc_cap = 0;  // entries (nq * k)
    size_t loc_q_cap = 0;
    unsigned merge_resident = 0;
    double last_exchange_s = 0.0;  // the row all-to-all of the most recent vers_sharded_ivf_build
    vers::GraphCache call_graph;   // vers_sharded_ivf_search: the device work of a repeated call shape, as a CUDA graph
};

namespace vers {
void ivf_state_stamp(const vers_ivf* ivf, uint64_t out[4]);  // ivf.cu
}

namespace vers {

__global__ void comm_gather_init_kernel(const float* __restrict__ rows, uint32_t ld, uint64_t n, uint64_t id_base,
                                        const uint64_t* __restrict__ init, uint32_t C, float* __restrict__ cents) {
    // centroid j = global row init[j] if this rank holds it, zeros otherwise: the INTEGER sum over ranks of the bit
    // patterns is then exact (and keeps -0.0, which a float add would turn into +0.0)
    const uint32_t ld4 = ld >> 2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (uint64_t)C * ld4;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(i / ld4), c = (uint32_t)(i - (uint64_t)j * ld4);
        const uint64_t g = init[j];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g >= id_base && g - id_base < n) v = reinterpret_cast<const float4*>(rows + (g - id_base) * ld)[c];
        reinterpret_cast<float4*>(cents)[i] = v;
    }
}

__global__ void comm_hist64_kernel(const uint32_t* __restrict__ assign, uint64_t n, unsigned long long* hist) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[assign[i]], 1ull);
}

__global__ void comm_dest_kernel(const uint32_t* __restrict__ assign, const uint32_t* __restrict__ owner, uint64_t n,
                                 uint32_t* __restrict__ dest, uint32_t* __restrict__ iota) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        dest[i] = owner[assign[i]];
        iota[i] = (uint32_t)i;
    }
}

// send buffers in destination order: rows, global ids, clusters
__global__ void comm_pack_kernel(const float* __restrict__ rows, uint32_t ld, const uint32_t* __restrict__ order,
                                 const uint32_t* __restrict__ assign, uint64_t n, uint64_t id_base,
                                 float* __restrict__ out_rows, uint64_t* __restrict__ out_ids,
                                 uint32_t* __restrict__ out_assign) {
    const uint32_t ld4 = ld >> 2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * ld4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / ld4;
        const uint32_t c = (uint32_t)(i - r * ld4);
        const uint32_t src = order[r];
        reinterpret_cast<float4*>(out_rows)[i] = reinterpret_cast<const float4*>(rows)[(uint64_t)src * ld4 + c];
        if (c == 0) {
            out_ids[r] = id_base + src;
            out_assign[r] = assign[src];
        }
    }
}

// owner rank of every inverted list: largest list first onto the least-loaded rank (ties: lowest rank, lowest list)
static std::vector<uint32_t> balanced_list_owners(const std::vector<unsigned long long>& sizes, uint32_t world) {
    std::vector<uint32_t> order(sizes.size());
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return sizes[a] > sizes[b]; });
    std::vector<unsigned long long> load(world, 0);
    std::vector<uint32_t> owner(sizes.size(), 0);
    for (uint32_t c : order) {
        uint32_t best = 0;
        for (uint32_t r = 1; r < world; ++r)
            if (load[r] < load[best]) best = r;
        owner[c] = best;
        load[best] += sizes[c];
    }
    return owner;
}

static int32_t comm_sync(vers_comm* cm) {
    VERS_CUDA(cudaStreamSynchronize(cm->ctx->stream));
    return VERS_OK;
}

// a stream-ordered rendezvous of all ranks (tiny all-reduce) followed by a host wait
static int32_t comm_barrier(vers_comm* cm, uint32_t* d_word) {
    if (cm->world > 1) VERS_NCCL(g_nccl.AllReduce(d_word, d_word, 1, ncclUint32, ncclSum, cm->nccl, cm->ctx->stream));
    return comm_sync(cm);
}

static void comm_close_peers(vers_comm* cm) {
    for (char* m : cm->opened) cudaIpcCloseMemHandle(m);
    cm->opened.clear();
}

// (re)creates the exchange buffer when a batch shape needs larger slots.  COLLECTIVE: every rank calls it with the same
// arguments (they derive from the batch shape, which is the same on every rank).
static int32_t comm_ensure_exchange(vers_comm* cm, uint64_t probe_bytes, uint64_t topk_bytes) {
    probe_bytes = (probe_bytes + 255) & ~uint64_t(255);
    topk_bytes = (topk_bytes + 255) & ~uint64_t(255);
    if (cm->d_buf && probe_bytes <= cm->probe_slot && topk_bytes <= cm->topk_slot) return VERS_OK;
    vers_ctx* ctx = cm->ctx;
    const uint32_t W = cm->world;
    uint32_t* d_word = nullptr;
    VERS_CUDA(cudaMalloc(&d_word, 256));
    VERS_CUDA(cudaMemsetAsync(d_word, 0, 256, ctx->stream));
    int32_t rc = comm_barrier(cm, d_word);  // nobody is still inside a step that uses the old buffer
    if (rc == VERS_OK && cm->d_buf) {
        comm_close_peers(cm);
        rc = comm_barrier(cm, d_word);  // every mapping of my old buffer is closed before I free it
        cudaFree(cm->d_buf);
        cm->d_buf = nullptr;
    }
    if (rc != VERS_OK) {
        cudaFree(d_word);
        return rc;
    }
    cm->probe_slot = std::max(cm->probe_slot, probe_bytes);
    cm->topk_slot = std::max(cm->topk_slot, topk_bytes);
    // layout: probe data | probe flags | probe ctl | topk data | topk flags | topk ctl
    auto region = [&](uint64_t& off, uint64_t slot, PeerRegion& g) {
        g.data_off = off;
        g.slot_bytes = slot;
        off += 2ull * W * slot;
        g.flags_off = off;
        off += (2ull * W * 4 + 255) & ~255ull;
        g.ctl_off = off;
        off += 256;
        g.world = W;
        g.rank = cm->rank;
    };
    uint64_t off = 0;
    region(off, cm->probe_slot, cm->probe_rg);
    region(off, cm->topk_slot, cm->topk_rg);
    cm->buf_bytes = off;
    cudaError_t e = cudaMalloc(&cm->d_buf, cm->buf_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(cm->d_buf, 0, cm->buf_bytes, ctx->stream);
    const uint32_t one = 1;  // the device-resident step counters start at 1 (flags start at 0)
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(cm->d_buf + cm->probe_rg.ctl_off + 8, &one, 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(cm->d_buf + cm->topk_rg.ctl_off + 8, &one, 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && !cm->d_peer_base) e = cudaMalloc(&cm->d_peer_base, sizeof(char*) * W);
    // swap the 64-byte CUDA IPC handles through NCCL
    cudaIpcMemHandle_t mine;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine, cm->d_buf);
    char* d_handles = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&d_handles, (size_t)64 * (W + 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_handles + 64 * W, &mine, 64, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(d_word);
        cudaFree(d_handles);
        return fail(e == cudaErrorMemoryAllocation ? VERS_ERR_NOMEM : VERS_ERR_CUDA, "vers_comm exchange buffer: %s",
                    cudaGetErrorString(e));
    }
    std::vector<char> all((size_t)64 * W);
    rc = VERS_OK;
    {
        ncclResult_t r = g_nccl.AllGather(d_handles + 64 * W, d_handles, 64, ncclChar, cm->nccl, ctx->stream);
        if (r != ncclSuccess) rc = fail(VERS_ERR_CUDA, "ncclAllGather(ipc handles): %s", g_nccl.GetErrorString(r));
    }
    if (rc == VERS_OK) {
        e = cudaMemcpyAsync(all.data(), d_handles, all.size(), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "vers_comm exchange buffer: %s", cudaGetErrorString(e));
    }
    std::vector<char*> base(W, nullptr);
    for (uint32_t r = 0; r < W && rc == VERS_OK; ++r) {
        if (r == cm->rank) {
            base[r] = cm->d_buf;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, all.data() + (size_t)r * 64, 64);
        void* ptr = nullptr;
        e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "cudaIpcOpenMemHandle(rank %u): %s", r, cudaGetErrorString(e));
        base[r] = (char*)ptr;
        if (ptr) cm->opened.push_back((char*)ptr);
    }
    if (rc == VERS_OK) {
        e = cudaMemcpyAsync(cm->d_peer_base, base.data(), sizeof(char*) * W, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "vers_comm exchange buffer: %s", cudaGetErrorString(e));
    }
    cm->probe_rg.peer_base = cm->d_peer_base;
    cm->topk_rg.peer_base = cm->d_peer_base;
    if (rc == VERS_OK) rc = comm_barrier(cm, d_word);  // every rank has mapped every buffer before anyone stores
    cudaFree(d_word);
    cudaFree(d_handles);
    return rc;
}

static int32_t comm_step_buffers(vers_comm* cm, size_t probe_bytes, uint32_t nq, uint32_t k) {
    if (probe_bytes > cm->probe_cap) {
        VERS_CUDA(cudaStreamSynchronize(cm->ctx->stream));
        cudaFree(cm->d_probe_local);
        cudaFree(cm->d_probe_all);
        cm->d_probe_local = cm->d_probe_all = nullptr;
        cm->probe_cap = 0;
        VERS_CUDA(cudaMalloc(&cm->d_probe_local, probe_bytes));
        VERS_CUDA(cudaMalloc(&cm->d_probe_all, probe_bytes * cm->world));
        cm->probe_cap = probe_bytes;
    }
    const size_t nk = (size_t)nq * k;
    if (nk > cm->loc_cap || nq > cm->loc_q_cap) {
        VERS_CUDA(cudaStreamSynchronize(cm->ctx->stream));
        cudaFree(cm->d_loc_ids);
        cudaFree(cm->d_loc_d);
        cudaFree(cm->d_loc_cnt);
        cm->d_loc_ids = nullptr, cm->d_loc_d = nullptr, cm->d_loc_cnt = nullptr;
        cm->loc_cap = cm->loc_q_cap = 0;
        VERS_CUDA(cudaMalloc(&cm->d_loc_ids, nk * 8));
        VERS_CUDA(cudaMalloc(&cm->d_loc_d, nk * 4));
        VERS_CUDA(cudaMalloc(&cm->d_loc_cnt, (size_t)nq * 4));
        cm->loc_cap = nk;
        cm->loc_q_cap = nq;
    }
    return VERS_OK;
}

static int32_t sharded_search_dev(vers_comm* cm, vers_ivf* ivf, const float* d_queries, uint32_t nq, uint32_t top_k,
                                  uint32_t nprobe, uint64_t* d_ids, float* d_d, uint32_t* d_cnt) {
    vers_ctx* ctx = cm->ctx;
    if (cm->world == 1) return vers_ivf_search_dev(ivf, d_queries, nq, top_k, nprobe, d_ids, d_d, d_cnt);
    if (nprobe == 0)
        return fail(VERS_ERR_UNSUPPORTED, "sharded search needs nprobe >= 1 (the reference's spill semantics walk the "
                                          "lists of ONE index in order, ivfflat.rs:163-197)");
    uint64_t n = 0;
    uint32_t dim = 0, C = 0;
    VERS_TRY(vers_ivf_info(ivf, &n, &dim, &C, nullptr, nullptr));
    const uint32_t ld = round_up(dim, 4), np = std::min(nprobe, C);
    if (np > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "nprobe %u > %u", np, VERS_MAX_TOPK);
    if (top_k > VERS_MAX_TOPK) return fail(VERS_ERR_UNSUPPORTED, "top_k %u > %u", top_k, VERS_MAX_TOPK);
    const uint32_t W = cm->world, per = (nq + W - 1) / W;
    const uint32_t q0 = std::min(nq, cm->rank * per), nql = std::min(nq, q0 + per) - q0;
    const uint64_t probe_bytes = ((uint64_t)per * np * 8 + 15) & ~15ull;
    VERS_TRY(comm_ensure_exchange(cm, probe_bytes, (uint64_t)nq * top_k * 12));
    VERS_TRY(comm_step_buffers(cm, probe_bytes, nq, top_k));
    // (a) my share of the centroid probe (exact order, ivfflat.rs:155-161); unused tail entries = u64::MAX
    if (nql < per) VERS_CUDA(cudaMemsetAsync(cm->d_probe_local, 0xff, probe_bytes, ctx->stream));
    if (nql) VERS_TRY(vers_ivf_probe_dev(ivf, d_queries + (size_t)q0 * ld, nql, np, cm->d_probe_local));
    {   // (b) all-gather of the probe lists over peer memory
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        const unsigned grid = (unsigned)std::min<uint64_t>(ceil_div(probe_bytes >> 4, 256), (uint64_t)ctx->sm_count);
        peer_publish_kernel<<<grid, 256, 0, ctx->stream>>>(cm->probe_rg, reinterpret_cast<const uint4*>(cm->d_probe_local),
                                                          probe_bytes, 0);
        VERS_LAUNCH_CHECK(ctx);
        const uint64_t elems = (uint64_t)per * np;  // u64 probe ids per rank: d_probe_all is the compact [world * per][np]
        const unsigned grid2 = (unsigned)std::min<uint64_t>(ceil_div(elems * W, 256), (uint64_t)ctx->sm_count);
        peer_wait_copy_kernel<<<grid2, 256, 0, ctx->stream>>>(cm->probe_rg,
                                                             reinterpret_cast<unsigned long long*>(cm->d_probe_all), elems, 0);
        VERS_LAUNCH_CHECK(ctx);
    }
    // (c) scan the lists this rank owns.  d_probe_all is [world * per][np]: query q = r * per + j is row q (rows past nq
    // are the unused tail of the last ranks)
    VERS_TRY(vers_ivf_search_probed_dev(ivf, d_queries, nq, top_k, np, cm->d_probe_all, cm->d_loc_ids, cm->d_loc_d,
                                        cm->d_loc_cnt));
    {   // (d) exchange + merge of the per-rank top-k, one kernel
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        const size_t smem = (size_t)PG_WARPS * top_k * 12;
        if (!cm->merge_resident) VERS_TRY(peer_resident_blocks(ctx, (size_t)PG_WARPS * VERS_MAX_TOPK * 12, &cm->merge_resident));
        const unsigned grid = std::min<unsigned>((unsigned)ceil_div(nq, PG_WARPS), cm->merge_resident);
        peer_gather_merge_kernel<<<grid, PG_WARPS * 32, smem, ctx->stream>>>(cm->topk_rg, 0, cm->d_loc_ids, cm->d_loc_d, nq,
                                                                            top_k, d_ids, d_d, d_cnt);
        VERS_LAUNCH_CHECK(ctx);
    }
    return VERS_OK;
}

}  // namespace vers

using namespace vers;

extern "C" int32_t vers_comm_unique_id(uint8_t id_out[128]) {
    if (!id_out) return fail(VERS_ERR_ARG, "comm_unique_id: null");
    VERS_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    VERS_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, 128);
    return VERS_OK;
}

extern "C" int32_t vers_comm_create(vers_ctx* ctx, uint32_t world, uint32_t rank, const uint8_t unique_id[128],
                                    vers_comm** out) {
    if (!ctx || !out || world == 0 || rank >= world || (world > 1 && !unique_id))
        return fail(VERS_ERR_ARG, "comm_create: bad argument");
    *out = nullptr;
    VERS_CUDA(cudaSetDevice(ctx->device));
    vers_comm* cm = new vers_comm();
    cm->ctx = ctx;
    cm->world = world;
    cm->rank = rank;
    if (world > 1) {
        int32_t rc = nccl_load();
        if (rc == VERS_OK) {
            ncclUniqueId id;
            memcpy(&id, unique_id, 128);
            ncclResult_t r = g_nccl.CommInitRank(&cm->nccl, (int)world, id, (int)rank);
            if (r != ncclSuccess) rc = fail(VERS_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
        }
        // Establish every connection the build and the search will use NOW (NCCL connects lazily, on the first
        // collective / the first send-recv of a pair: seconds at 8 ranks), so that index builds are not charged for it.
        uint32_t* d_w = nullptr;
        if (rc == VERS_OK && cudaMalloc(&d_w, 256 * (size_t)(world + 1)) != cudaSuccess)
            rc = fail(VERS_ERR_NOMEM, "comm_create: cudaMalloc");
        if (rc == VERS_OK) {
            auto run = [&]() -> int32_t {
                VERS_CUDA(cudaMemsetAsync(d_w, 0, 256 * (size_t)(world + 1), ctx->stream));
                VERS_NCCL(g_nccl.AllReduce(d_w, d_w, 1, ncclUint32, ncclSum, cm->nccl, ctx->stream));
                VERS_NCCL(g_nccl.Broadcast(d_w, d_w, 1, ncclUint32, (int)world - 1, cm->nccl, ctx->stream));
                VERS_NCCL(g_nccl.GroupStart());
                for (uint32_t r = 0; r < world; ++r) {
                    if (r == rank) continue;
                    VERS_NCCL(g_nccl.Send(d_w + 64 * world, 1, ncclUint32, (int)r, cm->nccl, ctx->stream));
                    VERS_NCCL(g_nccl.Recv(d_w + 64 * r, 1, ncclUint32, (int)r, cm->nccl, ctx->stream));
                }
                VERS_NCCL(g_nccl.GroupEnd());
                VERS_CUDA(cudaStreamSynchronize(ctx->stream));
                return VERS_OK;
            };
            rc = run();
        }
        cudaFree(d_w);
        if (rc == VERS_OK) rc = comm_ensure_exchange(cm, 1000 * 32 * 8 / world + 4096, 1000 * 10 * 12);
        if (rc != VERS_OK) {
            if (cm->nccl) g_nccl.CommDestroy(cm->nccl);
            delete cm;
            return rc;
        }
    }
    *out = cm;
    return VERS_OK;
}

extern "C" int32_t vers_comm_destroy(vers_comm* cm) {
    if (!cm) return VERS_OK;
    cudaSetDevice(cm->ctx->device);
    cudaStreamSynchronize(cm->ctx->stream);
    cm->call_graph.reset();
    comm_close_peers(cm);
    cudaFree(cm->d_peer_base);
    cudaFree(cm->d_buf);
    cudaFree(cm->d_probe_local);
    cudaFree(cm->d_probe_all);
    cudaFree(cm->d_loc_ids);
    cudaFree(cm->d_loc_d);
    cudaFree(cm->d_loc_cnt);
    if (cm->nccl) g_nccl.CommDestroy(cm->nccl);
    delete cm;
    return VERS_OK;
}

extern "C" int32_t vers_comm_info(const vers_comm* cm, uint32_t* world, uint32_t* rank, double* last_exchange_seconds) {
    if (!cm) return fail(VERS_ERR_ARG, "comm_info: null");
    if (world) *world = cm->world;
    if (rank) *rank = cm->rank;
    if (last_exchange_seconds) *last_exchange_seconds = cm->last_exchange_s;
    return VERS_OK;
}

extern "C" int32_t vers_debug_peer_times(vers_comm* cm, uint64_t out_ns[4]) {
    if (!cm || !out_ns) return fail(VERS_ERR_ARG, "debug_peer_times: null");
    VERS_CUDA(cudaSetDevice(cm->ctx->device));
    VERS_CUDA(cudaStreamSynchronize(cm->ctx->stream));
    VERS_CUDA(cudaMemcpyFromSymbol(out_ns, g_peer_dbg, 32));
    return VERS_OK;
}

extern "C" int32_t vers_comm_barrier(vers_comm* cm) {
    if (!cm) return fail(VERS_ERR_ARG, "comm_barrier: null");
    VERS_CUDA(cudaSetDevice(cm->ctx->device));
    if (cm->world == 1) return comm_sync(cm);
    uint32_t* d_word = nullptr;
    VERS_CUDA(cudaMalloc(&d_word, 256));
    cudaMemsetAsync(d_word, 0, 256, cm->ctx->stream);
    int32_t rc = comm_barrier(cm, d_word);
    cudaFree(d_word);
    return rc;
}

// max over ranks of a host value (bench: device-timed milliseconds, max over ranks)
extern "C" int32_t vers_comm_max_f64(vers_comm* cm, double* value_io) {
    if (!cm || !value_io) return fail(VERS_ERR_ARG, "comm_max_f64: null");
    if (cm->world == 1) return VERS_OK;
    VERS_CUDA(cudaSetDevice(cm->ctx->device));
    double* d = nullptr;
    VERS_CUDA(cudaMalloc(&d, 256));
    int32_t rc = VERS_OK;
    cudaError_t e = cudaMemcpyAsync(d, value_io, 8, cudaMemcpyHostToDevice, cm->ctx->stream);
    if (e == cudaSuccess) {
        ncclResult_t r = g_nccl.AllReduce(d, d, 1, ncclFloat64, ncclMax, cm->nccl, cm->ctx->stream);
        if (r != ncclSuccess) rc = fail(VERS_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    }
    if (rc == VERS_OK && e == cudaSuccess) e = cudaMemcpyAsync(value_io, d, 8, cudaMemcpyDeviceToHost, cm->ctx->stream);
    if (rc == VERS_OK && e == cudaSuccess) e = cudaStreamSynchronize(cm->ctx->stream);
    cudaFree(d);
    if (rc == VERS_OK && e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "comm_max_f64: %s", cudaGetErrorString(e));
    return rc;
}

// the list -> owner table of a list-sharded index (pure host arithmetic, exposed so that hosts and tests can reproduce
// which rank holds which list)
extern "C" int32_t vers_sharded_list_owners(const uint64_t* list_sizes, uint32_t num_clusters, uint32_t world,
                                            uint32_t* owner_out) {
    if (!list_sizes || !owner_out || world == 0) return fail(VERS_ERR_ARG, "sharded_list_owners: bad argument");
    std::vector<unsigned long long> sizes(list_sizes, list_sizes + num_clusters);
    const std::vector<uint32_t> owner = balanced_list_owners(sizes, world);
    std::copy(owner.begin(), owner.end(), owner_out);
    return VERS_OK;
}

// ---------------------------------------------------------------------------------------------- k-means over row shards
extern "C" int32_t vers_sharded_kmeans_fit(vers_comm* cm, vers_kmeans* km, const uint64_t* init_rows_global,
                                           uint32_t max_iterations, int32_t reduce, uint32_t* iterations_run) {
    if (!cm || !km || !init_rows_global) return fail(VERS_ERR_ARG, "sharded_kmeans_fit: null argument");
    if (reduce != VERS_REDUCE_CHAINED && reduce != VERS_REDUCE_ALLREDUCE)
        return fail(VERS_ERR_ARG, "sharded_kmeans_fit: unknown reduce mode %d", reduce);
    vers_dataset* ds = km->ds;
    vers_ctx* ctx = ds->ctx;
    if (ctx != cm->ctx) return fail(VERS_ERR_ARG, "sharded_kmeans_fit: the k-means state lives on another context");
    VERS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t C = km->C, ld = ds->ld, W = cm->world, rank = cm->rank;
    const size_t cl = (size_t)C * ld;
    {   // initialize_centroids (ivfflat.rs:18-27) with injected GLOBAL row numbers: the owner of a row contributes it
        uint64_t* d_init = nullptr;
        VERS_CUDA(cudaMalloc(&d_init, (size_t)C * 8));
        cudaError_t e = cudaMemcpyAsync(d_init, init_rows_global, (size_t)C * 8, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) {
            std::lock_guard<std::recursive_mutex> lk(ctx->mu);
            comm_gather_init_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(ds->d_rows, ld, ds->n, ds->id_base, d_init, C,
                                                                     km->d_cents);
            ctx->launches += 1;
            e = cudaGetLastError();
        }
        int32_t rc = e == cudaSuccess ? VERS_OK : fail(VERS_ERR_CUDA, "sharded_kmeans_fit: %s", cudaGetErrorString(e));
        if (rc == VERS_OK && W > 1) {
            ncclResult_t r = g_nccl.AllReduce(km->d_cents, km->d_cents, cl, ncclInt32, ncclSum, cm->nccl, s);
            if (r != ncclSuccess) rc = fail(VERS_ERR_CUDA, "ncclAllReduce(init centroids): %s", g_nccl.GetErrorString(r));
        }
        if (rc == VERS_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = fail(VERS_ERR_CUDA, "sharded_kmeans_fit: sync");
        cudaFree(d_init);
        VERS_TRY(rc);
    }
    uint32_t it = 0;
    while (it < max_iterations) {
        VERS_TRY(vers_kmeans_assign_step(km));
        if (W == 1 || reduce == VERS_REDUCE_CHAINED) {
            if (rank == 0) {
                VERS_CUDA(cudaMemsetAsync(km->d_sums, 0, cl * 4, s));
                VERS_CUDA(cudaMemsetAsync(km->d_counts, 0, (size_t)C * 8, s));
            } else {
                VERS_NCCL(g_nccl.GroupStart());
                VERS_NCCL(g_nccl.Recv(km->d_sums, cl, ncclFloat32, (int)rank - 1, cm->nccl, s));
                VERS_NCCL(g_nccl.Recv(km->d_counts, C, ncclUint64, (int)rank - 1, cm->nccl, s));
                VERS_NCCL(g_nccl.GroupEnd());
            }
            VERS_TRY(vers_kmeans_sums_step_dev(km, km->d_sums, km->d_counts));
            if (W > 1) {
                if (rank + 1 < W) {
                    VERS_NCCL(g_nccl.GroupStart());
                    VERS_NCCL(g_nccl.Send(km->d_sums, cl, ncclFloat32, (int)rank + 1, cm->nccl, s));
                    VERS_NCCL(g_nccl.Send(km->d_counts, C, ncclUint64, (int)rank + 1, cm->nccl, s));
                    VERS_NCCL(g_nccl.GroupEnd());
                }
                VERS_NCCL(g_nccl.Broadcast(km->d_sums, km->d_sums, cl, ncclFloat32, (int)W - 1, cm->nccl, s));
                VERS_NCCL(g_nccl.Broadcast(km->d_counts, km->d_counts, C, ncclUint64, (int)W - 1, cm->nccl, s));
            }
        } else {
            VERS_CUDA(cudaMemsetAsync(km->d_sums, 0, cl * 4, s));
            VERS_CUDA(cudaMemsetAsync(km->d_counts, 0, (size_t)C * 8, s));
            VERS_TRY(vers_kmeans_sums_step_dev(km, km->d_sums, km->d_counts));
            VERS_NCCL(g_nccl.AllReduce(km->d_sums, km->d_sums, cl, ncclFloat32, ncclSum, cm->nccl, s));
            VERS_NCCL(g_nccl.AllReduce(km->d_counts, km->d_counts, C, ncclUint64, ncclSum, cm->nccl, s));
        }
        uint32_t changed = 0;
        VERS_TRY(vers_kmeans_finalize_step_dev(km, km->d_sums, km->d_counts, &changed));
        it += 1;
        if (!changed) break;  // identical sums on every rank => the same decision everywhere
    }
    VERS_TRY(vers_kmeans_assign_step(km));  // the final assign of build_kmeans (ivfflat.rs:96-98)
    if (iterations_run) *iterations_run = it;
    return VERS_OK;
}

extern "C" int32_t vers_sharded_kmeans_cost(vers_comm* cm, vers_kmeans* km, float* cost) {
    if (!cm || !km || !cost) return fail(VERS_ERR_ARG, "sharded_kmeans_cost: null argument");
    vers_ctx* ctx = km->ds->ctx;
    if (ctx != cm->ctx) return fail(VERS_ERR_ARG, "sharded_kmeans_cost: the k-means state lives on another context");
    VERS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t W = cm->world, rank = cm->rank;
    float acc = 0.0f;  // calculate_kmeans_cost folds in global row order (ivfflat.rs:138-149): rank r continues r-1's value
    float* d_acc = nullptr;
    if (W > 1) {
        VERS_CUDA(cudaMalloc(&d_acc, 256));
        if (rank > 0) {
            VERS_NCCL(g_nccl.Recv(d_acc, 1, ncclFloat32, (int)rank - 1, cm->nccl, s));
            VERS_CUDA(cudaMemcpyAsync(&acc, d_acc, 4, cudaMemcpyDeviceToHost, s));
            VERS_CUDA(cudaStreamSynchronize(s));
        }
    }
    int32_t rc = vers_kmeans_cost_step(km, &acc);
    if (rc == VERS_OK && W > 1) {
        auto tail = [&]() -> int32_t {
            VERS_CUDA(cudaMemcpyAsync(d_acc, &acc, 4, cudaMemcpyHostToDevice, s));
            if (rank + 1 < W) VERS_NCCL(g_nccl.Send(d_acc, 1, ncclFloat32, (int)rank + 1, cm->nccl, s));
            VERS_NCCL(g_nccl.Broadcast(d_acc, d_acc, 1, ncclFloat32, (int)W - 1, cm->nccl, s));
            VERS_CUDA(cudaMemcpyAsync(&acc, d_acc, 4, cudaMemcpyDeviceToHost, s));
            VERS_CUDA(cudaStreamSynchronize(s));
            return VERS_OK;
        };
        rc = tail();
    }
    cudaFree(d_acc);
    if (rc == VERS_OK) *cost = acc;
    return rc;
}

// ---------------------------------------------------------------------------------------------- list-sharded index build
extern "C" int32_t vers_sharded_ivf_build(vers_comm* cm, vers_kmeans* km, vers_ivf** out) {
    if (!cm || !km || !out) return fail(VERS_ERR_ARG, "sharded_ivf_build: null argument");
    *out = nullptr;
    if (cm->world == 1) return vers_ivf_from_kmeans(km, out);
    vers_dataset* ds = km->ds;
    vers_ctx* ctx = ds->ctx;
    if (ctx != cm->ctx) return fail(VERS_ERR_ARG, "sharded_ivf_build: the k-means state lives on another context");
    VERS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t C = km->C, ld = ds->ld, W = cm->world, rank = cm->rank;
    const uint64_t n = ds->n;
    if (n >= 0xffffffffull) return fail(VERS_ERR_UNSUPPORTED, "more than 2^32-2 rows per GPU shard");

    struct Bufs {
        unsigned long long *hist = nullptr, *counts_all = nullptr;
        uint32_t *owner = nullptr, *dest = nullptr, *dest_sorted = nullptr, *iota = nullptr, *order = nullptr;
        void* cub = nullptr;
        float *send_rows = nullptr, *recv_rows = nullptr;
        uint64_t *send_ids = nullptr, *recv_ids = nullptr;
        uint32_t *send_assign = nullptr, *recv_assign = nullptr;
        ~Bufs() {
            cudaFree(hist), cudaFree(counts_all), cudaFree(owner), cudaFree(dest), cudaFree(dest_sorted), cudaFree(iota);
            cudaFree(order), cudaFree(cub), cudaFree(send_rows), cudaFree(recv_rows), cudaFree(send_ids);
            cudaFree(recv_ids), cudaFree(send_assign), cudaFree(recv_assign);
        }
    } b;
    const size_t n1 = n ? n : 1;
    // 1. list sizes: local histogram -> global sizes (all-reduce) -> owner table (same on every rank)
    VERS_CUDA(cudaMalloc(&b.hist, (size_t)C * 8 * 2));
    VERS_CUDA(cudaMemsetAsync(b.hist, 0, (size_t)C * 8 * 2, s));
    if (n) {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        comm_hist64_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(km->d_assign, n, b.hist);
        VERS_LAUNCH_CHECK(ctx);
    }
    VERS_NCCL(g_nccl.AllReduce(b.hist, b.hist + C, C, ncclUint64, ncclSum, cm->nccl, s));
    std::vector<unsigned long long> h_hist((size_t)C * 2);
    VERS_CUDA(cudaMemcpyAsync(h_hist.data(), b.hist, (size_t)C * 16, cudaMemcpyDeviceToHost, s));
    VERS_CUDA(cudaStreamSynchronize(s));
    std::vector<unsigned long long> sizes(h_hist.begin() + C, h_hist.end());
    const std::vector<uint32_t> owner = balanced_list_owners(sizes, W);
    std::vector<unsigned long long> send_cnt(W, 0);
    for (uint32_t c = 0; c < C; ++c) send_cnt[owner[c]] += h_hist[c];
    // 2. counts matrix [src][dst] (all-gather of every rank's send counts)
    VERS_CUDA(cudaMalloc(&b.counts_all, (size_t)W * (W + 1) * 8));
    VERS_CUDA(cudaMemcpyAsync(b.counts_all + (size_t)W * W, send_cnt.data(), (size_t)W * 8, cudaMemcpyHostToDevice, s));
    VERS_NCCL(g_nccl.AllGather(b.counts_all + (size_t)W * W, b.counts_all, W, ncclUint64, cm->nccl, s));
    std::vector<unsigned long long> mat((size_t)W * W);
    VERS_CUDA(cudaMemcpyAsync(mat.data(), b.counts_all, (size_t)W * W * 8, cudaMemcpyDeviceToHost, s));
    // 3. rows in destination order (stable: ascending local row = ascending id inside every destination block)
    VERS_CUDA(cudaMalloc(&b.owner, (size_t)C * 4));
    VERS_CUDA(cudaMemcpyAsync(b.owner, owner.data(), (size_t)C * 4, cudaMemcpyHostToDevice, s));
    VERS_CUDA(cudaMalloc(&b.dest, n1 * 4));
    VERS_CUDA(cudaMalloc(&b.dest_sorted, n1 * 4));
    VERS_CUDA(cudaMalloc(&b.iota, n1 * 4));
    VERS_CUDA(cudaMalloc(&b.order, n1 * 4));
    if (n) {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        comm_dest_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(km->d_assign, b.owner, n, b.dest, b.iota);
        VERS_LAUNCH_CHECK(ctx);
        int end_bit = 1;
        while ((1u << end_bit) < W) ++end_bit;
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, b.dest, b.dest_sorted, b.iota, b.order, (int64_t)n, 0, end_bit, s);
        VERS_CUDA(cudaMalloc(&b.cub, need));
        VERS_CUDA(cub::DeviceRadixSort::SortPairs(b.cub, need, b.dest, b.dest_sorted, b.iota, b.order, (int64_t)n, 0,
                                                  end_bit, s));
        ctx->launches += 1;
    }
    VERS_CUDA(cudaStreamSynchronize(s));
    std::vector<uint64_t> soff(W + 1, 0), roff(W + 1, 0);
    for (uint32_t r = 0; r < W; ++r) {
        soff[r + 1] = soff[r] + mat[(size_t)rank * W + r];
        roff[r + 1] = roff[r] + mat[(size_t)r * W + rank];
    }
    if (soff[W] != n) return fail(VERS_ERR_CUDA, "sharded_ivf_build: send counts do not add up");
    const uint64_t n_recv = roff[W];
    if (n_recv >= 0xffffffffull) return fail(VERS_ERR_UNSUPPORTED, "more than 2^32-2 rows per GPU shard");
    const size_t nr1 = n_recv ? n_recv : 1;
    VERS_CUDA(cudaMalloc(&b.send_rows, n1 * ld * 4));
    VERS_CUDA(cudaMalloc(&b.send_ids, n1 * 8));
    VERS_CUDA(cudaMalloc(&b.send_assign, n1 * 4));
    VERS_CUDA(cudaMalloc(&b.recv_rows, nr1 * ld * 4));
    VERS_CUDA(cudaMalloc(&b.recv_ids, nr1 * 8));
    VERS_CUDA(cudaMalloc(&b.recv_assign, nr1 * 4));
    if (n) {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        comm_pack_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(ds->d_rows, ld, b.order, km->d_assign, n, ds->id_base,
                                                          b.send_rows, b.send_ids, b.send_assign);
        VERS_LAUNCH_CHECK(ctx);
    }
    // 4. the all-to-all (blocks arrive in rank order; every rank holds an ascending block of global ids, so the rows
    //    of a list end up in ascending id order like ids[c], ivfflat.rs:123-127)
    VERS_CUDA(cudaStreamSynchronize(s));
    const auto t0 = std::chrono::steady_clock::now();
    VERS_NCCL(g_nccl.GroupStart());
    for (uint32_t r = 0; r < W; ++r) {
        const uint64_t ns = soff[r + 1] - soff[r], nr = roff[r + 1] - roff[r];
        if (r == rank) {
            if (ns != nr) return fail(VERS_ERR_CUDA, "sharded_ivf_build: self block size mismatch");
            continue;
        }
        if (ns) {
            VERS_NCCL(g_nccl.Send(b.send_rows + soff[r] * ld, ns * ld, ncclFloat32, (int)r, cm->nccl, s));
            VERS_NCCL(g_nccl.Send(b.send_ids + soff[r], ns, ncclUint64, (int)r, cm->nccl, s));
            VERS_NCCL(g_nccl.Send(b.send_assign + soff[r], ns, ncclUint32, (int)r, cm->nccl, s));
        }
        if (nr) {
            VERS_NCCL(g_nccl.Recv(b.recv_rows + roff[r] * ld, nr * ld, ncclFloat32, (int)r, cm->nccl, s));
            VERS_NCCL(g_nccl.Recv(b.recv_ids + roff[r], nr, ncclUint64, (int)r, cm->nccl, s));
            VERS_NCCL(g_nccl.Recv(b.recv_assign + roff[r], nr, ncclUint32, (int)r, cm->nccl, s));
        }
    }
    VERS_NCCL(g_nccl.GroupEnd());
    if (const uint64_t ns = soff[rank + 1] - soff[rank]) {
        VERS_CUDA(cudaMemcpyAsync(b.recv_rows + roff[rank] * ld, b.send_rows + soff[rank] * ld, ns * ld * 4,
                                  cudaMemcpyDeviceToDevice, s));
        VERS_CUDA(cudaMemcpyAsync(b.recv_ids + roff[rank], b.send_ids + soff[rank], ns * 8, cudaMemcpyDeviceToDevice, s));
        VERS_CUDA(cudaMemcpyAsync(b.recv_assign + roff[rank], b.send_assign + soff[rank], ns * 4, cudaMemcpyDeviceToDevice, s));
    }
    VERS_CUDA(cudaStreamSynchronize(s));
    cm->last_exchange_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cudaFree(b.send_rows), b.send_rows = nullptr;
    cudaFree(b.send_ids), b.send_ids = nullptr;
    cudaFree(b.send_assign), b.send_assign = nullptr;
    cudaFree(b.order), b.order = nullptr;
    cudaFree(b.iota), b.iota = nullptr;
    cudaFree(b.dest), b.dest = nullptr;
    cudaFree(b.dest_sorted), b.dest_sorted = nullptr;
    // 5. this rank's lists from what arrived (lists it does not own are empty here; the centroid table is whole)
    vers_dataset* local = nullptr;
    VERS_TRY(vers_dataset_wrap_device(ctx, b.recv_rows, n_recv, ds->dim, 0, &local));
    int32_t rc = vers_ivf_from_parts_dev(local, km->d_cents, C, b.recv_assign, b.recv_ids, out);
    if (rc == VERS_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = fail(VERS_ERR_CUDA, "sharded_ivf_build: sync");
    vers_dataset_free(local);
    return rc;
}

// ---------------------------------------------------------------------------------------------- sharded search
extern "C" int32_t vers_sharded_ivf_search_dev(vers_comm* cm, vers_ivf* ivf, const float* d_queries, uint32_t nq,
                                               uint32_t top_k, uint32_t nprobe, uint64_t* d_ids, float* d_dists,
                                               uint32_t* d_counts) {
    if (!cm || !ivf || (!d_queries && nq) || !d_ids || !d_dists) return fail(VERS_ERR_ARG, "sharded_ivf_search_dev: null argument");
    if (nq == 0 || top_k == 0) return VERS_OK;
    VERS_CUDA(cudaSetDevice(cm->ctx->device));
    return sharded_search_dev(cm, ivf, d_queries, nq, top_k, nprobe, d_ids, d_dists, d_counts);
}

extern "C" int32_t vers_sharded_ivf_search(vers_comm* cm, vers_ivf* ivf, const float* queries, uint32_t nq,
                                           uint32_t q_stride_floats, uint32_t top_k, uint32_t nprobe, uint64_t* ids,
                                           float* dists, uint32_t* counts) {
    if (!cm || !ivf || (!queries && nq) || (!ids && nq && top_k) || (!dists && nq && top_k))
        return fail(VERS_ERR_ARG, "sharded_ivf_search: null argument");
    if (cm->world == 1) return vers_ivf_search(ivf, queries, nq, q_stride_floats, top_k, nprobe, ids, dists, counts);
    if (nq == 0) return VERS_OK;
    if (top_k == 0) {
        if (counts) memset(counts, 0, sizeof(uint32_t) * nq);
        return VERS_OK;
    }
    uint64_t n = 0;
    uint32_t dim = 0, C = 0;
    VERS_TRY(vers_ivf_info(ivf, &n, &dim, &C, nullptr, nullptr));
    if (q_stride_floats < dim) return fail(VERS_ERR_ARG, "sharded_ivf_search: query stride < dim");
    const uint32_t ld = round_up(dim, 4);
    vers_ctx* ctx = cm->ctx;
    VERS_CUDA(cudaSetDevice(ctx->device));
    const size_t nk = (size_t)nq * top_k;
    float* d_q;
    uint64_t* d_ids;
    float* d_d;
    uint32_t* d_c;
    {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        ScratchCarver plan(nullptr);
        plan.plan<float>((size_t)nq * ld);
        plan.plan<uint64_t>(nk);
        plan.plan<float>(nk);
        plan.plan<uint32_t>(nq);
        VERS_TRY(io_reserve(ctx, plan.off + 256));
        ScratchCarver io(ctx->io);
        d_q = io.take<float>((size_t)nq * ld);
        d_ids = io.take<uint64_t>(nk);
        d_d = io.take<float>(nk);
        d_c = io.take<uint32_t>(nq);
        if (q_stride_floats == ld && ld == dim) {
            VERS_CUDA(cudaMemcpyAsync(d_q, queries, (size_t)nq * ld * 4, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            if (ld != dim) VERS_CUDA(cudaMemsetAsync(d_q, 0, (size_t)nq * ld * 4, ctx->stream));
            VERS_CUDA(cudaMemcpy2DAsync(d_q, (size_t)ld * 4, queries, (size_t)q_stride_floats * 4, (size_t)dim * 4, nq,
                                        cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    {   // every rank makes the same sequence of calls, so every rank takes the same eager / capture / replay decision
        uint64_t key[12] = {reinterpret_cast<uint64_t>(ivf), nq, top_k, nprobe, 0, 0, 0, 0,
                            reinterpret_cast<uint64_t>(ctx->scratch), reinterpret_cast<uint64_t>(d_q),
                            reinterpret_cast<uint64_t>(ctx->stream),
                            ctx->scratch_bytes ^ (reinterpret_cast<uint64_t>(cm->d_probe_all) << 1) ^
                                (reinterpret_cast<uint64_t>(cm->d_loc_ids) << 2)};
        ivf_state_stamp(ivf, key + 4);
        // the context mutex (recursive) is held across the whole call: while a capture swaps ctx->stream no other thread
        // may enqueue work through this context
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        VERS_TRY(graph_cached_run(ctx, cm->call_graph, key, [&]() {
            return sharded_search_dev(cm, ivf, d_q, nq, top_k, nprobe, d_ids, d_d, d_c);
        }));
    }
    VERS_CUDA(cudaMemcpyAsync(ids, d_ids, nk * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaMemcpyAsync(dists, d_d, nk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) VERS_CUDA(cudaMemcpyAsync(counts, d_c, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VERS_CUDA(cudaStreamSynchronize(ctx->stream));
    return VERS_OK;
}

// ---------------------------------------------------------------------------------------------- forest search, queries sharded
extern "C" int32_t vers_sharded_lsh_search(vers_comm* cm, vers_lsh* lsh, const float* queries, uint32_t nq,
                                           uint32_t q_stride_floats, uint32_t top_k, uint64_t* ids, float* dists,
                                           uint32_t* counts) {
    if (!cm || !lsh || (!queries && nq) || (!ids && nq && top_k) || (!dists && nq && top_k))
        return fail(VERS_ERR_ARG, "sharded_lsh_search: null argument");
    if (cm->world == 1) return vers_lsh_search(lsh, queries, nq, q_stride_floats, top_k, ids, dists, counts);
    if (nq == 0) return VERS_OK;
    if (top_k == 0) {
        if (counts) memset(counts, 0, sizeof(uint32_t) * nq);
        return VERS_OK;
    }
    vers_ctx* ctx = cm->ctx;
    VERS_CUDA(cudaSetDevice(ctx->device));
    const uint32_t W = cm->world, per = (nq + W - 1) / W;
    const uint32_t q0 = std::min(nq, cm->rank * per), nql = std::min(nq, q0 + per) - q0;
    // my slice on my replica (host buffers in and out), packed as [per*k ids | per*k dists | per counts]
    const size_t nk = (size_t)per * top_k, slot = (nk * 12 + (size_t)per * 4 + 15) & ~size_t(15);
    std::vector<unsigned char> mine(slot, 0), all((size_t)slot * W);
    uint64_t* m_ids = reinterpret_cast<uint64_t*>(mine.data());
    float* m_d = reinterpret_cast<float*>(mine.data() + nk * 8);
    uint32_t* m_c = reinterpret_cast<uint32_t*>(mine.data() + nk * 12);
    if (nql) VERS_TRY(vers_lsh_search(lsh, queries + (size_t)q0 * q_stride_floats, nql, q_stride_floats, top_k, m_ids, m_d, m_c));
    unsigned char* d_buf = nullptr;
    VERS_CUDA(cudaMalloc(&d_buf, slot * (W + 1)));
    int32_t rc = VERS_OK;
    cudaError_t e = cudaMemcpyAsync(d_buf + slot * W, mine.data(), slot, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        ncclResult_t r = g_nccl.AllGather(d_buf + slot * W, d_buf, slot, ncclChar, cm->nccl, ctx->stream);
        if (r != ncclSuccess) rc = fail(VERS_ERR_CUDA, "ncclAllGather(lsh results): %s", g_nccl.GetErrorString(r));
    }
    if (rc == VERS_OK && e == cudaSuccess) e = cudaMemcpyAsync(all.data(), d_buf, slot * W, cudaMemcpyDeviceToHost, ctx->stream);
    if (rc == VERS_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_buf);
    if (rc == VERS_OK && e != cudaSuccess) rc = fail(VERS_ERR_CUDA, "sharded_lsh_search: %s", cudaGetErrorString(e));
    VERS_TRY(rc);
    for (uint32_t r = 0; r < W; ++r) {
        const uint32_t r0 = std::min(nq, r * per), rn = std::min(nq, r0 + per) - r0;
        const unsigned char* blk = all.data() + (size_t)slot * r;
        memcpy(ids + (size_t)r0 * top_k, blk, (size_t)rn * top_k * 8);
        memcpy(dists + (size_t)r0 * top_k, blk + nk * 8, (size_t)rn * top_k * 4);
        if (counts) memcpy(counts + r0, blk + nk * 12, (size_t)rn * 4);
    }
    return VERS_OK;
}
