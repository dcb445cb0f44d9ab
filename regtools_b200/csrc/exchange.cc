// regtools_b200/csrc/exchange.cc — the path's one exchange, inside the library: NCCL over NVLink / NVSwitch.
//
// Contigs are sharded over the ranks (one process per GPU), so the per-rank junction tables are disjoint and the merge is a
// concatenation followed by the global first-seen ranking and the sort.  Round 1 did this in Python (table D2H -> pinned ->
// H2D -> all_gather_into_tensor of max-padded slots -> D2H of world x cap on every rank -> numpy); here the compacted table
// never leaves the device until it is final: counts are all-gathered (4 bytes per rank), every rank sends its compacted
// entries straight from HBM to the root with ncclSend / ncclRecv (exact sizes, no padding, only the root receives), and the
// root ranks, sorts and downloads the merged table once.
// NCCL is resolved with dlopen at the first rtjx_comm_* call: the single-GPU CLI never loads it, and inside a process that
// already loaded a libnccl.so.2 (PyTorch bundles one) that copy is the one used.
#include "engine.h"
#include "buffer_cache.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>

namespace rtjx {

namespace {
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
        bool ok = true;
        auto sym = [&](const char* n) -> void* { void* p = dlsym(lib, n); if (!p) { ok = false; err = std::string("NCCL symbol missing: ") + n; } return p; };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        Send = reinterpret_cast<decltype(Send)>(sym("ncclSend"));
        Recv = reinterpret_cast<decltype(Recv)>(sym("ncclRecv"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!ok) { lib = nullptr; return false; }
        return true;
    }
};
// One communicator per process (one process per GPU): handles come and go — bench.py's end-to-end leg makes a fresh one per
// step — the communicator must not (ncclCommInitRank costs hundreds of milliseconds).
struct Comm { ncclComm_t comm = nullptr; int rank = 0, world = 1, device = -1; uint32_t* d_counts = nullptr; uint32_t* h_counts = nullptr; };
Nccl g_nccl;
Comm g_comm;
std::mutex g_mu;
}  // namespace

int comm_unique_id(void* id, std::string* err) {
    std::lock_guard<std::mutex> g(g_mu);
    if (!g_nccl.load()) { *err = g_nccl.err; return RTJX_E_UNSUPPORTED; }
    static_assert(sizeof(ncclUniqueId) == RTJX_COMM_ID_BYTES, "rtjx.h: RTJX_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
    ncclUniqueId u;
    ncclResult_t r = g_nccl.GetUniqueId(&u);
    if (r != ncclSuccess) { *err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return RTJX_E_CUDA; }
    memcpy(id, &u, sizeof u);
    return RTJX_OK;
}

int comm_init(const void* id, int rank, int world, int device, std::string* err) {
    std::lock_guard<std::mutex> g(g_mu);
    if (world < 1 || rank < 0 || rank >= world) { *err = "rtjx_comm_init: bad rank / world"; return RTJX_E_ARG; }
    if (!g_nccl.load()) { *err = g_nccl.err; return RTJX_E_UNSUPPORTED; }
    if (g_comm.comm) { g_nccl.CommDestroy(g_comm.comm); g_comm.comm = nullptr; }
    cudaSetDevice(device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = g_nccl.CommInitRank(&g_comm.comm, world, u, rank);
    if (r != ncclSuccess) { g_comm.comm = nullptr; *err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); return RTJX_E_CUDA; }
    g_comm.rank = rank; g_comm.world = world; g_comm.device = device;
    if (!g_comm.d_counts) {
        if (cudaMalloc(&g_comm.d_counts, 1024 * sizeof(uint32_t)) != cudaSuccess || cudaHostAlloc(&g_comm.h_counts, 1024 * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) {
            *err = "rtjx_comm_init: allocation failed"; return RTJX_E_CUDA;
        }
    }
    if (world > 1023) { *err = "rtjx_comm_init: more than 1023 ranks"; return RTJX_E_ARG; }
    return RTJX_OK;
}

void comm_destroy() {
    std::lock_guard<std::mutex> g(g_mu);
    if (g_comm.comm && g_nccl.lib) g_nccl.CommDestroy(g_comm.comm);
    g_comm.comm = nullptr; g_comm.world = 1; g_comm.rank = 0;
}

#define CKX(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(RTJX_E_CUDA, std::string(#call " failed: ") + cudaGetErrorString(e__));      \
    } while (0)
#define CKN(call)                                                                                    \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess)                                                                      \
            return fail(RTJX_E_CUDA, std::string(#call " failed: ") + g_nccl.GetErrorString(r__));   \
    } while (0)

// rtjx_gather: after it, the ROOT's handle holds the merged table of all ranks (names ranked over the whole file, sorted by
// compare_junctions); the other ranks keep their own shard's table.
int Engine::gather(int root) {
    NvtxRange nvtx("rtjx:gather (NCCL counts + tables -> root, rank + sort on the root)");
    if (!g_comm.comm || g_comm.world == 1) return finalize(nullptr);
    if (bc_mode_) return fail(RTJX_E_UNSUPPORTED, "-b barcodes are not exchanged between contig shards");
    if (root < 0 || root >= g_comm.world) return fail(RTJX_E_ARG, "rtjx_gather: bad root");
    int rc = ensure_device();
    if (rc) return rc;
    if (prm_.device != g_comm.device) return fail(RTJX_E_ARG, "rtjx_gather: the communicator was created for another device");
    const int world = g_comm.world, rank = g_comm.rank;
    cudaStream_t st = stream_;
    // ---- how many junctions does every rank hold?
    uint32_t n = 0;
    if (d_table_) { if ((rc = sync_counters(st))) return rc; n = h_counters_[CTR_NUNIQUE]; }
    g_comm.h_counts[world] = n;
    CKX(cudaMemcpyAsync(g_comm.d_counts + world, g_comm.h_counts + world, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CKN(g_nccl.AllGather(g_comm.d_counts + world, g_comm.d_counts, 1, ncclUint32, g_comm.comm, st));
    CKX(cudaMemcpyAsync(g_comm.h_counts, g_comm.d_counts, (size_t)world * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CKX(cudaStreamSynchronize(st));
    uint64_t total = 0, my_off = 0;
    for (int r = 0; r < world; ++r) { if (r == rank) my_off = total; total += g_comm.h_counts[r]; }
    if (total >= 0xfffffff0ull) return fail(RTJX_E_STATE, "rtjx_gather: merged table too large");
    const bool is_root = rank == root;
    const uint32_t cap = is_root ? (uint32_t)total : n;
    final_.clear(); pinned_final_n_ = 0;
    if (cap == 0) { finalized_ = true; dirty_ = false; return RTJX_OK; }
    if ((rc = ensure_finalize_buffers(cap, contigs_.size()))) return rc;
    if (rank_dirty_) {
        std::vector<uint32_t> cr = contig_rank_table();
        if (!cr.empty()) CKX(cudaMemcpyAsync(d_rank_, cr.data(), cr.size() * 4, cudaMemcpyHostToDevice, st));
        CKX(cudaStreamSynchronize(st));
        rank_dirty_ = false;
    }
    // ---- this rank's entries, compacted in place where the merged table wants them; the others' arrive from HBM to HBM
    OutJunction* mine = d_out_ + (is_root ? my_off : 0);
    if (n) launch_table_compact(table_ref(), n, mine, st);
    CKN(g_nccl.GroupStart());
    if (is_root) {
        uint64_t off = 0;
        for (int r = 0; r < world; ++r) {
            const uint32_t c = g_comm.h_counts[r];
            if (r != rank && c) CKN(g_nccl.Recv(d_out_ + off, (size_t)c * sizeof(OutJunction), ncclUint8, r, g_comm.comm, st));
            off += c;
        }
    } else if (n) {
        CKN(g_nccl.Send(mine, (size_t)n * sizeof(OutJunction), ncclUint8, root, g_comm.comm, st));
    }
    CKN(g_nccl.GroupEnd());
    // ---- names (rank of first appearance in the file = (contig, first ordinal) order: a contig lives on one rank) and order
    launch_finalize_sort(d_out_, cap, d_rank_, (uint32_t)contigs_.size(), d_ws_, ws_cap_, st, /*rank_by_contig=*/true);
    CKX(cudaMemcpyAsync(h_final_, d_out_, (size_t)cap * sizeof(OutJunction), cudaMemcpyDeviceToHost, st));
    CKX(cudaStreamSynchronize(st));
    CKX(cudaGetLastError());
    pinned_final_n_ = cap;
    stats_.kernel_launches += 2; stats_.d2h_bytes += (size_t)cap * sizeof(OutJunction) + (size_t)world * 4;
    finalized_ = true; dirty_ = false;
    return RTJX_OK;
}

}  // namespace rtjx
