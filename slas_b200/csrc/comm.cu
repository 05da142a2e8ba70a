// comm.cu -- multi-GPU plumbing for the sharded vector ops.  One process per GPU;
// each rank owns a contiguous slice of the vector.  The only exchange on the whole hot path
// is ONE ncclAllReduce(sum) of the rank-local f64 partial, enqueued on the same stream right
// behind the kernel that produced it (no host sync in between).  The reference has no
// distributed code at all (single-threaded CPU crate), so nothing here restates it.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when the
// process imported torch, else the system one), so libslas_cuda.so itself loads without it.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"
#include "comm.cuh"

namespace slas {

namespace {
struct NcclUniqueId { char internal[128]; };
typedef void *ncclComm_t_;
typedef int (*fn_GetUniqueId)(NcclUniqueId *);
typedef int (*fn_CommInitRank)(ncclComm_t_ *, int, NcclUniqueId, int);
typedef int (*fn_CommDestroy)(ncclComm_t_);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t);
typedef const char *(*fn_GetErrorString)(int);
constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64
constexpr int kNcclSum = 0;      // ncclRedOp_t::ncclSum

struct Nccl {
    void *handle = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
} g_nccl;

struct Comm {
    ncclComm_t_ comm = nullptr;
    int nranks = 1, rank = 0;
    // peer-memory path (p2p.cuh): own mailbox + the peers' mailboxes mapped through CUDA IPC
    PeerSlot *own_mailbox = nullptr;
    PeerSlot *mailbox[kMaxPeers] = {};
    bool p2p = false;
    int p2p_nranks = 1, p2p_rank = 0;
    unsigned long long epoch = 0;
} g_comm;
std::mutex g_comm_mu;

int load_nccl() {
    if (g_nccl.handle) return SLAS_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void *h = nullptr;
    for (const char *nme : names) {
        h = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(SLAS_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (fn_AllReduce)dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (fn_GetErrorString)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce ||
        !g_nccl.GetErrorString)
        return fail(SLAS_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
    g_nccl.handle = h;
    return SLAS_OK;
}
}  // namespace

int comm_nranks() { return g_comm.p2p ? g_comm.p2p_nranks : (g_comm.comm ? g_comm.nranks : 1); }

bool comm_next_peer_exchange(PeerExchange *out) {
    if (!g_comm.p2p) return false;
    for (int r = 0; r < kMaxPeers; ++r) out->mailbox[r] = g_comm.mailbox[r];
    out->nranks = g_comm.p2p_nranks;
    out->rank = g_comm.p2p_rank;
    out->epoch = ++g_comm.epoch;
    return true;
}

void comm_rollback_peer_exchange(unsigned long long epoch) {
    if (g_comm.p2p && g_comm.epoch == epoch) --g_comm.epoch;
}

int comm_allreduce_f64(double *dev_inout, size_t count, cudaStream_t st) {
    if (!g_comm.comm || g_comm.nranks == 1) return SLAS_OK;  // single rank: the local partial is the sum
    const int r = g_nccl.AllReduce(dev_inout, dev_inout, count, kNcclFloat64, kNcclSum, g_comm.comm, st);
    if (r != 0) return fail(SLAS_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    return SLAS_OK;
}

}  // namespace slas

using namespace slas;

extern "C" {

int slas_cuda_nccl_unique_id(void *id128) {
    SLAS_REQUIRE(id128, "null id buffer");
    std::lock_guard<std::mutex> lk(g_comm_mu);
    SLAS_TRY(load_nccl());
    NcclUniqueId id;
    const int r = g_nccl.GetUniqueId(&id);
    if (r != 0) return fail(SLAS_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(id128, &id, sizeof(id));
    return SLAS_OK;
}

int slas_cuda_comm_init(int nranks, int rank, const void *id128) {
    SLAS_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (g_comm.comm) return fail(SLAS_ERR_INVALID, "communicator already initialised");
    if (nranks == 1) {
        g_comm.nranks = 1;
        g_comm.rank = 0;
        return SLAS_OK;
    }
    SLAS_REQUIRE(id128, "null unique id");
    SLAS_TRY(load_nccl());
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t_ comm = nullptr;
    const int r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r != 0) return fail(SLAS_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    g_comm.comm = comm;
    g_comm.nranks = nranks;
    g_comm.rank = rank;
    return SLAS_OK;
}

int slas_cuda_comm_destroy(void) {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (g_comm.comm) {
        g_nccl.CommDestroy(g_comm.comm);
        g_comm.comm = nullptr;
    }
    g_comm.nranks = 1;
    g_comm.rank = 0;
    return SLAS_OK;
}

int slas_cuda_comm_info(int *nranks, int *rank) {
    if (nranks) *nranks = comm_nranks();
    if (rank) *rank = g_comm.p2p ? g_comm.p2p_rank : (g_comm.comm ? g_comm.rank : 0);
    return SLAS_OK;
}

int slas_cuda_p2p_handle(void *handle64) {
    SLAS_REQUIRE(handle64, "null handle buffer");
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (!g_comm.own_mailbox) {
        SLAS_CUDA_TRY(cudaMalloc(&g_comm.own_mailbox, kMailboxBytes));
        SLAS_CUDA_TRY(cudaMemset(g_comm.own_mailbox, 0, kMailboxBytes));
        SLAS_CUDA_TRY(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    SLAS_CUDA_TRY(cudaIpcGetMemHandle(&h, g_comm.own_mailbox));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, sizeof(h));
    return SLAS_OK;
}

int slas_cuda_p2p_init(int nranks, int rank, const void *handles) {
    SLAS_REQUIRE(nranks >= 1 && nranks <= kMaxPeers && rank >= 0 && rank < nranks, "bad rank %d of %d (max %d peers)", rank,
                 nranks, kMaxPeers);
    SLAS_REQUIRE(handles, "null handles");
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    std::lock_guard<std::mutex> lk(g_comm_mu);
    SLAS_REQUIRE(g_comm.own_mailbox, "call slas_cuda_p2p_handle first");
    SLAS_REQUIRE(!g_comm.p2p, "peer exchange already initialised");
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            g_comm.mailbox[r] = g_comm.own_mailbox;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        SLAS_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        g_comm.mailbox[r] = (PeerSlot *)p;
    }
    g_comm.p2p_nranks = nranks;
    g_comm.p2p_rank = rank;
    g_comm.epoch = 0;
    g_comm.p2p = nranks > 1;
    return SLAS_OK;
}

int slas_cuda_p2p_destroy(void) {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (g_comm.p2p) {
        cudaDeviceSynchronize();
        for (int r = 0; r < g_comm.p2p_nranks; ++r)
            if (r != g_comm.p2p_rank && g_comm.mailbox[r]) cudaIpcCloseMemHandle(g_comm.mailbox[r]);
    }
    for (int r = 0; r < kMaxPeers; ++r) g_comm.mailbox[r] = nullptr;
    g_comm.p2p = false;
    g_comm.p2p_nranks = 1;
    g_comm.p2p_rank = 0;
    if (g_comm.own_mailbox) {
        cudaFree(g_comm.own_mailbox);
        g_comm.own_mailbox = nullptr;
    }
    return SLAS_OK;
}

int slas_cuda_shard_range(size_t len, size_t elem_size, int nranks, int rank, size_t *begin, size_t *end) {
    SLAS_REQUIRE(begin && end, "null output");
    SLAS_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
    SLAS_REQUIRE(elem_size >= 1 && elem_size <= 16 && 16 % elem_size == 0, "element size must divide 16");
    const size_t gran = 16 / elem_size;                       // elements per 16 bytes
    const size_t blocks = (len + gran - 1) / gran;            // 16-byte blocks, last may be ragged
    const size_t b0 = blocks * (size_t)rank / (size_t)nranks;
    const size_t b1 = blocks * (size_t)(rank + 1) / (size_t)nranks;
    size_t lo = b0 * gran, hi = b1 * gran;
    if (lo > len) lo = len;
    if (hi > len || rank == nranks - 1) hi = len;
    *begin = lo;
    *end = hi;
    return SLAS_OK;
}

}  // extern "C"
