// multi.cu -- several GPUs inside one process (multi.cuh): device set-up with peer access and mailboxes, one worker
// thread per device for ops on host operands, and the in-process sharded reductions (slas_cuda_*_multi): one
// device-resident slice per device, exchange inside the reduction kernels over NVLink peer memory (p2p.cuh), every
// device ends with the same bits.  The reference has no device or multi-threaded code at all (a single-threaded CPU
// crate): nothing here restates it; the op semantics are those of the single-object entries (include/slas_cuda.h).
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <memory>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "multi.cuh"

namespace slas {

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    const std::function<int(int)> *task = nullptr;
    int part = 0;
    bool has_task = false, done = false, quit = false;
    int status = SLAS_OK;
    char error[512] = "";
};

struct Multi {
    std::atomic<int> ndev{1};
    bool peers = false;                       // every pair of devices has peer access and a mailbox
    PeerSlot *mailbox[kMaxPeers] = {};        // mailbox[d] lives on device d
    std::vector<std::unique_ptr<Worker>> workers;
    std::mutex init_mu, run_mu;
    unsigned long long epoch = 0;
} g_multi;

void worker_main(Worker *w, int device) {
    cudaSetDevice(device);
    for (;;) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->has_task || w->quit; });
        if (w->quit) return;
        const std::function<int(int)> *task = w->task;
        const int part = w->part;
        lk.unlock();
        const int st = (*task)(part);
        lk.lock();
        w->status = st;
        if (st != SLAS_OK) {
            strncpy(w->error, slas_cuda_last_error(), sizeof(w->error) - 1);  // this thread's message, for the caller's thread
            w->error[sizeof(w->error) - 1] = 0;
        }
        w->has_task = false;
        w->done = true;
        w->cv.notify_all();
    }
}

}  // namespace

int multi_ndev() { return g_multi.ndev.load(std::memory_order_acquire); }

void multi_autoinit() {
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("SLAS_CUDA_DEVICES");
        if (!env || !*env) return;
        const int n = atoi(env);
        if (n > 1 || strcmp(env, "all") == 0) slas_cuda_init_devices(n > 1 ? n : 0);
    });
}

DevScope::DevScope(const void *p) {
    multi_autoinit();
    if (multi_ndev() <= 1 || !p) return;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return; }
    if (attr.type != cudaMemoryTypeDevice) return;
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return; }
    if (cur != attr.device && cudaSetDevice(attr.device) == cudaSuccess) prev = cur;
}
DevScope::~DevScope() {
    if (prev >= 0) cudaSetDevice(prev);
}

bool multi_split_wanted(const void *p, size_t bytes) {
    if (multi_ndev() <= 1 || !p) return false;
    const long mb = tunable("multi_split_mb");
    if (bytes < (size_t)(mb > 0 ? mb : 64) << 20) return false;
    bool dev = true;
    if (pointer_is_device(p, &dev) != SLAS_OK) return false;
    return !dev;
}

int multi_run(int parts, const std::function<int(int)> &fn) {
    std::lock_guard<std::mutex> run_lk(g_multi.run_mu);
    if (parts > (int)g_multi.workers.size()) return fail(SLAS_ERR_INVALID, "multi_run: %d parts but %zu devices", parts, g_multi.workers.size());
    for (int d = 0; d < parts; ++d) {
        Worker *w = g_multi.workers[d].get();
        std::lock_guard<std::mutex> lk(w->mu);
        w->task = &fn;
        w->part = d;
        w->done = false;
        w->has_task = true;
        w->cv.notify_all();
    }
    int status = SLAS_OK;
    for (int d = 0; d < parts; ++d) {
        Worker *w = g_multi.workers[d].get();
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->status != SLAS_OK && status == SLAS_OK) {
            status = w->status;
            set_error("device %d: %s", d, w->error);
        }
    }
    return status;
}

bool multi_peer_exchange(const int *devices, int nslices, int rank, unsigned long long epoch, PeerExchange *out) {
    if (!g_multi.peers || nslices > kMaxPeers) return false;
    for (int i = 0; i < nslices; ++i) {
        if (devices[i] < 0 || devices[i] >= multi_ndev()) return false;
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return false;  // two slices on one device: their kernels could not wait for each other
        out->mailbox[i] = g_multi.mailbox[devices[i]];
    }
    out->nranks = nslices;
    out->rank = rank;
    out->epoch = epoch;
    return true;
}
unsigned long long multi_next_epoch() { return ++g_multi.epoch; }

Context *multi_context(int device) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    Context *c = nullptr;
    if (cur != device) cudaSetDevice(device);
    const int r = get_context(&c);
    if (cur != device) cudaSetDevice(cur);
    return r == SLAS_OK ? c : nullptr;
}

// ---- in-process sharded reductions: slice i device-resident on its own GPU ----------------------------------------
// kind 0 dot, 1 squared norm (mode 1: sqrt).  normalize != 0: scale every slice by the global norm afterwards.
template <typename T>
static int api_reduce_multi(int kind, int mode, int nslices, const size_t *lens, const T *const *a, const T *const *b, T *out,
                            T *const *a_mut, bool normalize) {
    SLAS_REQUIRE(nslices >= 1 && nslices <= kMaxPeers, "multi: %d slices (1..%d)", nslices, kMaxPeers);
    SLAS_REQUIRE(lens && a && (kind == 1 || b), "multi: null argument");
    SLAS_REQUIRE(normalize || out, "null result pointer");
    int devices[kMaxPeers];
    Context *ctxs[kMaxPeers];
    int cur = -1;
    SLAS_CUDA_TRY(cudaGetDevice(&cur));
    for (int i = 0; i < nslices; ++i) {
        SLAS_REQUIRE(a[i] && (kind == 1 || b[i]), "multi: null slice %d", i);
        cudaPointerAttributes attr;
        SLAS_CUDA_TRY(cudaPointerGetAttributes(&attr, a[i]));
        SLAS_REQUIRE(attr.type == cudaMemoryTypeDevice, "multi: slice %d is not device memory (host vectors: call the plain entry, it splits them)", i);
        devices[i] = attr.device;
        if (kind != 1) {
            cudaPointerAttributes attr_b;
            SLAS_CUDA_TRY(cudaPointerGetAttributes(&attr_b, b[i]));
            SLAS_REQUIRE(attr_b.type == cudaMemoryTypeDevice && attr_b.device == attr.device, "multi: slice %d: a and b live on different devices", i);
        }
        ctxs[i] = multi_context(devices[i]);
        if (!ctxs[i]) return SLAS_ERR_CUDA;
    }
    // lock the contexts in device order (two slices on one device share a context: lock it once)
    int order[kMaxPeers];
    for (int i = 0; i < nslices; ++i) order[i] = i;
    for (int i = 1; i < nslices; ++i)
        for (int j = i; j > 0 && devices[order[j]] < devices[order[j - 1]]; --j) { const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    std::vector<std::unique_lock<std::mutex>> locks;
    for (int q = 0; q < nslices; ++q)
        if (q == 0 || devices[order[q]] != devices[order[q - 1]]) locks.emplace_back(ctxs[order[q]]->mu);
    for (int i = 0; i < nslices; ++i)
        SLAS_REQUIRE(!ctxs[i]->capturing, "sharded reductions cannot be recorded in a graph (the exchange epoch is per call)");

    int status = SLAS_OK;
    PeerExchange px[kMaxPeers];
    const unsigned long long epoch = multi_next_epoch();
    bool fused = nslices > 1;
    for (int i = 0; i < nslices && fused; ++i) fused = multi_peer_exchange(devices, nslices, i, epoch, &px[i]);
    auto typed_of = [](Context *c) { return reinterpret_cast<T *>(c->scalars + 4); };
    if (fused || nslices == 1) {
        // one launch per device; the last block of each exchanges its partial with all the others (rank-ordered sum)
        for (int i = 0; i < nslices && status == SLAS_OK; ++i) {
            cudaSetDevice(devices[i]);
            status = launch_reduce<T>(ctxs[i], ctxs[i]->stream, kind, lens[i], a[i], kind == 1 ? a[i] : b[i], ctxs[i]->scalars,
                                      typed_of(ctxs[i]), mode, 0, fused ? &px[i] : nullptr);
        }
    } else {
        // no peer access (or two slices on one device): f64 partials to the host, added there in slice order
        double sum = 0.0;
        for (int i = 0; i < nslices && status == SLAS_OK; ++i) {
            cudaSetDevice(devices[i]);
            status = launch_reduce<T>(ctxs[i], ctxs[i]->stream, kind, lens[i], a[i], kind == 1 ? a[i] : b[i], ctxs[i]->scalars, nullptr, mode, 0, nullptr);
            double part = 0.0;
            if (status == SLAS_OK) status = deliver_scalar(ctxs[i], ctxs[i]->scalars, &part, sizeof(double));
            sum += part;
        }
        const T typed = mode == 1 ? (T)sqrt(sum) : (T)sum;
        for (int i = 0; i < nslices && status == SLAS_OK; ++i) {
            cudaSetDevice(devices[i]);
            if (cudaMemcpyAsync(typed_of(ctxs[i]), &typed, sizeof(T), cudaMemcpyHostToDevice, ctxs[i]->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctxs[i]->stream) != cudaSuccess) {
                cudaGetLastError();
                status = fail(SLAS_ERR_CUDA, "multi: broadcasting the combined scalar to device %d failed", devices[i]);
            }
        }
    }
    if (status == SLAS_OK && normalize) {
        for (int i = 0; i < nslices && status == SLAS_OK; ++i) {
            cudaSetDevice(devices[i]);
            status = launch_scale_div<T>(ctxs[i], ctxs[i]->stream, lens[i], a_mut[i], a_mut[i], typed_of(ctxs[i]));
        }
    }
    if (status == SLAS_OK && out) {
        cudaSetDevice(devices[0]);
        status = deliver_scalar(ctxs[0], typed_of(ctxs[0]), out, sizeof(T));
        // a device-resident result was written with an asynchronous copy on slice 0's stream; the other devices'
        // streams have nothing the caller waits for except their own kernels
    }
    cudaSetDevice(cur);
    return status;
}

}  // namespace slas

using namespace slas;

extern "C" {

int slas_cuda_init_devices(int ndev) {
    std::lock_guard<std::mutex> lk(g_multi.init_mu);
    int count = 0;
    SLAS_CUDA_TRY(cudaGetDeviceCount(&count));
    if (ndev <= 0) ndev = count;
    SLAS_REQUIRE(ndev <= count, "init_devices: %d devices requested, %d visible", ndev, count);
    SLAS_REQUIRE(ndev <= 32, "init_devices: at most 32 devices");
    if (ndev <= multi_ndev() && !(ndev == 1 && g_multi.workers.empty())) return SLAS_OK;  // already there
    int cur = -1;
    SLAS_CUDA_TRY(cudaGetDevice(&cur));
    int status = SLAS_OK;
    for (int d = 0; d < ndev && status == SLAS_OK; ++d) {
        if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); status = fail(SLAS_ERR_CUDA, "init_devices: cannot select device %d", d); break; }
        Context *c;
        status = get_context(&c);
    }
    // peer access between every pair + one mailbox per device: the fused exchange of the *_multi reductions
    bool peers = status == SLAS_OK && ndev > 1 && ndev <= kMaxPeers;
    for (int d = 0; d < ndev && peers; ++d) {
        cudaSetDevice(d);
        for (int e = 0; e < ndev && peers; ++e) {
            if (e == d) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, d, e) != cudaSuccess || !can) { cudaGetLastError(); peers = false; break; }
            const cudaError_t pe = cudaDeviceEnablePeerAccess(e, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) peers = false;
            cudaGetLastError();
        }
    }
    for (int d = 0; d < ndev && peers; ++d) {
        cudaSetDevice(d);
        if (!g_multi.mailbox[d]) {
            if (cudaMalloc(&g_multi.mailbox[d], kMailboxBytes) != cudaSuccess || cudaMemset(g_multi.mailbox[d], 0, kMailboxBytes) != cudaSuccess ||
                cudaDeviceSynchronize() != cudaSuccess) {
                cudaGetLastError();
                peers = false;
            }
        }
    }
    cudaSetDevice(cur);
    SLAS_TRY(status);
    g_multi.peers = peers;
    while ((int)g_multi.workers.size() < ndev) {
        const int d = (int)g_multi.workers.size();
        g_multi.workers.emplace_back(new Worker());
        Worker *w = g_multi.workers.back().get();
        w->th = std::thread(worker_main, w, d);
    }
    g_multi.ndev.store(ndev, std::memory_order_release);
    return SLAS_OK;
}

int slas_cuda_devices(int *ndev, int *peer_exchange) {
    if (ndev) *ndev = multi_ndev();
    if (peer_exchange) *peer_exchange = g_multi.peers ? 1 : 0;
    return SLAS_OK;
}

int slas_cuda_malloc_on(int device, void **dptr, size_t bytes) {
    SLAS_REQUIRE(dptr, "null dptr");
    int cur = -1;
    SLAS_CUDA_TRY(cudaGetDevice(&cur));
    SLAS_CUDA_TRY(cudaSetDevice(device));
    Context *c;
    int status = get_context(&c);
    if (status == SLAS_OK && cudaMalloc(dptr, bytes ? bytes : 16) != cudaSuccess) {
        cudaGetLastError();
        status = fail(SLAS_ERR_CUDA, "cudaMalloc of %zu bytes on device %d failed", bytes, device);
    }
    cudaSetDevice(cur);
    return status;
}

int slas_cuda_sdot_multi(int nslices, const size_t *lens, const float *const *a, const float *const *b, float *out) {
    return api_reduce_multi<float>(0, 0, nslices, lens, a, b, out, nullptr, false);
}
int slas_cuda_ddot_multi(int nslices, const size_t *lens, const double *const *a, const double *const *b, double *out) {
    return api_reduce_multi<double>(0, 0, nslices, lens, a, b, out, nullptr, false);
}
int slas_cuda_snrm2_multi(int nslices, const size_t *lens, const float *const *a, float *out) {
    return api_reduce_multi<float>(1, 1, nslices, lens, a, a, out, nullptr, false);
}
int slas_cuda_dnrm2_multi(int nslices, const size_t *lens, const double *const *a, double *out) {
    return api_reduce_multi<double>(1, 1, nslices, lens, a, a, out, nullptr, false);
}
int slas_cuda_snormalize_multi(int nslices, const size_t *lens, float *const *a) {
    return api_reduce_multi<float>(1, 1, nslices, lens, a, a, nullptr, a, true);
}
int slas_cuda_dnormalize_multi(int nslices, const size_t *lens, double *const *a) {
    return api_reduce_multi<double>(1, 1, nslices, lens, a, a, nullptr, a, true);
}

}  // extern "C"

// workers are joined at process exit
namespace {
struct MultiShutdown {
    ~MultiShutdown() {
        for (auto &w : g_multi.workers) {
            {
                std::lock_guard<std::mutex> lk(w->mu);
                w->quit = true;
                w->cv.notify_all();
            }
            if (w->th.joinable()) w->th.join();
        }
    }
} g_multi_shutdown;
}  // namespace
