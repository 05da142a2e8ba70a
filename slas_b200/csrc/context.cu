// context.cu -- process-global context, error reporting and the device-resident storage
// entry points of the C ABI (include/slas_cuda.h).  A reference backend is a stateless ZST
// (src/backends/rust.rs:2-3), so all device state lives here.
#include <sched.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "multi.cuh"

namespace slas {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
    return code;
}

struct Tunable { const char *name; long value; };
static Tunable g_tunables[] = {{"transpose_variant", 0}, {"gemm32d_variant", 0}, {"gemm32s_variant", 0},
                               {"gemm16s_variant", 0}, {"gemm4s_variant", 0}, {"gemm_large_variant", 0},
                               {"tc_variant", 0}, {"gemm_tiny_variant", 0}, {"host_slot_mb", 0}, {"gemv_variant", 0}, {"bdot_variant", 0}, {"gemm_jit", 0}, {"bdot_limit", 0}, {"transpose_staged_limit", 0},
                               {"reduce_small", 0}, {"reduce_small_threads", 0}, {"reduce_small_cps", 0}, {"reduce_small_tail", 0},
                               {"reduce_phases", 0}, {"ewise_small", 0}, {"pdl", 0}, {"multi_split_mb", 0}, {"host_copy_threads", 0}, {"host_bounce", 0}, {"host_copy_nt", 0}, {"bdot_stage_kb", 0}, {"host_result", 0}, {"normalize_fused", 0}, {"staged_jit", 0}, {"bdot_gcap", 0}, {"transpose_stage_kb", 0}, {"staged_cta_cap", 0}, {"bnorm_cluster", 0}, {"bnorm_cluster_ctas", 0}, {"gemm4d_variant", 0}, {"gemm_jit_bulk_c", 0}};
long tunable(const char *name) {
    for (auto &t : g_tunables)
        if (strcmp(t.name, name) == 0) return t.value;
    return 0;
}

constexpr int kMaxDevices = 32;
static Context g_ctx[kMaxDevices];
static std::mutex g_init_mu;

static void release_context(Context *c) {
    cudaFree(c->partials);
    cudaFree(c->ticket);
    cudaFree(c->scalars);
    cudaFree(c->ll_slots);
    cudaFree(c->phases);
    c->ll_slots = nullptr; c->phases = nullptr;
    cudaFree(c->workspace);
    for (void *p : c->retired) cudaFree(p);
    c->retired.clear();
    cudaFree(c->staging);
    if (c->pinned_scalars) cudaFreeHost(c->pinned_scalars);
    if (c->bounce) cudaFreeHost(c->bounce);
    c->bounce = nullptr; c->bounce_bytes = 0;
    for (int i = 0; i < kBounceSlots; ++i) {
        if (c->bounce_events[i]) cudaEventDestroy(c->bounce_events[i]);
        c->bounce_events[i] = nullptr;
    }
    for (int i = 0; i < 3; ++i) {
        if (c->copy_streams[i]) cudaStreamDestroy(c->copy_streams[i]);
        if (c->copy_events[i]) cudaEventDestroy(c->copy_events[i]);
        c->copy_streams[i] = nullptr;
        c->copy_events[i] = nullptr;
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    c->own_stream = c->stream = nullptr;
    c->partials = nullptr; c->ticket = nullptr; c->scalars = nullptr;
    c->workspace = nullptr; c->workspace_bytes = 0; c->pinned_scalars = nullptr; c->pinned_scalars_dev = nullptr;
    c->staging = nullptr; c->staging_bytes = 0;
    c->live_graphs = 0; c->capturing = false;
    cudaGetLastError();
}

static int init_context_resources(Context *c, int dev) {
    cudaDeviceProp prop;
    SLAS_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(SLAS_ERR_CUDA, "device %d is sm_%d%d; libslas_cuda is built for sm_100a only", dev,
                    prop.major, prop.minor);
    c->sm_count = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    c->hbm_bytes = prop.totalGlobalMem;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    SLAS_CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    SLAS_CUDA_TRY(cudaMalloc(&c->partials, sizeof(double) * kMaxPartials * 2));
    SLAS_CUDA_TRY(cudaMemset(c->partials, 0, sizeof(double) * kMaxPartials * 2));
    SLAS_CUDA_TRY(cudaMalloc(&c->ticket, sizeof(unsigned int) * 4));
    SLAS_CUDA_TRY(cudaMemset(c->ticket, 0, sizeof(unsigned int) * 4));
    SLAS_CUDA_TRY(cudaMalloc(&c->scalars, sizeof(double) * 16));
    SLAS_CUDA_TRY(cudaMemset(c->scalars, 0, sizeof(double) * 16));
    // + past the sets: per set one norm-broadcast slot and one acknowledge counter of normalize_fused_kernel
    SLAS_CUDA_TRY(cudaMalloc(&c->ll_slots, (size_t)16 * (3 * kMaxPartials * kRedScratchSets + 2 * kRedScratchSets)));
    SLAS_CUDA_TRY(cudaMemset(c->ll_slots, 0, (size_t)16 * (3 * kMaxPartials * kRedScratchSets + 2 * kRedScratchSets)));
    SLAS_CUDA_TRY(cudaMalloc(&c->phases, sizeof(unsigned long long) * 8));
    SLAS_CUDA_TRY(cudaMemset(c->phases, 0, sizeof(unsigned long long) * 8));
    SLAS_CUDA_TRY(cudaHostAlloc(&c->pinned_scalars, sizeof(double) * 16, cudaHostAllocMapped | cudaHostAllocPortable));
    SLAS_CUDA_TRY(cudaHostGetDevicePointer(&c->pinned_scalars_dev, c->pinned_scalars, 0));
    for (int i = 0; i < 3; ++i) {
        SLAS_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_streams[i], cudaStreamNonBlocking));
        SLAS_CUDA_TRY(cudaEventCreateWithFlags(&c->copy_events[i], cudaEventDisableTiming));
    }
    SLAS_CUDA_TRY(cudaDeviceSynchronize());
    return SLAS_OK;
}

// Called under g_init_mu.  A failure part-way gives back what was already made, so a retry starts clean.
static int init_context(Context *c, int dev) {
    const int r = init_context_resources(c, dev);
    if (r != SLAS_OK) {
        release_context(c);
        return r;
    }
    c->device.store(dev, std::memory_order_release);  // publishes every field written above
    return SLAS_OK;
}

int get_context(Context **out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SLAS_ERR_CUDA, "no CUDA device: %s (libslas_cuda has no CPU fallback)",
                    cudaGetErrorString(e));
    }
    if (dev < 0 || dev >= kMaxDevices) return fail(SLAS_ERR_CUDA, "device index %d out of range", dev);
    Context *c = &g_ctx[dev];
    // double-checked: the acquire load pairs with init_context's release store (libtest calls in concurrently)
    if (c->device.load(std::memory_order_acquire) != dev) {
        std::lock_guard<std::mutex> lk(g_init_mu);
        if (c->device.load(std::memory_order_relaxed) != dev) SLAS_TRY(init_context(c, dev));
    }
    *out = c;
    return SLAS_OK;
}

// Blocks retired by ensure_workspace are only needed by instantiated graphs (and an open capture).
static void free_retired(Context *ctx) {
    if (ctx->retired.empty() || ctx->live_graphs > 0 || ctx->capturing) return;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return; }
    for (void *p : ctx->retired) cudaFree(p);
    ctx->retired.clear();
}

static inline bool spans_overlap(const Context::Span &a, const Context::Span &b) {
    return a.lo < b.hi && b.lo < a.hi;
}

bool hazard_early_ok(const Context *ctx, cudaStream_t st, const Context::Span *reads, int nr, const Context::Span *writes, int nw) {
    const Context::Hazard &h = ctx->hazard;
    // Nothing recorded, another stream, or some other kernel of the library launched since: what may still be running
    // is unknown.  (Copies / memsets / foreign kernels in between do not trigger early, so the dependent launch
    // degenerates to a full dependency -- the barrier kernel this returns false for costs nothing extra then.)
    if (h.n == 0 || h.n >= Context::Hazard::kWindow || h.seq != g_launches.load(std::memory_order_relaxed) || h.stream != st ||
        tunable("pdl") == 1)
        return false;
    for (int o = 0; o < h.n; ++o) {
        const Context::HazardOp &op = h.ops[o];
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < 2; ++j)
                if (spans_overlap(reads[i], op.writes[j])) return false;  // RAW
        for (int i = 0; i < nw; ++i)
            for (int j = 0; j < 2; ++j)
                if (spans_overlap(writes[i], op.writes[j]) || spans_overlap(writes[i], op.reads[j])) return false;  // WAW, WAR
    }
    return true;
}

void hazard_record(Context *ctx, cudaStream_t st, const Context::Span *reads, int nr, const Context::Span *writes, int nw, int early) {
    Context::Hazard &h = ctx->hazard;
    if (!early || h.n >= Context::Hazard::kWindow) h.n = 0;  // a barrier kernel: everything before it is done when its successors start
    Context::HazardOp &op = h.ops[h.n++];
    const Context::Span none = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i) {
        op.reads[i] = i < nr ? reads[i] : none;
        op.writes[i] = i < nw ? writes[i] : none;
    }
    h.seq = g_launches.load(std::memory_order_relaxed);
    h.stream = st;
}

int ensure_workspace(Context *ctx, size_t bytes, void **out) {
    if (bytes > ctx->workspace_bytes) {
        // Kernels already enqueued -- and kernel nodes of recorded graphs -- hold the old address.  While a capture
        // is open (no synchronising call is legal there) or a graph is alive the old block is retired, not freed.
        if (ctx->workspace) {
            if (ctx->capturing || ctx->live_graphs > 0) {
                ctx->retired.push_back(ctx->workspace);
            } else {
                SLAS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                for (int i = 0; i < 3; ++i) SLAS_CUDA_TRY(cudaStreamSynchronize(ctx->copy_streams[i]));
                SLAS_CUDA_TRY(cudaFree(ctx->workspace));
            }
            ctx->workspace = nullptr;
            ctx->workspace_bytes = 0;
        }
        size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        SLAS_CUDA_TRY(cudaMalloc(&ctx->workspace, want));  // legal during a relaxed-mode capture
        ctx->workspace_bytes = want;
    }
    *out = ctx->workspace;
    return SLAS_OK;
}

int ensure_staging(Context *ctx, size_t bytes, void **out) {
    // host operands need synchronising copies: they cannot be part of a recorded graph (include/slas_cuda.h)
    if (ctx->capturing)
        return fail(SLAS_ERR_INVALID, "host operands cannot be recorded in a graph: pass device pointers between "
                                      "slas_cuda_graph_begin and slas_cuda_graph_end");
    if (bytes > ctx->staging_bytes) {
        if (ctx->staging) {
            SLAS_CUDA_TRY(cudaDeviceSynchronize());
            SLAS_CUDA_TRY(cudaFree(ctx->staging));
            ctx->staging = nullptr;
            ctx->staging_bytes = 0;
        }
        size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        SLAS_CUDA_TRY(cudaMalloc(&ctx->staging, want));
        ctx->staging_bytes = want;
    }
    *out = ctx->staging;
    return SLAS_OK;
}

int pointer_is_device(const void *p, bool *is_dev) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SLAS_ERR_CUDA, "cudaPointerGetAttributes: %s", cudaGetErrorString(e));
    }
    *is_dev = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    return SLAS_OK;
}

int operands_side(const void *const *ptrs, int n, bool *is_dev) {
    bool first = false;
    for (int i = 0; i < n; ++i) {
        if (!ptrs[i]) return fail(SLAS_ERR_INVALID, "null data pointer (operand %d)", i);
        bool d = false;
        SLAS_TRY(pointer_is_device(ptrs[i], &d));
        if (i == 0) first = d;
        else if (d != first)
            return fail(SLAS_ERR_INVALID, "operands mix host and device pointers (operand %d)", i);
    }
    *is_dev = first;
    return SLAS_OK;
}

constexpr uint32_t kHostResultSentinel = 0x7fc5a5a5u;  // a quiet NaN with a payload no arithmetic produces

void *host_result_arm(Context *ctx, size_t bytes) {
    volatile uint32_t *slot = reinterpret_cast<volatile uint32_t *>(ctx->pinned_scalars + 8);
    for (size_t w = 0; w < bytes / 4; ++w) slot[w] = kHostResultSentinel;
    __sync_synchronize();
    return ctx->pinned_scalars_dev + 8;
}

int host_result_wait(Context *ctx, cudaStream_t st, void *out, size_t elem_bytes, int count) {
    volatile uint32_t *slot = reinterpret_cast<volatile uint32_t *>(ctx->pinned_scalars + 8);
    const size_t wpe = elem_bytes / 4;  // words per element: 1 (f32) or 2 (f64); each element is ONE store of the kernel
    for (unsigned spins = 0;; ++spins) {
        bool all = true;
        for (int e = 0; e < count && all; ++e) {
            bool changed = false;  // (one word of a 64-bit result may legitimately equal the sentinel)
            for (size_t w = 0; w < wpe; ++w) changed = changed || slot[e * wpe + w] != kHostResultSentinel;
            all = changed;
        }
        if (all) break;
        if ((spins & 0x3ff) == 0x3ff) {
            // a result that IS the sentinel, or a kernel that died: the stream tells
            const cudaError_t q = cudaStreamQuery(st);
            if (q == cudaSuccess) break;  // drained: whatever the kernel wrote is in the slot
            if (q != cudaErrorNotReady) {
                cudaGetLastError();
                return fail(SLAS_ERR_CUDA, "reduction failed: %s", cudaGetErrorString(q));
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    __sync_synchronize();
    memcpy(out, (const void *)slot, elem_bytes * (size_t)count);
    return SLAS_OK;
}

int deliver_scalar(Context *ctx, const void *dev_src, void *out, size_t bytes) {
    if (!out) return fail(SLAS_ERR_INVALID, "null result pointer");
    bool d = false;
    SLAS_TRY(pointer_is_device(out, &d));
    if (d) {
        SLAS_CUDA_TRY(cudaMemcpyAsync(out, dev_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return SLAS_OK;
    }
    if (ctx->capturing)
        return fail(SLAS_ERR_INVALID, "a host result pointer needs a synchronising copy and cannot be recorded in a graph: "
                                      "pass a device pointer for the result");
    SLAS_CUDA_TRY(cudaMemcpyAsync(ctx->pinned_scalars, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SLAS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->pinned_scalars, bytes);
    return SLAS_OK;
}

}  // namespace slas

using namespace slas;

extern "C" {

int slas_cuda_init(int device) {
    if (device >= 0) SLAS_CUDA_TRY(cudaSetDevice(device));
    Context *c;
    return get_context(&c);
}

int slas_cuda_set_tunable(const char *name, long value) {
    SLAS_REQUIRE(name, "null tunable name");
    for (auto &t : g_tunables)
        if (strcmp(t.name, name) == 0) { t.value = value; return SLAS_OK; }
    return fail(SLAS_ERR_INVALID, "unknown tunable '%s'", name);
}
int slas_cuda_get_tunable(const char *name, long *value) {
    SLAS_REQUIRE(name && value, "null argument");
    for (auto &t : g_tunables)
        if (strcmp(t.name, name) == 0) { *value = t.value; return SLAS_OK; }
    return fail(SLAS_ERR_INVALID, "unknown tunable '%s'", name);
}

int slas_cuda_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    int cur = -1;
    cudaGetDevice(&cur);
    for (int d = 0; d < kMaxDevices; ++d) {
        Context *c = &g_ctx[d];
        if (c->device.load(std::memory_order_acquire) != d) continue;
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        std::lock_guard<std::mutex> ck(c->mu);
        c->device.store(-1, std::memory_order_release);
        release_context(c);
    }
    if (cur >= 0) cudaSetDevice(cur);
    return SLAS_OK;
}

const char *slas_cuda_last_error(void) { return t_error; }

int slas_cuda_device_count(int *count) {
    SLAS_REQUIRE(count, "null count");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return fail(SLAS_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return SLAS_OK;
}

int slas_cuda_device_info(int *sm_count, size_t *l2_bytes, size_t *hbm_bytes, int *cc_major, int *cc_minor) {
    Context *c;
    SLAS_TRY(get_context(&c));
    if (sm_count) *sm_count = c->sm_count;
    if (l2_bytes) *l2_bytes = c->l2_bytes;
    if (hbm_bytes) *hbm_bytes = c->hbm_bytes;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    return SLAS_OK;
}

int slas_cuda_set_stream(void *cuda_stream) {
    Context *c;
    SLAS_TRY(get_context(&c));
    std::lock_guard<std::mutex> lk(c->mu);
    SLAS_REQUIRE(!c->capturing, "set_stream inside graph_begin / graph_end");
    cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    if (next != c->stream) {
        // ops share the context's scratch (reduction slots, workspace): work queued on the new stream must not
        // overtake what is still running on the old one
        SLAS_CUDA_TRY(cudaEventRecord(c->copy_events[0], c->stream));
        SLAS_CUDA_TRY(cudaStreamWaitEvent(next, c->copy_events[0], 0));
        c->stream = next;
    }
    return SLAS_OK;
}

int slas_cuda_synchronize(void) {
    Context *c;
    SLAS_TRY(get_context(&c));
    cudaStream_t st;
    { std::lock_guard<std::mutex> lk(c->mu); st = c->stream; }
    SLAS_CUDA_TRY(cudaStreamSynchronize(st));
    return SLAS_OK;
}

uint64_t slas_cuda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- CUDA graphs: a launch-bound sequence of device-pointer ops recorded once, replayed with one launch ----------
int slas_cuda_graph_begin(void) {
    Context *c;
    SLAS_TRY(get_context(&c));
    std::lock_guard<std::mutex> lk(c->mu);
    SLAS_REQUIRE(!c->capturing, "graph_begin: a capture is already open on this device");
    SLAS_CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    c->hazard.n = 0;  // the first recorded kernel is a barrier kernel: a replay may follow anything
    return SLAS_OK;
}

int slas_cuda_graph_end(void **graph_exec) {
    SLAS_REQUIRE(graph_exec, "null graph_exec");
    Context *c;
    SLAS_TRY(get_context(&c));
    std::lock_guard<std::mutex> lk(c->mu);
    SLAS_REQUIRE(c->capturing, "graph_end without graph_begin");
    c->capturing = false;
    c->hazard.n = 0;  // recorded kernels have not run: nothing of them is in flight
    cudaGraph_t graph = nullptr;
    SLAS_CUDA_TRY(cudaStreamEndCapture(c->stream, &graph));
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    SLAS_CUDA_TRY(e);
    ++c->live_graphs;  // from here on the scratch blocks its kernel nodes point at must stay allocated
    *graph_exec = exec;
    return SLAS_OK;
}

int slas_cuda_graph_launch(void *graph_exec) {
    SLAS_REQUIRE(graph_exec, "null graph_exec");
    Context *c;
    SLAS_TRY(get_context(&c));
    std::lock_guard<std::mutex> lk(c->mu);
    SLAS_CUDA_TRY(cudaGraphLaunch((cudaGraphExec_t)graph_exec, c->stream));
    return SLAS_OK;
}

int slas_cuda_graph_destroy(void *graph_exec) {
    if (!graph_exec) return SLAS_OK;
    Context *c;
    SLAS_TRY(get_context(&c));
    std::lock_guard<std::mutex> lk(c->mu);
    // (a garbage-collected handle may be destroyed while another capture is open: no synchronising call then; the
    // driver defers the destruction of an exec that is still running, and retired scratch is freed later)
    SLAS_CUDA_TRY(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    if (c->live_graphs > 0) --c->live_graphs;
    free_retired(c);
    return SLAS_OK;
}

int slas_cuda_malloc(void **dptr, size_t bytes) {
    SLAS_REQUIRE(dptr, "null dptr");
    Context *c;
    SLAS_TRY(get_context(&c));
    SLAS_CUDA_TRY(cudaMalloc(dptr, bytes ? bytes : 16));
    return SLAS_OK;
}

int slas_cuda_free(void *dptr) {
    if (!dptr) return SLAS_OK;
    SLAS_CUDA_TRY(cudaFree(dptr));
    return SLAS_OK;
}

// CPUs of the NUMA node the device's PCIe root port hangs off (sysfs), or false when the platform does not say.
static bool device_local_cpus(int device, cpu_set_t *set) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return false;
    for (char *p = bus; *p; ++p)
        if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return false;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return false;
    char list[4096] = {0};
    const bool ok = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int count = 0;
    char *save = nullptr;
    for (char *tok = strtok_r(list, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {
        int lo = 0, hi = 0;
        const int n = sscanf(tok, "%d-%d", &lo, &hi);
        if (n < 1) continue;
        if (n == 1) hi = lo;
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c) { CPU_SET(c, set); ++count; }
    }
    return count > 0;
}

// Pinned host memory on the NUMA node next to the current device: the pages are placed by the allocating thread
// (first touch inside the driver), so the thread is moved onto the device-local CPUs for the duration of the call.
// With one process per GPU on a two-socket host this is what keeps eight ranks' H2D/D2H traffic off the
// inter-socket link.  SLAS_CUDA_HOST_NUMA=0 turns it off.
int slas_cuda_malloc_host(void **hptr, size_t bytes) {
    SLAS_REQUIRE(hptr, "null hptr");
    Context *c;
    SLAS_TRY(get_context(&c));
    cpu_set_t before, local;
    bool moved = false;
    const char *env = getenv("SLAS_CUDA_HOST_NUMA");
    if (!(env && env[0] == '0') && sched_getaffinity(0, sizeof(before), &before) == 0 && device_local_cpus(c->device.load(), &local)) {
        cpu_set_t want;
        CPU_AND(&want, &before, &local);  // never leave the cpuset the process was given
        if (CPU_COUNT(&want) > 0) moved = sched_setaffinity(0, sizeof(want), &want) == 0;
    }
    const cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 16);
    if (moved) sched_setaffinity(0, sizeof(before), &before);
    SLAS_CUDA_TRY(e);
    return SLAS_OK;
}

// Page-lock memory the caller already owns (a Vec / Box / numpy array), in place: from then on it is a pinned operand
// -- copied from directly by the DMA engines instead of through the bounce ring.  The caller unregisters before freeing.
int slas_cuda_host_register(void *hptr, size_t bytes) {
    SLAS_REQUIRE(hptr && bytes, "host_register: null pointer or zero size");
    Context *c;
    SLAS_TRY(get_context(&c));
    SLAS_CUDA_TRY(cudaHostRegister(hptr, bytes, cudaHostRegisterPortable));
    return SLAS_OK;
}
int slas_cuda_host_unregister(void *hptr) {
    if (!hptr) return SLAS_OK;
    SLAS_CUDA_TRY(cudaHostUnregister(hptr));
    return SLAS_OK;
}

int slas_cuda_free_host(void *hptr) {
    if (!hptr) return SLAS_OK;
    SLAS_CUDA_TRY(cudaFreeHost(hptr));
    return SLAS_OK;
}

static int copy_sync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
    if (bytes == 0) return SLAS_OK;
    SLAS_REQUIRE(dst && src, "null pointer in memcpy");
    DevScope scope(kind == cudaMemcpyHostToDevice ? dst : src);  // several devices: the stream of the one that owns the memory
    Context *c;
    SLAS_TRY(get_context(&c));
    cudaStream_t st;
    { std::lock_guard<std::mutex> lk(c->mu); st = c->stream; }
    SLAS_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, kind, st));
    SLAS_CUDA_TRY(cudaStreamSynchronize(st));
    return SLAS_OK;
}

int slas_cuda_memcpy_h2d(void *dst, const void *src, size_t bytes) {
    return copy_sync(dst, src, bytes, cudaMemcpyHostToDevice);
}
int slas_cuda_memcpy_d2h(void *dst, const void *src, size_t bytes) {
    return copy_sync(dst, src, bytes, cudaMemcpyDeviceToHost);
}
int slas_cuda_memcpy_d2d(void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return SLAS_OK;
    SLAS_REQUIRE(dst && src, "null pointer in memcpy");
    DevScope scope(dst);
    Context *c;
    SLAS_TRY(get_context(&c));
    std::lock_guard<std::mutex> lk(c->mu);
    SLAS_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return SLAS_OK;
}
int slas_cuda_memset(void *dst, int byte, size_t bytes) {
    if (bytes == 0) return SLAS_OK;
    SLAS_REQUIRE(dst, "null pointer in memset");
    DevScope scope(dst);
    Context *c;
    SLAS_TRY(get_context(&c));
    std::lock_guard<std::mutex> lk(c->mu);
    SLAS_CUDA_TRY(cudaMemsetAsync(dst, byte, bytes, c->stream));
    return SLAS_OK;
}

int slas_cuda_is_device_ptr(const void *p, int *is_device) {
    SLAS_REQUIRE(p && is_device, "null pointer");
    bool d = false;
    SLAS_TRY(pointer_is_device(p, &d));
    *is_device = d ? 1 : 0;
    return SLAS_OK;
}

}  // extern "C"
