// common.cuh -- shared plumbing of libslas_cuda: status/error handling, the
// per-process context (device, stream, scratch), pointer classification and small
// device helpers.  No reference code here: the reference has no device side at all.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/slas_cuda.h"

namespace slas {

// ---- errors ------------------------------------------------------------------
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);

#define SLAS_CUDA_TRY(expr)                                                                 \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            cudaGetLastError(); /* do not leave it behind for the next launch check */      \
            return ::slas::fail(SLAS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,              \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                \
        }                                                                                   \
    } while (0)

#define SLAS_TRY(expr)                  \
    do {                                \
        int _s = (expr);                \
        if (_s != SLAS_OK) return _s;   \
    } while (0)

#define SLAS_REQUIRE(cond, ...)                                        \
    do {                                                               \
        if (!(cond)) return ::slas::fail(SLAS_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// ---- context -------------------------------------------------------------------
struct Context {
    std::atomic<int> device{-1};         // -1 until init_context has finished (acquire / release: see get_context)
    int sm_count = 0;
    size_t l2_bytes = 0;
    size_t hbm_bytes = 0;
    int cc_major = 0, cc_minor = 0;
    cudaStream_t own_stream = nullptr;   // created non-blocking at init
    cudaStream_t stream = nullptr;       // the stream ops are enqueued on
    // Reduction scratch: partial sums + self-resetting ticket + device scalars.
    double *partials = nullptr;          // [kMaxPartials * 2]
    unsigned int *ticket = nullptr;      // [4]
    double *scalars = nullptr;           // [16] device scalars (norm, sharded partial, ...)
    void *ll_slots = nullptr;            // [kRedScratchSets][3 * kMaxPartials] 16-byte flagged partial slots (reduce.cuh), zero between launches
    unsigned red_seq = 0;                // rotates the scratch set: overlapping reductions never share one
    unsigned long long *phases = nullptr;  // [8] %globaltimer stamps of the last reduction (tunable "reduce_phases")
    // Grow-only device workspace (transpose_inplace temp, tcgen05 hi/lo split, host staging).  A block that is
    // outgrown is RETIRED, not freed, while a stream capture is open or an instantiated graph may still hold
    // its address in a kernel argument (slas_cuda_graph_*); retired blocks are freed once no graph is alive.
    void *workspace = nullptr;
    size_t workspace_bytes = 0;
    std::vector<void *> retired;
    int live_graphs = 0;
    bool capturing = false;              // between slas_cuda_graph_begin and _end
    // Host-pointer path: staging area in HBM, copy streams + events, pinned bounce scalars.
    void *staging = nullptr;
    size_t staging_bytes = 0;
    cudaStream_t copy_streams[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t copy_events[3] = {nullptr, nullptr, nullptr};   // kernels of chunk i done (orders the shared scratch)
    double *pinned_scalars = nullptr;    // [16]; mapped: [8..11] is the slot reductions write a host-bound scalar into directly
    double *pinned_scalars_dev = nullptr;  // the same memory as the device sees it
    // Pageable host operands (hostcopy.cu): pinned bounce ring + one event per slot
    void *bounce = nullptr;
    size_t bounce_bytes = 0;
    cudaEvent_t bounce_events[4] = {nullptr, nullptr, nullptr, nullptr};
    bool bounce_pending[4] = {false, false, false, false};  // a DMA out of / into that slot may still be running
    unsigned bounce_next = 0;                                 // round-robin slot of the whole-operand copies
    // Programmatic dependent launch of the launch-bound vector ops (reduce.cuh): what the kernels that may still be
    // running read and write, so that the next one knows whether it may touch its operands before they are done.
    // A kernel that waits for its predecessor BEFORE it lets its successor go ("early" = 0) is a barrier: everything
    // before it has completed by the time anything after it starts.  The window holds that kernel and every "early"
    // kernel launched since; a full window forces the next kernel to be a barrier.
    struct Span { const char *lo, *hi; };
    struct HazardOp { Span reads[2], writes[2]; };
    struct Hazard {
        static constexpr int kWindow = 32;  // == kRedScratchSets: a kernel of the window owns one scratch set
        uint64_t seq = ~0ull;            // g_launches right after the last recorded launch; anything launched since voids it
        cudaStream_t stream = nullptr;
        HazardOp ops[kWindow];
        int n = 0;
    } hazard;
    std::mutex mu;                       // serialises enqueue + scratch use
};

// True when a kernel that reads `reads` and writes `writes` BEFORE its griddepcontrol.wait may do so while the kernels
// of the window are still running (no RAW, WAR or WAW overlap with any of them, nothing else launched in between).
bool hazard_early_ok(const Context *ctx, cudaStream_t st, const Context::Span *reads, int nr, const Context::Span *writes, int nw);
// Call right after the launch (after SLAS_LAUNCH_CHECK) of a dependent-launch kernel: `reads` / `writes` = everything it
// touches at any time; early = what hazard_early_ok allowed (0: the kernel is a barrier, the window restarts at it).
void hazard_record(Context *ctx, cudaStream_t st, const Context::Span *reads, int nr, const Context::Span *writes, int nw, int early);

constexpr int kMaxPartials = 148 * 16;
constexpr int kRedScratchSets = 32;

// Returns the context of the current device, initialising it on first use.
int get_context(Context **out);
int ensure_workspace(Context *ctx, size_t bytes, void **out);
int ensure_staging(Context *ctx, size_t bytes, void **out);

// hostcopy.cu -- copies that know about pageable host memory (what a slas StaticVec is)
constexpr int kBounceSlots = 4;
struct HostCopyJob { void *dst; const void *src; size_t bytes; };
void parallel_memcpy(const HostCopyJob *jobs, int njobs);           // the copy-thread pool, fork-join
bool host_pointer_pageable(const void *p);                          // neither pinned nor registered
int ensure_bounce(Context *ctx, size_t bytes, char **out);          // grow-only pinned ring
// enqueue-or-stage: pinned memory -> one cudaMemcpyAsync; pageable -> through the bounce ring (h2d returns with every
// chunk enqueued and the ring idle, d2h returns with the bytes in the caller's memory)
int copy_h2d(Context *ctx, cudaStream_t st, void *dev, const void *host, size_t bytes);
int copy_d2h(Context *ctx, cudaStream_t st, void *host, const void *dev, size_t bytes);

// Counts every kernel launch of this library (slas_cuda_launch_count).
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define SLAS_LAUNCH_CHECK()                                                                 \
    do {                                                                                    \
        ::slas::count_launch();                                                             \
        cudaError_t _e = cudaPeekAtLastError();                                             \
        if (_e != cudaSuccess) {                                                            \
            cudaGetLastError();                                                             \
            return ::slas::fail(SLAS_ERR_CUDA, "kernel launch failed: %s (%s:%d)",          \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                \
        }                                                                                   \
    } while (0)

// Pointer side: 1 device, 0 host (pageable or pinned).
int pointer_is_device(const void *p, bool *is_dev);
// Classify all data operands of a call; fails on a mix.
int operands_side(const void *const *ptrs, int n, bool *is_dev);

// A reduction whose scalar goes back to the HOST (`let x = a.dot(&b)`: the reference's own calling pattern) lets the
// kernel store it straight into mapped pinned memory and the host spin on that word, instead of a device scalar + a
// D2H copy + a stream synchronisation.  host_result_arm fills the slot with a sentinel and returns the device view;
// host_result_wait returns once every word has changed (or the stream has drained) and copies the slot to `out`.
void *host_result_arm(Context *ctx, size_t bytes);
int host_result_wait(Context *ctx, cudaStream_t st, void *out, size_t elem_bytes, int count);

// Write a device scalar (value computed on ctx->stream) to `out`, which may be a host or a
// device pointer.  Host: copies through the pinned bounce buffer and synchronises the stream.
int deliver_scalar(Context *ctx, const void *dev_src, void *out, size_t bytes);

// Benchmark knobs (slas_cuda_set_tunable); every knob defaults to 0 = the shipped choice.
long tunable(const char *name);

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers ----------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum in a fixed order (warp butterflies, then warp 0 over the per-warp sums):
// the same launch geometry always adds in the same order => deterministic.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *smem /* [THREADS/32] */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = lane < THREADS / 32 ? smem[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;  // valid in warp 0
}
#endif

}  // namespace slas
