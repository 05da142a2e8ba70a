// hostcopy.cu -- host <-> device copies for the memory a slas StaticVec really is.  What the reference hands a
// backend is PAGEABLE host memory (`[T; LEN]` on the stack / in a Vec, src/static_vec.rs:126, src/dynamic_vec.rs:70-76).
// cudaMemcpyAsync from pageable memory is a synchronous bounce inside the driver on the calling thread (measured here:
// 11.6 GB/s, and no overlap with the kernels or the other direction), so pageable operands go through the library's own
// pinned bounce ring instead: a small pool of copy threads moves the caller's bytes into / out of pinned slots with
// plain memcpy while the DMA engines move the previous slots over PCIe.  Pinned (or registered) operands keep the
// direct cudaMemcpyAsync.
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <thread>
#include <vector>

#include "common.cuh"

namespace slas {

namespace {

// ---- copy-thread pool: fork-join memcpy over fixed-size pieces ------------------------------------------------------
struct CopyPool {
    struct Piece { char *dst; const char *src; size_t bytes; };
    std::vector<std::thread> threads;
    std::mutex mu;              // pool state
    std::mutex call_mu;         // one parallel copy at a time (it uses every thread anyway)
    std::condition_variable cv_work, cv_done;
    std::vector<Piece> pieces;
    std::atomic<size_t> next{0};
    size_t generation = 0;
    int active = 0;
    bool quit = false;

    void worker() {
        size_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return quit || generation != seen; });
                if (quit) return;
                seen = generation;
            }
            drain();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--active == 0) cv_done.notify_all();
            }
        }
    }
    void drain() {
        for (;;) {
            const size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= pieces.size()) return;
            memcpy(pieces[i].dst, pieces[i].src, pieces[i].bytes);
        }
    }
    void start(int n) {
        for (int i = 0; i < n; ++i) threads.emplace_back([this] { worker(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
            cv_work.notify_all();
        }
        for (auto &t : threads)
            if (t.joinable()) t.join();
    }
} g_pool;
std::once_flag g_pool_once;

int pool_threads() {
    const long knob = tunable("host_copy_threads");
    if (knob > 0) return (int)(knob > 64 ? 64 : knob);
    const unsigned hw = std::thread::hardware_concurrency();
    int n = hw ? (int)hw / 2 : 4;
    if (n < 2) n = 2;
    if (n > 8) n = 8;
    return n;
}

}  // namespace

void parallel_memcpy(const HostCopyJob *jobs, int njobs) {
    size_t total = 0;
    for (int j = 0; j < njobs; ++j) total += jobs[j].bytes;
    if (total < ((size_t)1 << 20)) {  // not worth waking anybody
        for (int j = 0; j < njobs; ++j)
            if (jobs[j].bytes) memcpy(jobs[j].dst, jobs[j].src, jobs[j].bytes);
        return;
    }
    std::call_once(g_pool_once, [] { g_pool.start(pool_threads() - 1); });  // the caller is the last copier
    std::lock_guard<std::mutex> call_lk(g_pool.call_mu);
    constexpr size_t kPiece = (size_t)1 << 20;
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        g_pool.pieces.clear();
        for (int j = 0; j < njobs; ++j)
            for (size_t off = 0; off < jobs[j].bytes; off += kPiece)
                g_pool.pieces.push_back({(char *)jobs[j].dst + off, (const char *)jobs[j].src + off,
                                         jobs[j].bytes - off < kPiece ? jobs[j].bytes - off : kPiece});
        g_pool.next.store(0, std::memory_order_relaxed);
        g_pool.active = (int)g_pool.threads.size();
        ++g_pool.generation;
        g_pool.cv_work.notify_all();
    }
    g_pool.drain();
    std::unique_lock<std::mutex> lk(g_pool.mu);
    g_pool.cv_done.wait(lk, [&] { return g_pool.active == 0; });
}

bool host_pointer_pageable(const void *p) {
    if (!p || tunable("host_bounce") == 1) return false;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeUnregistered;
}

int ensure_bounce(Context *ctx, size_t bytes, char **out) {
    if (bytes > ctx->bounce_bytes) {
        if (ctx->bounce) {
            SLAS_CUDA_TRY(cudaDeviceSynchronize());
            SLAS_CUDA_TRY(cudaFreeHost(ctx->bounce));
            ctx->bounce = nullptr;
            ctx->bounce_bytes = 0;
        }
        const size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        SLAS_CUDA_TRY(cudaMallocHost(&ctx->bounce, want));
        ctx->bounce_bytes = want;
        for (int i = 0; i < kBounceSlots; ++i)
            if (!ctx->bounce_events[i]) SLAS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->bounce_events[i], cudaEventDisableTiming));
    }
    *out = (char *)ctx->bounce;
    return SLAS_OK;
}

// Whole-operand copies (the single-object entries).  Pinned source: one asynchronous copy.  Pageable source: chunks of
// kBounceChunk through the ring, the copy threads filling slot i + 1 while the DMA engine drains slot i.
constexpr size_t kBounceChunk = (size_t)16 << 20;

int copy_h2d(Context *ctx, cudaStream_t st, void *dev, const void *host, size_t bytes) {
    if (bytes == 0) return SLAS_OK;
    if (bytes < ((size_t)2 << 20) || !host_pointer_pageable(host)) {
        SLAS_CUDA_TRY(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st));
        return SLAS_OK;
    }
    char *ring = nullptr;
    SLAS_TRY(ensure_bounce(ctx, kBounceChunk * kBounceSlots, &ring));
    size_t i = 0;
    for (size_t off = 0; off < bytes; off += kBounceChunk, ++i) {
        const int slot = (int)(i % kBounceSlots);
        const size_t cnt = bytes - off < kBounceChunk ? bytes - off : kBounceChunk;
        if (i >= (size_t)kBounceSlots) SLAS_CUDA_TRY(cudaEventSynchronize(ctx->bounce_events[slot]));  // its last DMA is done
        const HostCopyJob job = {ring + (size_t)slot * kBounceChunk, (const char *)host + off, cnt};
        parallel_memcpy(&job, 1);
        SLAS_CUDA_TRY(cudaMemcpyAsync((char *)dev + off, job.dst, cnt, cudaMemcpyHostToDevice, st));
        SLAS_CUDA_TRY(cudaEventRecord(ctx->bounce_events[slot], st));
    }
    // the ring may be reused by the next call on another stream: leave it idle
    for (int s = 0; s < kBounceSlots && (size_t)s < i; ++s) SLAS_CUDA_TRY(cudaEventSynchronize(ctx->bounce_events[s]));
    return SLAS_OK;
}

int copy_d2h(Context *ctx, cudaStream_t st, void *host, const void *dev, size_t bytes) {
    if (bytes == 0) return SLAS_OK;
    if (bytes < ((size_t)2 << 20) || !host_pointer_pageable(host)) {
        SLAS_CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
        return SLAS_OK;
    }
    char *ring = nullptr;
    SLAS_TRY(ensure_bounce(ctx, kBounceChunk * kBounceSlots, &ring));
    const size_t nchunks = (bytes + kBounceChunk - 1) / kBounceChunk;
    // DMA runs kBounceSlots - 1 chunks ahead of the copy threads
    size_t issued = 0;
    for (size_t done = 0; done < nchunks; ++done) {
        while (issued < nchunks && issued < done + kBounceSlots) {
            const int slot = (int)(issued % kBounceSlots);
            const size_t off = issued * kBounceChunk, cnt = bytes - off < kBounceChunk ? bytes - off : kBounceChunk;
            SLAS_CUDA_TRY(cudaMemcpyAsync(ring + (size_t)slot * kBounceChunk, (const char *)dev + off, cnt, cudaMemcpyDeviceToHost, st));
            SLAS_CUDA_TRY(cudaEventRecord(ctx->bounce_events[slot], st));
            ++issued;
        }
        const int slot = (int)(done % kBounceSlots);
        const size_t off = done * kBounceChunk, cnt = bytes - off < kBounceChunk ? bytes - off : kBounceChunk;
        SLAS_CUDA_TRY(cudaEventSynchronize(ctx->bounce_events[slot]));
        const HostCopyJob job = {(char *)host + off, ring + (size_t)slot * kBounceChunk, cnt};
        parallel_memcpy(&job, 1);
    }
    return SLAS_OK;
}

}  // namespace slas
