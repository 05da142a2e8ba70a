// hostcopy.cu -- host <-> device copies for the memory a slas StaticVec really is.  What the reference hands a
// backend is PAGEABLE host memory (`[T; LEN]` on the stack / in a Vec, src/static_vec.rs:126, src/dynamic_vec.rs:70-76).
// cudaMemcpyAsync from pageable memory is a synchronous bounce inside the driver on the calling thread (measured here:
// 11.6 GB/s, and no overlap with the kernels or the other direction), so pageable operands go through the library's own
// pinned bounce ring instead: a small pool of copy threads moves the caller's bytes into / out of pinned slots with
// streaming (non-temporal) stores while the DMA engines move the previous slots over PCIe (hostcopy_pool.cpp).
// Pinned (or registered) operands keep the direct cudaMemcpyAsync.
#include <string.h>

#include "common.cuh"

namespace slas {

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
