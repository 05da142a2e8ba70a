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
        for (int i = 0; i < kBounceSlots; ++i) ctx->bounce_pending[i] = false;  // (the device was synchronised above)
        for (int i = 0; i < kBounceSlots; ++i)
            if (!ctx->bounce_events[i]) SLAS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->bounce_events[i], cudaEventDisableTiming));
    }
    *out = (char *)ctx->bounce;
    return SLAS_OK;
}

// Whole-operand copies (the single-object entries).  Pinned source: one asynchronous copy.  Pageable source: chunks of
// kBounceChunk through the ring, the copy threads filling the next slot while the DMA engine drains the previous ones.
// Slots are handed out round-robin across calls and operands; a slot is waited for only when it comes round again, so
// the copies of consecutive operands (a, then b of one dot) overlap each other and nothing is synchronised at the end.
constexpr size_t kBounceChunk = (size_t)16 << 20;

static int bounce_slot(Context *ctx, char *ring, char **slot_ptr, int *slot_idx) {
    const int slot = (int)(ctx->bounce_next++ % kBounceSlots);
    if (ctx->bounce_pending[slot]) {
        SLAS_CUDA_TRY(cudaEventSynchronize(ctx->bounce_events[slot]));  // its last DMA is done
        ctx->bounce_pending[slot] = false;
    }
    *slot_ptr = ring + (size_t)slot * kBounceChunk;
    *slot_idx = slot;
    return SLAS_OK;
}

int copy_h2d(Context *ctx, cudaStream_t st, void *dev, const void *host, size_t bytes) {
    if (bytes == 0) return SLAS_OK;
    if (bytes < ((size_t)1 << 20) || !host_pointer_pageable(host)) {
        SLAS_CUDA_TRY(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st));
        return SLAS_OK;
    }
    char *ring = nullptr;
    SLAS_TRY(ensure_bounce(ctx, kBounceChunk * kBounceSlots, &ring));
    for (size_t off = 0; off < bytes; off += kBounceChunk) {
        const size_t cnt = bytes - off < kBounceChunk ? bytes - off : kBounceChunk;
        char *slot_ptr = nullptr;
        int slot = 0;
        SLAS_TRY(bounce_slot(ctx, ring, &slot_ptr, &slot));
        const HostCopyJob job = {slot_ptr, (const char *)host + off, cnt};
        parallel_memcpy(&job, 1);
        SLAS_CUDA_TRY(cudaMemcpyAsync((char *)dev + off, slot_ptr, cnt, cudaMemcpyHostToDevice, st));
        SLAS_CUDA_TRY(cudaEventRecord(ctx->bounce_events[slot], st));
        ctx->bounce_pending[slot] = true;
    }
    return SLAS_OK;
}

int copy_d2h(Context *ctx, cudaStream_t st, void *host, const void *dev, size_t bytes) {
    if (bytes == 0) return SLAS_OK;
    if (bytes < ((size_t)1 << 20) || !host_pointer_pageable(host)) {
        SLAS_CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
        return SLAS_OK;
    }
    char *ring = nullptr;
    SLAS_TRY(ensure_bounce(ctx, kBounceChunk * kBounceSlots, &ring));
    const size_t nchunks = (bytes + kBounceChunk - 1) / kBounceChunk;
    // DMA runs up to kBounceSlots - 1 chunks ahead of the copy threads
    char *slot_ptrs[kBounceSlots];
    int slots[kBounceSlots];
    size_t issued = 0;
    for (size_t done = 0; done < nchunks; ++done) {
        while (issued < nchunks && issued < done + kBounceSlots - 1) {
            const size_t off = issued * kBounceChunk, cnt = bytes - off < kBounceChunk ? bytes - off : kBounceChunk;
            const int q = (int)(issued % kBounceSlots);
            SLAS_TRY(bounce_slot(ctx, ring, &slot_ptrs[q], &slots[q]));
            SLAS_CUDA_TRY(cudaMemcpyAsync(slot_ptrs[q], (const char *)dev + off, cnt, cudaMemcpyDeviceToHost, st));
            SLAS_CUDA_TRY(cudaEventRecord(ctx->bounce_events[slots[q]], st));
            ctx->bounce_pending[slots[q]] = true;
            ++issued;
        }
        const int q = (int)(done % kBounceSlots);
        const size_t off = done * kBounceChunk, cnt = bytes - off < kBounceChunk ? bytes - off : kBounceChunk;
        SLAS_CUDA_TRY(cudaEventSynchronize(ctx->bounce_events[slots[q]]));
        ctx->bounce_pending[slots[q]] = false;
        const HostCopyJob job = {(char *)host + off, slot_ptrs[q], cnt};
        parallel_memcpy(&job, 1);
    }
    return SLAS_OK;
}

}  // namespace slas
