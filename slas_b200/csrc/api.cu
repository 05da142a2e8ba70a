// api.cu -- the op entry points of the C ABI (include/slas_cuda.h): argument checks that mirror
// the reference's asserts, host/device pointer dispatch, kernel selection.
//
// Device pointers: the op is enqueued on the context's stream and the call returns.
// Host pointers (the reference's own calling convention: StaticVec::as_ptr() is host memory,
// src/static_vec.rs:126): operands are staged into HBM.  Unit-parallel ops (element-wise and
// every *_batched entry) run a 3-slot pipeline -- H2D copy, kernel, D2H copy of consecutive
// chunks on three streams -- so that both PCIe directions and the SMs overlap; everything
// else stages whole operands.  There is no CPU arithmetic path.
#include <string.h>

#include <algorithm>

#include "comm.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "multi.cuh"

namespace slas {

static int g_sgemm_path = 2;  // 0 FFMA only, 1 tcgen05 3xTF32 wherever its tiling applies, 2 auto (default)

// ---- host staging ----------------------------------------------------------------------------
enum Dir { IN = 1, OUT = 2, INOUT = 3 };
struct Operand {
    const void *host;
    size_t unit_bytes;  // bytes per unit (element / object)
    int dir;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Whole-operand staging: sizes[i] bytes each.  Fills dev[i]; caller runs the op on ctx->stream and
// then calls stage_all_finish.  Pageable operands go through the pinned bounce ring (hostcopy.cu).
static int stage_all_begin(Context *ctx, const Operand *ops, const size_t *sizes, int n, void **dev) {
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += align_up(sizes[i], 256);
    void *base = nullptr;
    SLAS_TRY(ensure_staging(ctx, total ? total : 256, &base));
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        dev[i] = (char *)base + off;
        off += align_up(sizes[i], 256);
        if ((ops[i].dir & IN) && sizes[i]) SLAS_TRY(copy_h2d(ctx, ctx->stream, dev[i], ops[i].host, sizes[i]));
    }
    return SLAS_OK;
}
static int stage_all_finish(Context *ctx, const Operand *ops, const size_t *sizes, int n, void *const *dev) {
    for (int i = 0; i < n; ++i)
        if ((ops[i].dir & OUT) && sizes[i]) SLAS_TRY(copy_d2h(ctx, ctx->stream, const_cast<void *>(ops[i].host), dev[i], sizes[i]));
    SLAS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SLAS_OK;
}

// Chunked 3-slot pipeline over `units` independent units.  `granule`: chunk sizes are multiples of
// it (keeps every chunk's device pointers 16-byte aligned and whole objects together).
// Pinned operands: H2D copy, kernel and D2H copy of consecutive chunks on three streams, straight from / to the
// caller's memory.  Pageable operands (what a slas StaticVec is, static_vec.rs:126): the same pipeline with a pinned
// bounce slot per chunk -- the copy threads fill the slot of chunk i + 1 and empty the slot of chunk i - 2 while the
// DMA engines and the SMs work on the chunks in between.
template <typename F>
static int run_chunked(Context *ctx, size_t units, size_t granule, const Operand *ops, int n, F launch) {
    if (units == 0) return SLAS_OK;
    size_t unit_total = 0;
    bool pageable = false;
    for (int i = 0; i < n; ++i) {
        unit_total += ops[i].unit_bytes;
        if (!pageable && ops[i].unit_bytes && host_pointer_pageable(ops[i].host)) pageable = true;
    }
    if (pageable && units * unit_total < ((size_t)4 << 20)) pageable = false;  // small: the driver's own bounce is fine
    // 96 MB slots by default (64 MB through the bounce ring: measured 58.7 ms per 3 GiB step against 66.6 at 24 MB and
    // 99 at 8 MB, profiles/r2e_tune_pageable.log); "host_slot_mb" is the sweep knob (profiles/tune.py e2e / pageable)
    const long slot_mb = tunable("host_slot_mb");
    const size_t kSlotBytes = (size_t)(slot_mb > 0 ? slot_mb : (pageable ? 64 : 96)) << 20;
    size_t chunk = kSlotBytes / (unit_total ? unit_total : 1);
    chunk = chunk / granule * granule;
    if (chunk == 0) chunk = granule;
    if (chunk > units) chunk = align_up(units, granule);
    size_t slot_bytes = 0;
    for (int i = 0; i < n; ++i) slot_bytes += align_up(chunk * ops[i].unit_bytes, 256);
    const int nslots = units > chunk ? 3 : 1;
    void *base = nullptr;
    SLAS_TRY(ensure_staging(ctx, slot_bytes * nslots, &base));
    char *ring = nullptr;
    if (pageable) SLAS_TRY(ensure_bounce(ctx, slot_bytes * nslots, &ring));
    // the copy streams must not overtake work already queued on the op stream (this also drains whatever the
    // whole-operand copies left in flight in the bounce ring, which is laid out differently below)
    SLAS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < kBounceSlots; ++i) ctx->bounce_pending[i] = false;
    struct Pending { bool live = false; size_t first = 0, cnt = 0; } pend[3];
    // bring the results of the chunk that last used `slot` from the ring into the caller's memory
    auto drain_slot = [&](int slot) -> int {
        if (!pend[slot].live) return SLAS_OK;
        SLAS_CUDA_TRY(cudaEventSynchronize(ctx->bounce_events[slot]));
        HostCopyJob jobs[8];
        int nj = 0;
        size_t off = 0;
        for (int i = 0; i < n; ++i) {
            if (ops[i].dir & OUT)
                jobs[nj++] = {(char *)const_cast<void *>(ops[i].host) + pend[slot].first * ops[i].unit_bytes,
                              ring + (size_t)slot * slot_bytes + off, pend[slot].cnt * ops[i].unit_bytes};
            off += align_up(chunk * ops[i].unit_bytes, 256);
        }
        parallel_memcpy(jobs, nj);
        pend[slot].live = false;
        return SLAS_OK;
    };
    size_t done = 0;
    for (size_t ci = 0; done < units; ++ci) {
        const int slot = (int)(ci % nslots);
        cudaStream_t st = ctx->copy_streams[slot];
        const size_t cnt = std::min(chunk, units - done);
        void *dev[8];
        if (pageable) {
            SLAS_TRY(drain_slot(slot));  // also guarantees the slot's last H2D has left the ring
            HostCopyJob jobs[8];
            int nj = 0;
            size_t off = 0;
            for (int i = 0; i < n; ++i) {
                if (ops[i].dir & IN)
                    jobs[nj++] = {ring + (size_t)slot * slot_bytes + off, (const char *)ops[i].host + done * ops[i].unit_bytes,
                                  cnt * ops[i].unit_bytes};
                off += align_up(chunk * ops[i].unit_bytes, 256);
            }
            parallel_memcpy(jobs, nj);
        }
        size_t off = 0;
        for (int i = 0; i < n; ++i) {
            dev[i] = (char *)base + (size_t)slot * slot_bytes + off;
            if (ops[i].dir & IN) {
                const void *src = pageable ? (const void *)(ring + (size_t)slot * slot_bytes + off)
                                           : (const void *)((const char *)ops[i].host + done * ops[i].unit_bytes);
                SLAS_CUDA_TRY(cudaMemcpyAsync(dev[i], src, cnt * ops[i].unit_bytes, cudaMemcpyHostToDevice, st));
            }
            off += align_up(chunk * ops[i].unit_bytes, 256);
        }
        // The kernels of consecutive chunks run on different streams but share the context's scratch (the grow-only
        // workspace: pre-transposed operands, the 3xTF32 split, gemv / segmented-reduction partials; the reduction
        // ticket).  Chunk i + 1's kernels therefore wait for chunk i's -- they would queue for the same SMs anyway;
        // its H2D copy above and chunk i's D2H copy below still overlap with them.
        if (ci > 0) SLAS_CUDA_TRY(cudaStreamWaitEvent(st, ctx->copy_events[(ci - 1) % 3], 0));
        SLAS_TRY(launch(dev, cnt, st));
        SLAS_CUDA_TRY(cudaEventRecord(ctx->copy_events[ci % 3], st));
        off = 0;
        for (int i = 0; i < n; ++i) {
            if (ops[i].dir & OUT) {
                void *dst = pageable ? (void *)(ring + (size_t)slot * slot_bytes + off)
                                     : (void *)((char *)const_cast<void *>(ops[i].host) + done * ops[i].unit_bytes);
                SLAS_CUDA_TRY(cudaMemcpyAsync(dst, dev[i], cnt * ops[i].unit_bytes, cudaMemcpyDeviceToHost, st));
            }
            off += align_up(chunk * ops[i].unit_bytes, 256);
        }
        if (pageable) {
            SLAS_CUDA_TRY(cudaEventRecord(ctx->bounce_events[slot], st));
            pend[slot].live = true;
            pend[slot].first = done;
            pend[slot].cnt = cnt;
        }
        done += cnt;
    }
    if (pageable) {
        // the last chunks, oldest first
        for (int q = 0; q < nslots; ++q) {
            int oldest = -1;
            for (int s = 0; s < nslots; ++s)
                if (pend[s].live && (oldest < 0 || pend[s].first < pend[oldest].first)) oldest = s;
            if (oldest >= 0) SLAS_TRY(drain_slot(oldest));
        }
    }
    for (int s = 0; s < nslots; ++s) SLAS_CUDA_TRY(cudaStreamSynchronize(ctx->copy_streams[s]));
    return SLAS_OK;
}

// ---- element-wise ----------------------------------------------------------------------------
template <typename T>
static int api_ewise(int op, size_t len, const T *a, const T *b, T *c) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    if (len == 0) return SLAS_OK;
    const void *ptrs[3] = {a, b, c};
    bool dev = false;
    SLAS_TRY(operands_side(ptrs, 3, &dev));
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (dev) return launch_ewise<T>(ctx, ctx->stream, op, len, a, b, c);
    const Operand ops[3] = {{a, sizeof(T), IN}, {b, sizeof(T), IN}, {c, sizeof(T), OUT}};
    return run_chunked(ctx, len, 4096, ops, 3, [&](void **d, size_t cnt, cudaStream_t st) {
        return launch_ewise<T>(ctx, st, op, cnt, (const T *)d[0], (const T *)d[1], (T *)d[2]);
    });
}

// ---- reductions -------------------------------------------------------------------------------
// kind: 0 dot, 1 squared norm, 2 complex dotu; `len` counts real scalars.  mode 1 = sqrt.
// sharded: allreduce the f64 partial over the communicator before the typed result is formed.
template <typename T>
static int api_reduce(int kind, int mode, size_t len, const T *a, const T *b, T *out, bool sharded,
                      double *partial_out) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    SLAS_REQUIRE(out || partial_out, "null result pointer");
    const int nres = kind == 2 ? 2 : 1;
    std::lock_guard<std::mutex> lk(ctx->mu);
    double *res = ctx->scalars;                        // [0..1] f64 sums
    T *typed = reinterpret_cast<T *>(ctx->scalars + 4);  // typed result(s)
    const T *da = a, *db = b;
    bool dev = true;
    if (len > 0) {
        const void *ptrs[2] = {a, kind == 1 ? a : b};
        SLAS_TRY(operands_side(ptrs, 2, &dev));
    }
    if (len > 0 && !dev) {
        const Operand ops[2] = {{a, sizeof(T), IN}, {b, sizeof(T), IN}};
        const size_t sizes[2] = {len * sizeof(T), kind == 1 ? 0 : len * sizeof(T)};
        void *d[2];
        SLAS_TRY(stage_all_begin(ctx, ops, sizes, 2, d));
        da = (const T *)d[0];
        db = kind == 1 ? da : (const T *)d[1];
    }
    bool need_collective = sharded && comm_nranks() > 1;
    // a device-resident result pointer is written by the kernel itself: one launch, no copy, no sync
    bool out_dev = false;
    if (out) SLAS_TRY(pointer_is_device(out, &out_dev));
    // a collective's epoch / NCCL call is not replayable: sharded ops stay out of recorded graphs
    if (need_collective && ctx->capturing)
        return fail(SLAS_ERR_INVALID, "sharded reductions cannot be recorded in a graph (the exchange epoch is per call)");
    // fused collective: the reduction kernel itself exchanges the partial over NVLink peer memory.  The epoch is taken
    // last, after every check that can fail, and given back if the launch itself fails: a rank whose epoch ran ahead
    // of its peers would make every later collective spin.
    PeerExchange px;
    const bool fused = need_collective && comm_next_peer_exchange(&px);
    if (fused) need_collective = false;
    if (out_dev && (len > 0 || fused)) typed = out;
    // a scalar bound for the HOST (`let x = a.dot(&b)`): the kernel stores it into mapped pinned memory and the host
    // spins on that word -- no device scalar, no D2H copy, no stream synchronisation
    const bool direct_host = out && !out_dev && len > 0 && !need_collective && !fused && !partial_out && !ctx->capturing &&
                             tunable("host_result") != 1;
    if (direct_host) typed = static_cast<T *>(host_result_arm(ctx, sizeof(T) * nres));
    if (len == 0 && !fused) {
        // the reference returns the empty sum: 0 (and sqrt(0) = 0)
        SLAS_CUDA_TRY(cudaMemsetAsync(ctx->scalars, 0, sizeof(double) * 8, ctx->stream));
    } else {
        // (a rank with an empty slice still takes part in the fused exchange: its partial is 0)
        // the f64 sums go to the context's scalars only when somebody reads them (a collective behind the kernel, the
        // *_partial entries): every reduction writing that one slot would order all of them one behind the other
        double *res_arg = (need_collective || partial_out) ? res : nullptr;
        const int r = launch_reduce<T>(ctx, ctx->stream, kind, len, da, db, res_arg, need_collective ? nullptr : typed, mode, 0,
                                       fused ? &px : nullptr);
        if (r != SLAS_OK) {
            if (fused) comm_rollback_peer_exchange(px.epoch);
            return r;
        }
    }
    if (direct_host) return host_result_wait(ctx, ctx->stream, out, sizeof(T), nres);
    if (need_collective) {
        SLAS_TRY(comm_allreduce_f64(res, nres, ctx->stream));
        SLAS_TRY(launch_finish_scalar<T>(ctx->stream, res, typed, mode));
        if (nres == 2) SLAS_TRY(launch_finish_scalar<T>(ctx->stream, res + 1, typed + 1, mode));
    }
    if (partial_out) SLAS_TRY(deliver_scalar(ctx, res, partial_out, sizeof(double) * nres));
    if (out && typed != out) SLAS_TRY(deliver_scalar(ctx, typed, out, sizeof(T) * nres));
    return SLAS_OK;
}

template <typename T>
static int api_normalize(size_t len, T *a, bool sharded) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    if (len == 0 && !(sharded && comm_nranks() > 1)) return SLAS_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    double *res = ctx->scalars;
    T *typed = reinterpret_cast<T *>(ctx->scalars + 4);
    bool dev = true;
    if (len > 0) SLAS_TRY(pointer_is_device(a, &dev));
    T *da = a;
    const Operand ops[1] = {{a, sizeof(T), INOUT}};
    const size_t sizes[1] = {len * sizeof(T)};
    void *d[1];
    if (!dev) {
        SLAS_TRY(stage_all_begin(ctx, ops, sizes, 1, d));
        da = (T *)d[0];
    }
    bool need_collective = sharded && comm_nranks() > 1;
    if (need_collective && ctx->capturing)
        return fail(SLAS_ERR_INVALID, "sharded reductions cannot be recorded in a graph (the exchange epoch is per call)");
    if (!need_collective && len > 0) {
        // a short vector: one launch, the vector stays in registers between the norm and the division
        bool handled = false;
        SLAS_TRY(launch_normalize_fused<T>(ctx, ctx->stream, len, da, (T *)nullptr, &handled));  // nobody reads the norm itself
        if (handled) {
            if (!dev) SLAS_TRY(stage_all_finish(ctx, ops, sizes, 1, d));
            return SLAS_OK;
        }
    }
    PeerExchange px;
    const bool fused = need_collective && comm_next_peer_exchange(&px);  // after every check that can fail
    if (fused) need_collective = false;
    if (len == 0 && !fused) SLAS_CUDA_TRY(cudaMemsetAsync(ctx->scalars, 0, sizeof(double) * 8, ctx->stream));
    else {
        const int r = launch_reduce<T>(ctx, ctx->stream, 1, len, da, da, res, need_collective ? nullptr : typed, 1, 0,
                                       fused ? &px : nullptr);
        if (r != SLAS_OK) {
            if (fused) comm_rollback_peer_exchange(px.epoch);
            return r;
        }
    }
    if (need_collective) {
        SLAS_TRY(comm_allreduce_f64(res, 1, ctx->stream));
        SLAS_TRY(launch_finish_scalar<T>(ctx->stream, res, typed, 1));
    }
    SLAS_TRY(launch_scale_div<T>(ctx, ctx->stream, len, da, da, typed));
    if (!dev) SLAS_TRY(stage_all_finish(ctx, ops, sizes, 1, d));
    return SLAS_OK;
}

// ---- fused chains ------------------------------------------------------------------------------
// chain 0..3: c = a op b (c may be null: not materialised), out[0] = dot(c, d);  chain 4: out[0..2] = dot(a,b), |a|, |b|
template <typename T>
static int api_fused_chain(int chain, size_t len, const T *a, const T *b, T *c, const T *dd, T *out) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    SLAS_REQUIRE(out, "null result pointer");
    const int nres = chain == 4 ? 3 : 1;
    std::lock_guard<std::mutex> lk(ctx->mu);
    T *typed = reinterpret_cast<T *>(ctx->scalars + 4);
    bool out_dev = false;
    SLAS_TRY(pointer_is_device(out, &out_dev));
    if (len == 0) {
        SLAS_CUDA_TRY(cudaMemsetAsync(typed, 0, sizeof(T) * nres, ctx->stream));
        return deliver_scalar(ctx, typed, out, sizeof(T) * nres);
    }
    const void *ptrs[4] = {a, b, chain == 4 ? a : dd, c ? c : a};
    bool dev = false;
    SLAS_TRY(operands_side(ptrs, 4, &dev));
    if (dev) {
        // device operands, host-bound scalars: stored by the kernel into mapped pinned memory, the host spins (as api_reduce)
        const bool direct_host = !out_dev && !ctx->capturing && tunable("host_result") != 1;
        if (direct_host) typed = static_cast<T *>(host_result_arm(ctx, sizeof(T) * nres));
        SLAS_TRY(launch_fused_chain<T>(ctx, ctx->stream, chain, len, a, b, c, dd, out_dev ? out : typed));
        if (direct_host) return host_result_wait(ctx, ctx->stream, out, sizeof(T), nres);
        return out_dev ? SLAS_OK : deliver_scalar(ctx, typed, out, sizeof(T) * nres);
    }
    const Operand ops[4] = {{a, sizeof(T), IN}, {b, sizeof(T), IN}, {dd, sizeof(T), IN}, {c, sizeof(T), OUT}};
    const size_t sizes[4] = {len * sizeof(T), len * sizeof(T), chain == 4 ? 0 : len * sizeof(T), c ? len * sizeof(T) : 0};
    void *d[4];
    SLAS_TRY(stage_all_begin(ctx, ops, sizes, 4, d));
    SLAS_TRY(launch_fused_chain<T>(ctx, ctx->stream, chain, len, (const T *)d[0], (const T *)d[1], c ? (T *)d[3] : nullptr,
                                   (const T *)d[2], out_dev ? out : typed));
    SLAS_TRY(stage_all_finish(ctx, ops, sizes, 4, d));
    return out_dev ? SLAS_OK : deliver_scalar(ctx, typed, out, sizeof(T) * nres);
}

// out = a / |a| without touching a: the copy-on-write + normalize chain (static_vec.rs:304-318 then rust.rs:122-127)
template <typename T>
static int api_normalize_into(size_t len, const T *a, T *out) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    if (len == 0) return SLAS_OK;
    SLAS_REQUIRE((const void *)a != (const void *)out, "normalize_into: operands alias (use normalize)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    double *res = ctx->scalars;
    T *typed = reinterpret_cast<T *>(ctx->scalars + 4);
    const void *ptrs[2] = {a, out};
    bool dev = false;
    SLAS_TRY(operands_side(ptrs, 2, &dev));
    const Operand ops[2] = {{a, sizeof(T), IN}, {out, sizeof(T), OUT}};
    const size_t sizes[2] = {len * sizeof(T), len * sizeof(T)};
    void *d[2] = {const_cast<T *>(a), out};
    if (!dev) SLAS_TRY(stage_all_begin(ctx, ops, sizes, 2, d));
    SLAS_TRY(launch_reduce<T>(ctx, ctx->stream, 1, len, (const T *)d[0], (const T *)d[0], res, typed, 1, 0, nullptr));
    SLAS_TRY(launch_scale_div<T>(ctx, ctx->stream, len, (const T *)d[0], (T *)d[1], typed));
    if (!dev) SLAS_TRY(stage_all_finish(ctx, ops, sizes, 2, d));
    return SLAS_OK;
}

// ---- transpose ---------------------------------------------------------------------------------
static int api_transpose(size_t batch, size_t len, size_t elem, const void *a, void *buffer, size_t columns,
                         bool inplace) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    SLAS_REQUIRE(columns != 0, "transpose: columns == 0 (the reference divides by it, src/backends/rust.rs:169)");
    SLAS_REQUIRE(elem >= 1, "transpose: element size 0");
    if (len == 0 || batch == 0) return SLAS_OK;
    const void *ptrs[2] = {a, inplace ? a : buffer};
    bool dev = false;
    SLAS_TRY(operands_side(ptrs, 2, &dev));
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t bytes = batch * len * elem;
    if (inplace) {
        // rust.rs:152-160: transpose into a zeroed temporary, then copy the temporary back
        void *tmp = nullptr;
        void *target = const_cast<void *>(a);
        const Operand ops[1] = {{a, len * elem, INOUT}};
        const size_t sizes[1] = {bytes};
        void *d[1];
        if (!dev) {
            SLAS_TRY(stage_all_begin(ctx, ops, sizes, 1, d));
            target = d[0];
        }
        // objects a thread / warp can hold, and square ones (tile pairs swapped across the diagonal), are
        // transposed in place: one read and one write per element
        bool handled = false;
        SLAS_TRY(launch_transpose_inplace(ctx, ctx->stream, batch, len, elem, target, columns, &handled));
        if (!handled) {
            SLAS_TRY(ensure_workspace(ctx, bytes, &tmp));
            if (len % columns || len < columns) SLAS_CUDA_TRY(cudaMemsetAsync(tmp, 0, bytes, ctx->stream));
            SLAS_TRY(launch_transpose(ctx, ctx->stream, batch, len, elem, target, tmp, columns));
            SLAS_CUDA_TRY(cudaMemcpyAsync(target, tmp, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (!dev) SLAS_TRY(stage_all_finish(ctx, ops, sizes, 1, d));
        return SLAS_OK;
    }
    if (dev) return launch_transpose(ctx, ctx->stream, batch, len, elem, a, buffer, columns);
    if (batch > 1) {
        // elements the reference never writes (len % columns != 0) must keep the caller's bytes: INOUT
        const int out_dir = (len % columns) ? INOUT : OUT;
        const Operand ops[2] = {{a, len * elem, IN}, {buffer, len * elem, out_dir}};
        return run_chunked(ctx, batch, 1, ops, 2, [&](void **d, size_t cnt, cudaStream_t st) {
            return launch_transpose(ctx, st, cnt, len, elem, d[0], d[1], columns);
        });
    }
    const Operand ops[2] = {{a, len * elem, IN}, {buffer, len * elem, (len % columns) ? INOUT : OUT}};
    const size_t sizes[2] = {bytes, bytes};
    void *d[2];
    SLAS_TRY(stage_all_begin(ctx, ops, sizes, 2, d));
    SLAS_TRY(launch_transpose(ctx, ctx->stream, 1, len, elem, d[0], d[1], columns));
    return stage_all_finish(ctx, ops, sizes, 2, d);
}

// ---- matrix_mul ---------------------------------------------------------------------------------
template <typename T>
static int gemm_device(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m, size_t n,
                       size_t k, size_t lda, size_t ldb, size_t ldc, int ta, int tb, bool dense) {
    if (dense) {
        const int r = launch_gemm_small_batched<T>(ctx, st, batch, a, b, c, m, n, k, ta, tb);
        if (r != SLAS_ERR_UNSUPPORTED) return r;
    }
    if (batch == 1 && m * n >= 64 * 64 && k >= 8) {
        // auto: the tensor-core path from 2^27 multiply-adds up -- measured 2.3x (512^3) to 5x (4096^3) faster than
        // the FFMA kernel, profiles/r1e_tune_large.log -- unless padding (m, n, k) up to whole tiles would add more
        // than 60 % of work.  Its 3xTF32 result is fp32-class (<= 2e-6 * sum|a||b|, inside the 1e-5 * sqrt(K) bar);
        // the FFMA kernel below is the ascending-k fp32 chain.
        bool tc_auto = false;
        if (sizeof(T) == 4 && g_sgemm_path == 2 && m * n * k >= ((size_t)1 << 27) && tunable("gemm_large_variant") == 0) {
            size_t mp, np, kp;
            bool pair;
            sgemm_3xtf32_padded(m, n, k, &mp, &np, &kp, &pair);
            tc_auto = (double)mp * (double)np * (double)kp <= 1.6 * (double)m * (double)n * (double)k;
        }
        if (sizeof(T) == 4 && (g_sgemm_path == 1 || tc_auto)) {
            const int r = launch_sgemm_3xtf32(ctx, st, (const float *)a, (const float *)b, (float *)c, m, n, k, lda, ldb,
                                              ldc, ta, tb);
            if (r != SLAS_ERR_UNSUPPORTED) return r;
        }
        // The FFMA kernel is fastest when A arrives k-major (a_trans) and B n-major (!b_trans): both tiles then go
        // to shared memory with 128-bit stores (measured 50.6 vs 43.4 TFLOP/s at 4096^3).  For a big enough dense
        // f32 product it pays to put the operands into that form first with the HBM-speed transpose kernels
        // (2 x 64 MB at ~6 TB/s = 0.02 ms against 0.4 ms gained); the FMA order per C element is unchanged.
        if (sizeof(T) == 4 && tunable("gemm_large_variant") != 2 && m * n * k >= ((size_t)1 << 30)) {
            const bool fix_a = !ta && lda == k, fix_b = tb && ldb == k;
            if (fix_a || fix_b) {
                void *ws = nullptr;
                const size_t ea = fix_a ? m * k : 0, eb = fix_b ? n * k : 0;
                SLAS_TRY(ensure_workspace(ctx, (ea + eb) * sizeof(T), &ws));
                T *at = (T *)ws, *bt = at + ea;
                if (fix_a) SLAS_TRY(launch_transpose(ctx, st, 1, m * k, sizeof(T), a, at, m));  // [m][k] -> [k][m]
                if (fix_b) SLAS_TRY(launch_transpose(ctx, st, 1, n * k, sizeof(T), b, bt, n));  // [n][k] -> [k][n]
                return launch_gemm_large<T>(ctx, st, fix_a ? at : a, fix_b ? bt : b, c, m, n, k, fix_a ? m : lda,
                                            fix_b ? n : ldb, ldc, fix_a ? 1 : ta, fix_b ? 0 : tb);
            }
        }
        // The DFMA kernels are the other way round (measured at 4096^3, profiles/r1e_tune_dgemm*.log: row-major A and
        // k-contiguous B 23.8 TFLOP/s, row-major B 22.3, transposed A 14.3 on the 64 x 64 kernel): put both operands
        // into k-contiguous form first (two HBM-speed transposes, ~1 % of the product's time)
        if (sizeof(T) == 8 && m * n * k >= ((size_t)1 << 30) && tunable("gemm_large_variant") != 2) {
            const bool fix_a = ta && lda == m, fix_b = !tb && ldb == n;
            if (fix_a || fix_b) {
                void *ws = nullptr;
                const size_t ea = fix_a ? m * k : 0, eb = fix_b ? n * k : 0;
                SLAS_TRY(ensure_workspace(ctx, (ea + eb) * sizeof(T), &ws));
                T *ar = (T *)ws, *br = ar + ea;
                if (fix_a) SLAS_TRY(launch_transpose(ctx, st, 1, m * k, sizeof(T), a, ar, k));  // [k][m] -> [m][k]
                if (fix_b) SLAS_TRY(launch_transpose(ctx, st, 1, n * k, sizeof(T), b, br, k));  // [k][n] -> [n][k]
                return launch_gemm_large<T>(ctx, st, fix_a ? ar : a, fix_b ? br : b, c, m, n, k, fix_a ? k : lda,
                                            fix_b ? k : ldb, ldc, fix_a ? 0 : ta, fix_b ? 1 : tb);
            }
        }
        return launch_gemm_large<T>(ctx, st, a, b, c, m, n, k, lda, ldb, ldc, ta, tb);
    }
    // a batch of medium-sized products (past the shared-memory staged kernels): the register-tiled kernel, one grid layer
    // per object (100x100 f32, batch 4096: 3.9 TFLOP/s on the plain kernel)
    if (batch > 1 && m * n >= 32 * 32 && k >= 8)
        return launch_gemm_large_batched<T>(ctx, st, batch, a, b, c, m, n, k, lda, ldb, ldc, ta, tb, m * k, k * n, m * n);
    return launch_gemm_generic<T>(ctx, st, batch, a, b, c, m, n, k, lda, ldb, ldc, ta, tb, m * k, k * n, m * n);
}

template <typename T>
static int api_gemm(const T *a, const T *b, T *c, size_t m, size_t n, size_t k, size_t lda, size_t ldb, size_t ldc,
                    int ta, int tb) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    if (m == 0 || n == 0) return SLAS_OK;
    // what cblas_?gemm(RowMajor) requires of the leading dimensions (the reference passes them through)
    const size_t a_rows = ta ? k : m, a_cols = ta ? m : k, b_rows = tb ? n : k, b_cols = tb ? k : n;
    SLAS_REQUIRE(lda >= std::max<size_t>(a_cols, 1), "matrix_mul: lda %zu < %zu", lda, a_cols);
    SLAS_REQUIRE(ldb >= std::max<size_t>(b_cols, 1), "matrix_mul: ldb %zu < %zu", ldb, b_cols);
    SLAS_REQUIRE(ldc >= n, "matrix_mul: ldc %zu < n %zu", ldc, n);
    const bool dense = lda == a_cols && ldb == b_cols && ldc == n;
    const void *ptrs[3] = {c, k ? (const void *)a : (const void *)c, k ? (const void *)b : (const void *)c};
    bool dev = false;
    SLAS_TRY(operands_side(ptrs, 3, &dev));
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (dev) return gemm_device<T>(ctx, ctx->stream, 1, a, b, c, m, n, k, lda, ldb, ldc, ta, tb, dense);
    const size_t ea = a_rows && a_cols ? (a_rows - 1) * lda + a_cols : 0;
    const size_t eb = b_rows && b_cols ? (b_rows - 1) * ldb + b_cols : 0;
    const size_t ec = (m - 1) * ldc + n;
    // C is INOUT when ldc > n: the gaps between rows keep the caller's bytes (beta = 0 only covers m x n)
    const Operand ops[3] = {{a, sizeof(T), IN}, {b, sizeof(T), IN}, {c, sizeof(T), ldc > n ? INOUT : OUT}};
    const size_t sizes[3] = {ea * sizeof(T), eb * sizeof(T), ec * sizeof(T)};
    void *d[3];
    SLAS_TRY(stage_all_begin(ctx, ops, sizes, 3, d));
    SLAS_TRY(gemm_device<T>(ctx, ctx->stream, 1, (const T *)d[0], (const T *)d[1], (T *)d[2], m, n, k, lda, ldb, ldc, ta,
                            tb, dense));
    return stage_all_finish(ctx, ops, sizes, 3, d);
}

template <typename T>
static int api_gemm_batched(size_t batch, const T *a, const T *b, T *c, size_t m, size_t n, size_t k, int ta, int tb) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    if (batch == 0 || m == 0 || n == 0) return SLAS_OK;
    const size_t lda = ta ? m : k, ldb = tb ? k : n, ldc = n;  // src/tensor.rs:376-382 on dense objects
    const void *ptrs[3] = {c, k ? (const void *)a : (const void *)c, k ? (const void *)b : (const void *)c};
    bool dev = false;
    SLAS_TRY(operands_side(ptrs, 3, &dev));
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (dev) return gemm_device<T>(ctx, ctx->stream, batch, a, b, c, m, n, k, lda, ldb, ldc, ta, tb, true);
    const Operand ops[3] = {{a, m * k * sizeof(T), IN}, {b, k * n * sizeof(T), IN}, {c, m * n * sizeof(T), OUT}};
    // granule: keep chunk starts 16-byte aligned for every operand
    size_t granule = 1;
    while ((granule * m * k * sizeof(T)) % 16 || (granule * k * n * sizeof(T)) % 16 || (granule * m * n * sizeof(T)) % 16)
        granule *= 2;
    return run_chunked(ctx, batch, granule, ops, 3, [&](void **d, size_t cnt, cudaStream_t st) {
        return gemm_device<T>(ctx, st, cnt, (const T *)d[0], (const T *)d[1], (T *)d[2], m, n, k, lda, ldb, ldc, ta, tb,
                              true);
    });
}

// ---- matrix_vector_mul ---------------------------------------------------------------------------
template <typename T>
static int api_gemv(size_t batch, const T *a, const T *x, T *y, size_t m, size_t n, size_t lda, int ta, bool batched) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    const size_t xlen = ta ? n : m, ylen = ta ? m : n;
    if (batch == 0 || ylen == 0) return SLAS_OK;
    SLAS_REQUIRE(lda >= std::max<size_t>(m, 1), "matrix_vector_mul: lda %zu < width %zu", lda, m);
    const void *ptrs[3] = {y, xlen ? (const void *)a : (const void *)y, xlen ? (const void *)x : (const void *)y};
    bool dev = false;
    SLAS_TRY(operands_side(ptrs, 3, &dev));
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t ea = n ? (n - 1) * lda + m : 0;
    const size_t sa = batched ? n * m : 0;
    if (dev) return launch_gemv<T>(ctx, ctx->stream, batch, a, x, y, m, n, lda, ta, sa, xlen, ylen);
    if (batched) {
        const Operand ops[3] = {{a, n * m * sizeof(T), IN}, {x, xlen * sizeof(T), IN}, {y, ylen * sizeof(T), OUT}};
        return run_chunked(ctx, batch, 1, ops, 3, [&](void **d, size_t cnt, cudaStream_t st) {
            return launch_gemv<T>(ctx, st, cnt, (const T *)d[0], (const T *)d[1], (T *)d[2], m, n, lda, ta, sa, xlen, ylen);
        });
    }
    const Operand ops[3] = {{a, sizeof(T), IN}, {x, sizeof(T), IN}, {y, sizeof(T), OUT}};
    const size_t sizes[3] = {ea * sizeof(T), xlen * sizeof(T), ylen * sizeof(T)};
    void *d[3];
    SLAS_TRY(stage_all_begin(ctx, ops, sizes, 3, d));
    SLAS_TRY(launch_gemv<T>(ctx, ctx->stream, 1, (const T *)d[0], (const T *)d[1], (T *)d[2], m, n, lda, ta, 0, 0, 0));
    return stage_all_finish(ctx, ops, sizes, 3, d);
}

// ---- batched reductions -----------------------------------------------------------------------------
template <typename T>
static int api_batched_reduce(int mode, size_t batch, size_t len, const T *a, const T *b, T *out, T *a_mut) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    if (batch == 0) return SLAS_OK;
    if (mode == 2 && len == 0) return SLAS_OK;
    const T *base = mode == 2 ? a_mut : a;
    const void *anchor = mode == 2 ? (const void *)a_mut : (const void *)out;
    const void *ptrs[3] = {anchor, len ? (const void *)base : anchor, (mode == 0 && len) ? (const void *)b : anchor};
    bool dev = false;
    SLAS_TRY(operands_side(ptrs, 3, &dev));
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (dev) return launch_batched_reduce<T>(ctx, ctx->stream, mode, batch, len, a, b, out, a_mut);
    // granule 4: out chunks (4 bytes x 4 = 16) stay aligned
    const size_t granule = 4;
    if (mode == 0) {
        const Operand ops[3] = {{a, len * sizeof(T), IN}, {b, len * sizeof(T), IN}, {out, sizeof(T), OUT}};
        return run_chunked(ctx, batch, granule, ops, 3, [&](void **d, size_t cnt, cudaStream_t st) {
            return launch_batched_reduce<T>(ctx, st, 0, cnt, len, (const T *)d[0], (const T *)d[1], (T *)d[2], nullptr);
        });
    }
    if (mode == 1) {
        const Operand ops[2] = {{a, len * sizeof(T), IN}, {out, sizeof(T), OUT}};
        return run_chunked(ctx, batch, granule, ops, 2, [&](void **d, size_t cnt, cudaStream_t st) {
            return launch_batched_reduce<T>(ctx, st, 1, cnt, len, (const T *)d[0], nullptr, (T *)d[1], nullptr);
        });
    }
    const Operand ops[1] = {{a_mut, len * sizeof(T), INOUT}};
    size_t g2 = 1;
    while ((g2 * len * sizeof(T)) % 16) g2 *= 2;
    return run_chunked(ctx, batch, g2, ops, 1, [&](void **d, size_t cnt, cudaStream_t st) {
        return launch_batched_reduce<T>(ctx, st, 2, cnt, len, nullptr, nullptr, nullptr, (T *)d[0]);
    });
}

template <typename T>
static int api_fill(size_t len, T *x, uint64_t seed, uint64_t offset, T lo, T hi) {
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    if (len == 0) return SLAS_OK;
    bool dev = false;
    SLAS_TRY(pointer_is_device(x, &dev));
    SLAS_REQUIRE(dev, "fill_uniform writes device memory only");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return launch_fill_uniform<T>(ctx, ctx->stream, len, x, seed, offset, lo, hi);
}


// ---- several devices in one process (multi.cuh): host operands are split over them -----------------------------------
// units [lo, lo + cnt) of `units`, in multiples of `granule`, one range per device, all devices at once
template <typename F>
static int split_units(size_t units, size_t granule, F fn) {
    const int n = multi_ndev();
    return multi_run(n, [&](int d) {
        const size_t lo = units * (size_t)d / (size_t)n / granule * granule;
        const size_t hi = d == n - 1 ? units : units * (size_t)(d + 1) / (size_t)n / granule * granule;
        return hi > lo ? fn(lo, hi - lo) : (int)SLAS_OK;
    });
}

template <typename T>
static int entry_ewise(int op, size_t len, const T *a, const T *b, T *c) {
    DevScope scope(c);
    if (multi_split_wanted(c, 3 * len * sizeof(T)))
        return split_units(len, 4096, [&](size_t lo, size_t cnt) { return api_ewise<T>(op, cnt, a + lo, b + lo, c + lo); });
    return api_ewise<T>(op, len, a, b, c);
}

// dot / norm / dotu of HOST vectors over several devices: slice per device, f64 partials added in device order here
template <typename T>
static int entry_reduce(int kind, int mode, size_t len, const T *a, const T *b, T *out) {
    DevScope scope(a);
    const size_t unit = kind == 2 ? 2 : 1;  // complex: whole (re, im) pairs per slice
    if (!(out && multi_split_wanted(a, (kind == 1 ? 1 : 2) * len * sizeof(T)))) return api_reduce<T>(kind, mode, len, a, b, out, false, nullptr);
    const int n = multi_ndev();
    double parts[32][2] = {};
    SLAS_TRY(split_units(len / unit, 16, [&](size_t lo, size_t cnt) {
        int cur = 0;
        cudaGetDevice(&cur);  // worker d runs with device d current: parts[] is in slice order
        return api_reduce<T>(kind, mode, cnt * unit, a + lo * unit, (kind == 1 ? a : b) + lo * unit, nullptr, false, parts[cur]);
    }));
    double s0 = 0.0, s1 = 0.0;
    for (int d = 0; d < n; ++d) { s0 += parts[d][0]; s1 += parts[d][1]; }
    bool out_dev = false;
    SLAS_TRY(pointer_is_device(out, &out_dev));
    T res[2] = {mode == 1 ? (T)sqrt(s0) : (T)s0, (T)s1};
    if (out_dev) return slas_cuda_memcpy_h2d(out, res, sizeof(T) * (kind == 2 ? 2 : 1));
    out[0] = res[0];
    if (kind == 2) out[1] = res[1];
    return SLAS_OK;
}

// normalize of a HOST vector over several devices: every device stages its slice once, the workers meet to add the
// partial sums of squares in device order, then each scales its own slice (true division by the same norm bits)
template <typename T>
static int entry_normalize(size_t len, T *a) {
    DevScope scope(a);
    if (!multi_split_wanted(a, 2 * len * sizeof(T))) return api_normalize<T>(len, a, false);
    const int n = multi_ndev();
    double parts[32] = {};
    Rendezvous meet(n);
    return multi_run(n, [&](int d) {
        const size_t lo = len * (size_t)d / (size_t)n / 16 * 16, hi = d == n - 1 ? len : len * (size_t)(d + 1) / (size_t)n / 16 * 16;
        const size_t cnt = hi - lo;
        Context *ctx = nullptr;
        int st = get_context(&ctx);
        std::unique_lock<std::mutex> lk;
        void *dv[1] = {nullptr};
        const Operand ops[1] = {{a + lo, sizeof(T), INOUT}};
        const size_t sizes[1] = {cnt * sizeof(T)};
        if (st == SLAS_OK) lk = std::unique_lock<std::mutex>(ctx->mu);
        if (st == SLAS_OK && cnt) st = stage_all_begin(ctx, ops, sizes, 1, dv);
        if (st == SLAS_OK && cnt) st = launch_reduce<T>(ctx, ctx->stream, 1, cnt, (const T *)dv[0], (const T *)dv[0], ctx->scalars, nullptr, 1, 0, nullptr);
        if (st == SLAS_OK && cnt) st = deliver_scalar(ctx, ctx->scalars, &parts[d], sizeof(double));
        if (!meet.arrive(st == SLAS_OK)) return st != SLAS_OK ? st : fail(SLAS_ERR_CUDA, "normalize: another device failed");
        if (!cnt) return (int)SLAS_OK;
        double sum = 0.0;
        for (int i = 0; i < n; ++i) sum += parts[i];
        const T nrm = (T)sqrt(sum);
        T *typed = reinterpret_cast<T *>(ctx->scalars + 4);
        if (cudaMemcpyAsync(typed, &nrm, sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            cudaGetLastError();
            return fail(SLAS_ERR_CUDA, "normalize: cannot upload the norm");
        }
        SLAS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // nrm lives on this thread's stack
        SLAS_TRY(launch_scale_div<T>(ctx, ctx->stream, cnt, (const T *)dv[0], (T *)dv[0], typed));
        return stage_all_finish(ctx, ops, sizes, 1, dv);
    });
}

template <typename T>
static int entry_gemm_batched(size_t batch, const T *a, const T *b, T *c, size_t m, size_t n, size_t k, int ta, int tb) {
    DevScope scope(c);
    if (multi_split_wanted(c, batch * (m * k + k * n + m * n) * sizeof(T))) {
        size_t granule = 1;
        while ((granule * m * k * sizeof(T)) % 16 || (granule * k * n * sizeof(T)) % 16 || (granule * m * n * sizeof(T)) % 16) granule *= 2;
        return split_units(batch, granule, [&](size_t lo, size_t cnt) {
            return api_gemm_batched<T>(cnt, a + lo * m * k, b + lo * k * n, c + lo * m * n, m, n, k, ta, tb);
        });
    }
    return api_gemm_batched<T>(batch, a, b, c, m, n, k, ta, tb);
}

template <typename T>
static int entry_gemv_batched(size_t batch, const T *a, const T *x, T *y, size_t m, size_t n, int ta) {
    DevScope scope(y);
    const size_t xlen = ta ? n : m, ylen = ta ? m : n;
    if (multi_split_wanted(y, batch * (m * n + xlen + ylen) * sizeof(T))) {
        size_t granule = 1;
        while ((granule * m * n * sizeof(T)) % 16 || (granule * xlen * sizeof(T)) % 16 || (granule * ylen * sizeof(T)) % 16) granule *= 2;
        return split_units(batch, granule, [&](size_t lo, size_t cnt) {
            return api_gemv<T>(cnt, a + lo * m * n, x + lo * xlen, y + lo * ylen, m, n, m, ta, true);
        });
    }
    return api_gemv<T>(batch, a, x, y, m, n, m, ta, true);
}

static int entry_transpose_batched(size_t batch, size_t len, size_t elem, const void *a, void *buffer, size_t columns, bool inplace) {
    DevScope scope(a);
    if (batch > 1 && multi_split_wanted(a, 2 * batch * len * elem)) {
        size_t granule = 1;
        while ((granule * len * elem) % 16) granule *= 2;
        return split_units(batch, granule, [&](size_t lo, size_t cnt) {
            return api_transpose(cnt, len, elem, (const char *)a + lo * len * elem, inplace ? nullptr : (char *)buffer + lo * len * elem, columns, inplace);
        });
    }
    return api_transpose(batch, len, elem, a, buffer, columns, inplace);
}

template <typename T>
static int entry_batched_reduce(int mode, size_t batch, size_t len, const T *a, const T *b, T *out, T *a_mut) {
    const void *anchor = mode == 2 ? (const void *)a_mut : (const void *)out;
    DevScope scope(anchor);
    if (multi_split_wanted(anchor, batch * len * sizeof(T) * (mode == 0 ? 2 : 1))) {
        size_t granule = 4;
        while ((granule * len * sizeof(T)) % 16) granule *= 2;
        return split_units(batch, granule, [&](size_t lo, size_t cnt) {
            return api_batched_reduce<T>(mode, cnt, len, a ? a + lo * len : nullptr, b ? b + lo * len : nullptr, out ? out + lo : nullptr,
                                         a_mut ? a_mut + lo * len : nullptr);
        });
    }
    return api_batched_reduce<T>(mode, batch, len, a, b, out, a_mut);
}

}  // namespace slas

using namespace slas;

extern "C" {

int slas_cuda_sfill_uniform(size_t len, float *x, uint64_t seed, uint64_t offset, float lo, float hi) {
    DevScope scope(x);
    return api_fill<float>(len, x, seed, offset, lo, hi);
}
int slas_cuda_dfill_uniform(size_t len, double *x, uint64_t seed, uint64_t offset, double lo, double hi) {
    DevScope scope(x);
    return api_fill<double>(len, x, seed, offset, lo, hi);
}

// DotProduct
int slas_cuda_sdot(size_t len, const float *a, const float *b, float *out) {
    return entry_reduce<float>(0, 0, len, a, b, out);
}
int slas_cuda_ddot(size_t len, const double *a, const double *b, double *out) {
    return entry_reduce<double>(0, 0, len, a, b, out);
}
int slas_cuda_cdotu(size_t len, const float *a, const float *b, float *out) {
    return entry_reduce<float>(2, 0, 2 * len, a, b, out);
}
int slas_cuda_zdotu(size_t len, const double *a, const double *b, double *out) {
    return entry_reduce<double>(2, 0, 2 * len, a, b, out);
}

// Normalize
int slas_cuda_snrm2(size_t len, const float *a, float *out) { return entry_reduce<float>(1, 1, len, a, a, out); }
int slas_cuda_dnrm2(size_t len, const double *a, double *out) { return entry_reduce<double>(1, 1, len, a, a, out); }
int slas_cuda_scnrm2(size_t len, const float *a, float *out) { return entry_reduce<float>(1, 1, 2 * len, a, a, out); }
int slas_cuda_dznrm2(size_t len, const double *a, double *out) { return entry_reduce<double>(1, 1, 2 * len, a, a, out); }
int slas_cuda_snormalize(size_t len, float *a) { return entry_normalize<float>(len, a); }
int slas_cuda_dnormalize(size_t len, double *a) { return entry_normalize<double>(len, a); }
int slas_cuda_cnormalize(size_t len, float *a) { return entry_normalize<float>(2 * len, a); }
int slas_cuda_znormalize(size_t len, double *a) { return entry_normalize<double>(2 * len, a); }

// Addition / Subtraction / Multiplication / Divition
int slas_cuda_sadd(size_t len, const float *a, const float *b, float *c) { return entry_ewise<float>(0, len, a, b, c); }
int slas_cuda_ssub(size_t len, const float *a, const float *b, float *c) { return entry_ewise<float>(1, len, a, b, c); }
int slas_cuda_smul(size_t len, const float *a, const float *b, float *c) { return entry_ewise<float>(2, len, a, b, c); }
int slas_cuda_sdiv(size_t len, const float *a, const float *b, float *c) { return entry_ewise<float>(3, len, a, b, c); }
int slas_cuda_dadd(size_t len, const double *a, const double *b, double *c) { return entry_ewise<double>(0, len, a, b, c); }
int slas_cuda_dsub(size_t len, const double *a, const double *b, double *c) { return entry_ewise<double>(1, len, a, b, c); }
int slas_cuda_dmul(size_t len, const double *a, const double *b, double *c) { return entry_ewise<double>(2, len, a, b, c); }
int slas_cuda_ddiv(size_t len, const double *a, const double *b, double *c) { return entry_ewise<double>(3, len, a, b, c); }

#define SLAS_FUSED_DOT(NAME, CHAIN)                                                                                         \
    int slas_cuda_s##NAME##_dot(size_t len, const float *a, const float *b, float *c, const float *d, float *out) {         \
        DevScope scope(a);                                                                                                  \
        return api_fused_chain<float>(CHAIN, len, a, b, c, d, out);                                                         \
    }                                                                                                                       \
    int slas_cuda_d##NAME##_dot(size_t len, const double *a, const double *b, double *c, const double *d, double *out) {    \
        DevScope scope(a);                                                                                                  \
        return api_fused_chain<double>(CHAIN, len, a, b, c, d, out);                                                        \
    }
SLAS_FUSED_DOT(add, 0)
SLAS_FUSED_DOT(sub, 1)
SLAS_FUSED_DOT(mul, 2)
SLAS_FUSED_DOT(div, 3)
#undef SLAS_FUSED_DOT
int slas_cuda_sdot_norms(size_t len, const float *a, const float *b, float *out3) {
    DevScope scope(a);
    return api_fused_chain<float>(4, len, a, b, nullptr, nullptr, out3);
}
int slas_cuda_ddot_norms(size_t len, const double *a, const double *b, double *out3) {
    DevScope scope(a);
    return api_fused_chain<double>(4, len, a, b, nullptr, nullptr, out3);
}
int slas_cuda_snormalize_into(size_t len, const float *a, float *out) {
    DevScope scope(a);
    return api_normalize_into<float>(len, a, out);
}
int slas_cuda_dnormalize_into(size_t len, const double *a, double *out) {
    DevScope scope(a);
    return api_normalize_into<double>(len, a, out);
}

// Transpose
int slas_cuda_transpose(size_t len, size_t elem_size, const void *a, void *buffer, size_t columns) {
    DevScope scope(a);
    return api_transpose(1, len, elem_size, a, buffer, columns, false);
}
int slas_cuda_transpose_inplace(size_t len, size_t elem_size, void *a, size_t columns) {
    DevScope scope(a);
    return api_transpose(1, len, elem_size, a, nullptr, columns, true);
}
int slas_cuda_transpose_inplace_batched(size_t batch, size_t len, size_t elem_size, void *a, size_t columns) {
    return entry_transpose_batched(batch, len, elem_size, a, nullptr, columns, true);
}
int slas_cuda_transpose_batched(size_t batch, size_t len, size_t elem_size, const void *a, void *buffer,
                                size_t columns) {
    return entry_transpose_batched(batch, len, elem_size, a, buffer, columns, false);
}

// MatrixMul
int slas_cuda_sgemm(const float *a, const float *b, float *c, size_t m, size_t n, size_t k, size_t lda, size_t ldb,
                    size_t ldc, int a_trans, int b_trans) {
    DevScope scope(c);
    return api_gemm<float>(a, b, c, m, n, k, lda, ldb, ldc, a_trans, b_trans);
}
int slas_cuda_dgemm(const double *a, const double *b, double *c, size_t m, size_t n, size_t k, size_t lda, size_t ldb,
                    size_t ldc, int a_trans, int b_trans) {
    DevScope scope(c);
    return api_gemm<double>(a, b, c, m, n, k, lda, ldb, ldc, a_trans, b_trans);
}
int slas_cuda_set_sgemm_path(int path) {
    SLAS_REQUIRE(path >= 0 && path <= 2, "sgemm path must be 0 (FFMA), 1 (tcgen05 3xTF32) or 2 (auto)");
    g_sgemm_path = path;
    return SLAS_OK;
}
int slas_cuda_get_sgemm_path(int *path) {
    SLAS_REQUIRE(path, "null path");
    *path = g_sgemm_path;
    return SLAS_OK;
}
int slas_cuda_sgemm_batched(size_t batch, const float *a, const float *b, float *c, size_t m, size_t n, size_t k,
                            int a_trans, int b_trans) {
    return entry_gemm_batched<float>(batch, a, b, c, m, n, k, a_trans, b_trans);
}
int slas_cuda_dgemm_batched(size_t batch, const double *a, const double *b, double *c, size_t m, size_t n, size_t k,
                            int a_trans, int b_trans) {
    return entry_gemm_batched<double>(batch, a, b, c, m, n, k, a_trans, b_trans);
}

// matrix_vector_mul
int slas_cuda_sgemv(const float *a, const float *x, float *y, size_t m, size_t n, size_t lda, int a_trans) {
    DevScope scope(y);
    return api_gemv<float>(1, a, x, y, m, n, lda, a_trans, false);
}
int slas_cuda_dgemv(const double *a, const double *x, double *y, size_t m, size_t n, size_t lda, int a_trans) {
    DevScope scope(y);
    return api_gemv<double>(1, a, x, y, m, n, lda, a_trans, false);
}
int slas_cuda_sgemv_batched(size_t batch, const float *a, const float *x, float *y, size_t m, size_t n, int a_trans) {
    return entry_gemv_batched<float>(batch, a, x, y, m, n, a_trans);
}
int slas_cuda_dgemv_batched(size_t batch, const double *a, const double *x, double *y, size_t m, size_t n, int a_trans) {
    return entry_gemv_batched<double>(batch, a, x, y, m, n, a_trans);
}

// batched dot / norm / normalize
int slas_cuda_sdot_batched(size_t batch, size_t len, const float *a, const float *b, float *out) {
    return entry_batched_reduce<float>(0, batch, len, a, b, out, nullptr);
}
int slas_cuda_ddot_batched(size_t batch, size_t len, const double *a, const double *b, double *out) {
    return entry_batched_reduce<double>(0, batch, len, a, b, out, nullptr);
}
int slas_cuda_snrm2_batched(size_t batch, size_t len, const float *a, float *out) {
    return entry_batched_reduce<float>(1, batch, len, a, nullptr, out, nullptr);
}
int slas_cuda_dnrm2_batched(size_t batch, size_t len, const double *a, double *out) {
    return entry_batched_reduce<double>(1, batch, len, a, nullptr, out, nullptr);
}
int slas_cuda_snormalize_batched(size_t batch, size_t len, float *a) {
    return entry_batched_reduce<float>(2, batch, len, nullptr, nullptr, nullptr, a);
}
int slas_cuda_dnormalize_batched(size_t batch, size_t len, double *a) {
    return entry_batched_reduce<double>(2, batch, len, nullptr, nullptr, nullptr, a);
}

// sharded (one slice per rank + one allreduce)
int slas_cuda_sdot_sharded(size_t local_len, const float *a, const float *b, float *out) {
    DevScope scope(a);
    return api_reduce<float>(0, 0, local_len, a, b, out, true, nullptr);
}
int slas_cuda_ddot_sharded(size_t local_len, const double *a, const double *b, double *out) {
    DevScope scope(a);
    return api_reduce<double>(0, 0, local_len, a, b, out, true, nullptr);
}
int slas_cuda_snrm2_sharded(size_t local_len, const float *a, float *out) {
    DevScope scope(a);
    return api_reduce<float>(1, 1, local_len, a, a, out, true, nullptr);
}
int slas_cuda_dnrm2_sharded(size_t local_len, const double *a, double *out) {
    DevScope scope(a);
    return api_reduce<double>(1, 1, local_len, a, a, out, true, nullptr);
}
int slas_cuda_snormalize_sharded(size_t local_len, float *a) { return api_normalize<float>(local_len, a, true); }
int slas_cuda_dnormalize_sharded(size_t local_len, double *a) { return api_normalize<double>(local_len, a, true); }
int slas_cuda_sdot_partial(size_t len, const float *a, const float *b, double *partial) {
    DevScope scope(a);
    return api_reduce<float>(0, 0, len, a, b, nullptr, false, partial);
}
int slas_cuda_ddot_partial(size_t len, const double *a, const double *b, double *partial) {
    DevScope scope(a);
    return api_reduce<double>(0, 0, len, a, b, nullptr, false, partial);
}

}  // extern "C"
