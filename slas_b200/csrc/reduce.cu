// reduce.cu -- grid-wide reductions of the vector half of the hot path:
//   dot / complex dotu   (replaces src/backends/rust.rs:20-41, src/backends/blas.rs:15-52)
//   norm                 (replaces src/backends/rust.rs:116-121, 129-140, src/backends/blas.rs:156-170)
// 128-bit coalesced streaming loads, four independent 16-byte loads per operand in flight per thread, fp
// accumulators flushed into f64, warp shuffles + block reduction, and a deterministic single-launch grid combine
// (reduce.cuh).  No float atomics anywhere: same input + same launch geometry => same bits.
#include <algorithm>

#include "common.cuh"
#include "divby.cuh"
#include "kernels.cuh"
#include "reduce.cuh"

namespace slas {

// ---- geometry (shared with fused_chain.cu) -------------------------------------------------------------------------
RedGeom reduce_geometry(const Context *ctx, size_t work, bool vectorised) {
    RedGeom g;
    const size_t sms = (size_t)ctx->sm_count;
    const size_t tile = (size_t)256 * kRedUnroll;
    // balanced single-wave geometry while the tiled one would not yet fill its cap of 8 CTAs per SM
    const long mode = tunable("reduce_small");          // 0 default, 1 = tiled everywhere (the round-1 geometry)
    const bool small = vectorised && mode != 1 && work < sms * 8 * tile;
    if (!small) {
        size_t blocks = (work + tile - 1) / tile;
        size_t cap = sms * 8;
        if (cap > 1536) cap = 1536;                      // 3 results x cap partials fit the scratch (2 * kMaxPartials)
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        g.blocks = (unsigned)blocks;
        g.threads = 256;
        return g;
    }
    // Measured on B200 (profiles/r2d_tune_small.log, 256 distinct 2^20-element pairs replayed from a graph): back-to-back
    // calls overlap (dependent launch), and the smaller a kernel's footprint per SM the more of its successor fits next
    // to it -- 256 threads x 2 CTAs per SM: 2.85 us per call, 512 x 1: 3.05, 512 x 2: 3.33, 1024 x 1: 3.58 (an isolated
    // kernel is fastest the other way round: 1024 x 1 writes its result 3.3 us after it starts, 256 x 2 after 5.3 us).
    long threads = tunable("reduce_small_threads");      // 256 / 512 / 1024; 0 = default
    if (threads != 256 && threads != 512 && threads != 1024) threads = 256;
    long cps = tunable("reduce_small_cps");              // CTAs per SM; 0 = default
    if (cps <= 0) cps = threads == 256 ? 2 : 1024 / threads;
    if (cps * threads > 2048) cps = 2048 / threads;
    g.threads = (unsigned)threads;
    g.balanced = 1;
    g.poll = tunable("reduce_small_tail") == 1 ? 0 : 1;  // 1 = ticket tail (for the comparison), default flagged words
    // at least one full group of loads per thread before another block is worth its cross-block step
    size_t blocks = (work + (size_t)threads * 2 - 1) / ((size_t)threads * 2);
    const size_t wave = sms * (size_t)cps;
    if (blocks >= wave) blocks = wave;
    if (blocks == 0) blocks = 1;
    size_t per = (work + blocks - 1) / blocks;
    per = (per + 7) / 8 * 8;                             // block ranges start on 128-byte lines
    if (per == 0) per = 8;                               // an empty vector still launches one block (it writes the zero)
    blocks = (work + per - 1) / per;
    if (blocks == 0) blocks = 1;
    g.blocks = (unsigned)blocks;
    g.per = per;
    return g;
}

RedScratch reduce_scratch(Context *ctx) {
    RedScratch sc;
    sc.partials = ctx->partials;
    sc.ticket = ctx->ticket;
    // consecutive launches use consecutive sets of flagged slots: kernels that overlap (dependent launch, at most
    // Hazard::kWindow == kRedScratchSets of them between two barriers) never publish into the same set
    sc.ll = reinterpret_cast<LLSlot *>(ctx->ll_slots) + (size_t)(ctx->red_seq++ % kRedScratchSets) * 3 * kMaxPartials;
    sc.phases = tunable("reduce_phases") ? ctx->phases : nullptr;
    return sc;
}

// result[0] (and result[1] for KIND 2) receive the f64 sums; typed_out (may be null) receives
// mode 0: (T)sum [, (T)sum_im]   mode 1: (T)sqrt(sum).  If accumulate != 0 the previous
// content of result[] is added first (chunked host vectors).
template <typename T, int KIND, int THREADS>
__global__ void __launch_bounds__(THREADS, RedTile<THREADS>::MINB)
reduce_kernel(const T *__restrict__ a, const T *__restrict__ b, size_t len, size_t nvec, int vectorised, const RedGeom g,
              const RedScratch sc, double *__restrict__ result, T *__restrict__ typed_out, int mode, int accumulate,
              const PeerExchange px, unsigned head, int early) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    constexpr int NRES = KIND == 2 ? 2 : 1;
    constexpr int U = RedTile<THREADS>::U;
    __shared__ double red[THREADS / 32];
    __shared__ double peer_vals[2 * kMaxPeers];
    __shared__ bool is_last;
    // programmatic dependent launch (reduce.cuh): the successor may be scheduled from here on; this kernel's own
    // loads wait for its predecessor unless the host found no overlap with that kernel's operands (`early`)
    if (!early) pdl_wait();  // a barrier kernel: its successor starts only after everything before this kernel is done
    pdl_trigger();
    const bool stamp = sc.phases != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (stamp) sc.phases[0] = red_timer_ns();
    Reducer<T, KIND> r;
    r.clear();
    const size_t tid = (size_t)blockIdx.x * THREADS + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * THREADS;
    if (vectorised) {
        // operands that share a misalignment (views at an odd offset): `head` elements up to the first 16-byte boundary
        // are scalar work like the tail, the body starts on the boundary
        const V *av = reinterpret_cast<const V *>(a + head);
        const V *bv = reinterpret_cast<const V *>(b + head);
        size_t i, end, stride;
        red_range<THREADS>(g, nvec, i, end, stride);
        int since_flush = 0;
        // full groups: all U (RedTile) vectors of this thread are in range
        for (; i + (size_t)(U - 1) * THREADS < end; i += stride) {
            V x[U], y[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = ld_stream(av + i + (size_t)u * THREADS);
            if (KIND != 1) {
#pragma unroll
                for (int u = 0; u < U; ++u) y[u] = ld_stream(bv + i + (size_t)u * THREADS);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) r.add(u, x[u], KIND == 1 ? x[u] : y[u]);
            if (++since_flush == kFlushEvery) { r.flush(); since_flush = 0; }
        }
        // ragged last group: every load is issued before the first use (this IS the whole body of a short vector)
        {
            V x[U], y[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t j = i + (size_t)u * THREADS;
                ok[u] = j < end;
                x[u] = ok[u] ? ld_stream(av + j) : vzero((V *)nullptr);
            }
            if (KIND != 1) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const size_t j = i + (size_t)u * THREADS;
                    y[u] = ok[u] ? ld_stream(bv + j) : vzero((V *)nullptr);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) r.add(u, x[u], KIND == 1 ? x[u] : y[u]);
        }
        r.flush();
        // scalar tail: the len % N elements past the last whole vector (N is even, so a complex
        // pair is never split: the complex tail is whole (re, im) pairs)
        if (KIND != 2) {
            const size_t j = head + nvec * N + tid;
            if (j < len) r.add_scalar(a[j], KIND == 1 ? a[j] : b[j]);
            if (tid < head) r.add_scalar(a[tid], KIND == 1 ? a[tid] : b[tid]);
        } else {
            const size_t j = head + nvec * N + 2 * tid;
            if (j + 1 < len) {
                const double ar = a[j], ai = a[j + 1], br = b[j], bi = b[j + 1];
                r.sum += ar * br - ai * bi;
                r.sum_im += ar * bi + ai * br;
            }
            if (2 * tid + 1 < head) {
                const double ar = a[2 * tid], ai = a[2 * tid + 1], br = b[2 * tid], bi = b[2 * tid + 1];
                r.sum += ar * br - ai * bi;
                r.sum_im += ar * bi + ai * br;
            }
        }
    } else {
        // unaligned operands: scalar grid-stride loop, f64 accumulation
        if (KIND == 2) {
            for (size_t j = tid; j < len / 2; j += nthreads) {
                const double ar = a[2 * j], ai = a[2 * j + 1], br = b[2 * j], bi = b[2 * j + 1];
                r.sum += ar * br - ai * bi;
                r.sum_im += ar * bi + ai * br;
            }
        } else {
            // operands with different misalignments: coalesced 4-byte / 8-byte loads, four in flight per thread, T
            // accumulators flushed into f64 like the vector body (one DFMA per element was 0.35 of HBM peak)
            size_t j = tid;
            int since_flush = 0;
            for (; j + 3 * nthreads < len; j += 4 * nthreads) {
                T x[4], y[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = a[j + u * nthreads];
#pragma unroll
                for (int u = 0; u < 4; ++u) y[u] = KIND == 1 ? x[u] : b[j + u * nthreads];
#pragma unroll
                for (int u = 0; u < 4; ++u) r.acc[u] = fma(x[u], y[u], r.acc[u]);
                if (++since_flush == kFlushEvery * 4) { r.flush(); since_flush = 0; }
            }
            r.flush();
            for (; j < len; j += nthreads) r.add_scalar(a[j], KIND == 1 ? a[j] : b[j]);
        }
    }
    if (stamp) sc.phases[1] = red_timer_ns();
    double s[NRES], p[NRES];
    s[0] = block_sum<THREADS>(r.sum, red);
    if (KIND == 2) s[NRES - 1] = block_sum<THREADS>(r.sum_im, red);
    if (stamp) sc.phases[2] = red_timer_ns();
    // early kernels: the flagged-slot combine runs in this launch's own scratch set and the result goes to memory no
    // running kernel touches (host-checked), so it need not wait either; only the ticket tail shares its counter.
    // Every block still waits for the predecessor before it exits: completion order stays the stream order.
    const bool wait_before_combine = early && !g.poll;
    if (wait_before_combine) pdl_wait();
    if (!grid_combine<THREADS, NRES>(g, sc, s, p, red, &is_last)) {
        if (early && !wait_before_combine) pdl_wait();
        return;
    }
    const bool stamp_end = sc.phases != nullptr && threadIdx.x == 0;
    if (stamp_end) sc.phases[3] = red_timer_ns();
    double pr = p[0], pi = KIND == 2 ? p[NRES - 1] : 0.0;
    // multi-GPU: exchange the rank-local partial with the peers over NVLink and add in rank order
    if (px.nranks > 1) peer_allreduce_sum(px, pr, pi, peer_vals);
    if (threadIdx.x == 0) {
        if (result) {
            if (accumulate) { pr += result[0]; if (KIND == 2) pi += result[1]; }
            result[0] = pr;
            if (KIND == 2) result[1] = pi;
        }
        if (typed_out) {
            if (mode == 1) typed_out[0] = (T)sqrt(pr);
            else {
                typed_out[0] = (T)pr;
                if (KIND == 2) typed_out[1] = (T)pi;
            }
        }
        if (stamp_end) sc.phases[4] = red_timer_ns();
    }
    if (early && !wait_before_combine) pdl_wait();
}

template <typename T, int KIND>
static int launch_reduce_kind(Context *ctx, cudaStream_t st, size_t len, const T *a, const T *b,
                              double *result, T *typed_out, int mode, int accumulate, const PeerExchange &px) {
    constexpr int N = Vec16<T>::N;
    // vector body when the operands are aligned, or share one misalignment that is whole elements (for complex: whole
    // pairs) away from a 16-byte boundary
    const uintptr_t ma = (uintptr_t)a % 16, mb = (uintptr_t)(KIND == 1 ? a : b) % 16;
    const size_t unit = sizeof(T) * (KIND == 2 ? 2 : 1);
    const bool vec = ma == mb && ma % unit == 0;
    size_t head = vec && ma ? (16 - ma) / sizeof(T) : 0;
    if (head > len) head = len;
    const size_t nvec = vec ? (len - head) / N : 0;
    const RedGeom g = reduce_geometry(ctx, vec ? nvec : len, vec);
    const RedScratch sc = reduce_scratch(ctx);
    const T *bb = KIND == 1 ? a : b;
    // what this kernel reads and writes, for the early-start decision of it and of its successor
    const Context::Span reads[2] = {{(const char *)a, (const char *)(a + len)}, {(const char *)bb, (const char *)(bb + len)}};
    constexpr int NRES = KIND == 2 ? 2 : 1;
    const Context::Span writes[2] = {{(const char *)result, (const char *)(result ? result + NRES : result)},
                                     {(const char *)typed_out, (const char *)(typed_out ? typed_out + NRES : typed_out)}};
    // overlap with the predecessor only on a stream nobody else enqueues kernels on (the library's own, or one the
    // caller vouches for: tunable "pdl" = 2), without the accumulate / multi-GPU forms, and when nothing overlaps
    const bool own = st == ctx->own_stream || tunable("pdl") == 2;
    // (an early kernel runs to its end without waiting: loads, combine in its own scratch set, result store)
    const int early = g.balanced && own && !accumulate && px.nranks <= 1 && hazard_early_ok(ctx, st, reads, 2, writes, 2) ? 1 : 0;
    cudaError_t le = cudaSuccess;
#define SLAS_REDUCE_LAUNCH(THREADS)                                                                                         \
    do {                                                                                                                    \
        if (g.balanced && tunable("pdl") != 1)                                                                              \
            le = launch_pdl(reduce_kernel<T, KIND, THREADS>, g.blocks, THREADS, st, a, bb, len, nvec, vec ? 1 : 0, g, sc,    \
                            result, typed_out, mode, accumulate, px, (unsigned)head, early);                                \
        else                                                                                                                \
            reduce_kernel<T, KIND, THREADS><<<g.blocks, THREADS, 0, st>>>(a, bb, len, nvec, vec ? 1 : 0, g, sc, result,      \
                                                                          typed_out, mode, accumulate, px, (unsigned)head, 0); \
    } while (0)
    if (g.threads == 1024) SLAS_REDUCE_LAUNCH(1024);
    else if (g.threads == 512) SLAS_REDUCE_LAUNCH(512);
    else SLAS_REDUCE_LAUNCH(256);
#undef SLAS_REDUCE_LAUNCH
    SLAS_CUDA_TRY(le);
    SLAS_LAUNCH_CHECK();
    hazard_record(ctx, st, reads, 2, writes, 2, early);
    return SLAS_OK;
}

template <typename T>
int launch_reduce(Context *ctx, cudaStream_t st, int kind, size_t len, const T *a, const T *b, double *result,
                  T *typed_out, int mode, int accumulate, const PeerExchange *peers) {
    const PeerExchange px = peers ? *peers : PeerExchange{};
    switch (kind) {
    case 0: return launch_reduce_kind<T, 0>(ctx, st, len, a, b, result, typed_out, mode, accumulate, px);
    case 1: return launch_reduce_kind<T, 1>(ctx, st, len, a, a, result, typed_out, mode, accumulate, px);
    case 2: return launch_reduce_kind<T, 2>(ctx, st, len, a, b, result, typed_out, mode, accumulate, px);
    }
    return fail(SLAS_ERR_INVALID, "unknown reduction kind %d", kind);
}
template int launch_reduce<float>(Context *, cudaStream_t, int, size_t, const float *, const float *, double *,
                                  float *, int, int, const PeerExchange *);
template int launch_reduce<double>(Context *, cudaStream_t, int, size_t, const double *, const double *,
                                   double *, double *, int, int, const PeerExchange *);

// ---- normalize of a short vector in ONE launch ------------------------------------------------------------------------
// (rust.rs:122-127: a[i] /= norm(a).)  For vectors that fit the registers of one CTA wave (<= 4 vectors of 16 bytes per
// thread: LEN <= ~1.2 M f32) the two passes collapse: every thread keeps its part of the vector in registers, the blocks
// combine their partial sums of squares exactly like reduce_kernel<T, 1> does (same geometry, accumulators and order:
// the norm carries the bits slas_cuda_?nrm2 returns), block 0 publishes the norm as flagged words, every block picks it
// up and divides its registers (IEEE div.rn) -- 8 bytes per element instead of 12, one launch instead of two.  All
// blocks are resident at once (one wave by construction), so waiting for block 0 cannot deadlock.  Dependent launch as
// for reduce_kernel: the partial slots, the broadcast slot and the acknowledge counter belong to this launch's scratch
// set, so an "early" kernel (no overlap with what may still be running: it reads AND writes its vector) runs to its end
// without waiting for its predecessor.
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS)
normalize_fused_kernel(T *__restrict__ a, size_t len, size_t nvec, const RedGeom g, const RedScratch sc, LLSlot *bcast,
                       unsigned int *ack, T *__restrict__ norm_out, int early) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    constexpr int U = RedTile<THREADS>::U;
    __shared__ double red[THREADS / 32];
    __shared__ bool is_last;
    __shared__ T nrm_s;
    if (!early) pdl_wait();
    pdl_trigger();
    Reducer<T, 1> r;
    r.clear();
    V *av = reinterpret_cast<V *>(a);
    size_t i, end, stride;
    red_range<THREADS>(g, nvec, i, end, stride);
    V x[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t j = i + (size_t)u * THREADS;
        ok[u] = j < end;
        x[u] = ok[u] ? ld_stream(av + j) : vzero((V *)nullptr);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (ok[u]) r.add(u, x[u], x[u]);
    r.flush();
    const size_t tid = (size_t)blockIdx.x * THREADS + threadIdx.x;
    const size_t jt = nvec * N + tid;  // the len % N elements past the last whole vector
    T tail = T(0);
    if (jt < len) { tail = a[jt]; r.add_scalar(tail, tail); }
    double s[1], p[1];
    s[0] = block_sum<THREADS>(r.sum, red);
    const bool combiner = grid_combine<THREADS, 1>(g, sc, s, p, red, &is_last);
    if (combiner && threadIdx.x == 0) {
        const T nrm = (T)sqrt(p[0]);
        if (norm_out) norm_out[0] = nrm;
        nrm_s = nrm;
        if (gridDim.x > 1) ll_store(bcast, (double)nrm);  // exact: T -> double -> T
    }
    if (!combiner && threadIdx.x == 0) {
        // every block but the combiner waits for the norm (without consuming the slot: the last reader resets it)
        unsigned long long w0, w1;
        const unsigned long long t0 = red_timer_ns();
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0) : "l"(&bcast->w[0]) : "memory");
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w1) : "l"(&bcast->w[1]) : "memory");
            if ((unsigned)(w0 >> 32) == kLLFlag && (unsigned)(w1 >> 32) == kLLFlag) break;
            if ((++spins & 0xfff) == 0 && red_timer_ns() - t0 > 20000000000ull) __trap();
        }
        nrm_s = (T)__longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    }
    __syncthreads();
    DivBy<T> by;
    by.init(nrm_s);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (ok[u]) {
            T *e = reinterpret_cast<T *>(&x[u]);
#pragma unroll
            for (int q = 0; q < N; ++q) e[q] = by.div(e[q]);  // the bits of IEEE div.rn (divby.cuh): not a reciprocal multiply
            st_stream(av + i + (size_t)u * THREADS, x[u]);
        }
    }
    if (jt < len) a[jt] = by.div(tail);
    // the last block to have read the norm clears the slot for the next launch
    if (gridDim.x > 1 && threadIdx.x == 0) {
        const unsigned int t = atomicInc(ack, gridDim.x - 1);
        if (t == gridDim.x - 1) {
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&bcast->w[0]), "l"(0ull) : "memory");
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&bcast->w[1]), "l"(0ull) : "memory");
        }
    }
    if (early) pdl_wait();  // completion order stays the stream order
}

// *handled = false: the vector does not fit one wave's registers (or is a misaligned view): use the two-pass path
template <typename T>
int launch_normalize_fused(Context *ctx, cudaStream_t st, size_t len, T *a, T *norm_out, bool *handled) {
    constexpr int N = Vec16<T>::N;
    *handled = false;
    if (len == 0 || !aligned16(a) || tunable("normalize_fused") == 1) return SLAS_OK;
    const size_t nvec = len / N;
    const RedGeom g = reduce_geometry(ctx, nvec, true);
    if (!g.balanced || !g.poll) return SLAS_OK;
    // every thread must hold its whole share in one group of loads
    if (g.per > (unsigned long long)g.threads * (g.threads == 1024 ? 2 : kRedUnroll)) return SLAS_OK;
    // this launch's scratch set: partial slots (reduce_scratch rotates), one broadcast slot and one acknowledge counter
    const unsigned set = ctx->red_seq % kRedScratchSets;
    const RedScratch sc = reduce_scratch(ctx);
    LLSlot *tail = reinterpret_cast<LLSlot *>(ctx->ll_slots) + (size_t)kRedScratchSets * 3 * kMaxPartials;  // past the sets
    LLSlot *bcast = tail + set;
    unsigned int *ack = reinterpret_cast<unsigned int *>(tail + kRedScratchSets) + set;
    const Context::Span rw[1] = {{(const char *)a, (const char *)(a + len)}};
    const Context::Span w2[2] = {{(const char *)a, (const char *)(a + len)}, {(const char *)norm_out, (const char *)(norm_out ? norm_out + 1 : norm_out)}};
    const bool own = st == ctx->own_stream || tunable("pdl") == 2;
    const int early = own && hazard_early_ok(ctx, st, rw, 1, w2, 2) ? 1 : 0;
    cudaError_t le;
    if (g.threads == 1024) le = launch_pdl(normalize_fused_kernel<T, 1024>, g.blocks, 1024u, st, a, len, nvec, g, sc, bcast, ack, norm_out, early);
    else if (g.threads == 512) le = launch_pdl(normalize_fused_kernel<T, 512>, g.blocks, 512u, st, a, len, nvec, g, sc, bcast, ack, norm_out, early);
    else le = launch_pdl(normalize_fused_kernel<T, 256>, g.blocks, 256u, st, a, len, nvec, g, sc, bcast, ack, norm_out, early);
    SLAS_CUDA_TRY(le);
    SLAS_LAUNCH_CHECK();
    hazard_record(ctx, st, rw, 1, w2, 2, early);
    *handled = true;
    return SLAS_OK;
}
template int launch_normalize_fused<float>(Context *, cudaStream_t, size_t, float *, float *, bool *);
template int launch_normalize_fused<double>(Context *, cudaStream_t, size_t, double *, double *, bool *);

// f64 sum -> typed scalar (used after the NCCL allreduce of the sharded ops)
template <typename T>
__global__ void finish_scalar_kernel(const double *in, T *out, int mode) {
    out[0] = mode == 1 ? (T)sqrt(in[0]) : (T)in[0];
}
template <typename T>
int launch_finish_scalar(cudaStream_t st, const double *in, T *out, int mode) {
    finish_scalar_kernel<T><<<1, 1, 0, st>>>(in, out, mode);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}
template int launch_finish_scalar<float>(cudaStream_t, const double *, float *, int);
template int launch_finish_scalar<double>(cudaStream_t, const double *, double *, int);

}  // namespace slas
