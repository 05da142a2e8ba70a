// staged_objects_kernel.cuh -- device side of the bulk-TMA staged kernels for large batches of SMALL objects whose byte
// length is not a whole number of 16-byte words (3-vectors, 3x3 matrices, LEN 5, 7, 30 ...): batched dot / norm /
// normalize (replaces a user's loop over rust.rs:20-41, 116-127) and batched transposes (rust.rs:151-177).
//
// HBM only ever sees 1-D bulk copies: a producer warp streams chunks of G consecutive objects into a three-stage
// shared-memory ring (cp.async.bulk + mbarrier), 256 consumer threads work on the chunk in shared memory, and results
// that are themselves objects (normalize, transpose) leave through a double-buffered shared tile with one bulk store per
// chunk.
//
// Object shapes are compile-time constants in slas (StaticVec<T, LEN>, MatrixShape<M, K>: rustc monomorphises every
// call).  The kernels below take the shape BOTH ways: template arguments CLEN / CROWS / CCOLS != 0 make it a constant
// (loops unroll, index arithmetic folds, a small object lives in its thread's registers), 0 means "use the run-time
// argument".  vector_ops.cu / transpose.cu instantiate the run-time form and the commonest constants with nvcc;
// staged_jit.cu compiles any other shape with NVRTC the first time it is seen -- this very text, embedded at build time.
// ncu, 2^24 objects, run-time shape (profiles/r2s_*): 3x3 f32 transpose 23 instructions per element, 3-vector normalize
// ~100 per vector -- both bound by instruction issue (IPC 2.5-2.7 of 4), not by HBM.
// Free of host headers for that reason.
#pragma once
#include "async.cuh"
#include "divby.cuh"

namespace slas {

// ---- batched dot (MODE 0) / norm (1) / normalize in place (2) ------------------------------------------------------------
// Each consumer thread owns whole vectors of the chunk: a plain ascending loop of FMAs in T from 0 -- the reference's
// naive loop (rust.rs:116-121) -- then sqrt.rn / div.rn.  The same operations in the same order for every CLEN, so the
// compile-time and the run-time form return the same bits.
template <typename T, int MODE, unsigned CLEN>
__global__ void __launch_bounds__(288)
batched_reduce_staged_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ out, T *a_mut, size_t batch,
                             size_t nchunks, unsigned len_arg, unsigned G, unsigned stage_bytes) {
    constexpr int CONSUMERS = 256, STAGES = 3;
    constexpr int NIN = MODE == 0 ? 2 : 1;  // operands per stage
    constexpr bool IN_REGS = CLEN != 0 && CLEN <= 16;
    const unsigned len = CLEN ? CLEN : len_arg;
    extern __shared__ __align__(128) unsigned char staged_smem[];
    unsigned char *out_base = staged_smem + (size_t)STAGES * NIN * stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(out_base + (MODE == 2 ? 2 * (size_t)stage_bytes : 0));
    uint64_t *empty = full + STAGES;
    DivBy<T> *norms = reinterpret_cast<DivBy<T> *>(empty + STAGES);  // MODE 2: the chunk's G norms (as divisors, divby.cuh)
    // measured: up to 8 elements the owning thread scales its own vector (LEN 3 f32 0.72 against 0.59); longer ones are
    // scaled by all threads over consecutive elements (LEN 38 f32 0.64 against 0.53, LEN 19 f64 0.80 against 0.67)
    const bool coop_scale = len > 8;
    const T *abase = MODE == 2 ? a_mut : a;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, CONSUMERS / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == CONSUMERS / 32) {
        if (lane == 0) {
            uint32_t it = 0;
            for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
                const uint32_t s = it % STAGES, use = it / STAGES;
                if (use > 0) mbar_wait_backoff(empty + s, (use - 1) & 1);
                const size_t first = ch * G;
                const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
                const uint32_t bytes = (cnt * len * (uint32_t)sizeof(T) + 15u) & ~15u;  // the arrays end on a 16-byte word
                mbar_expect_tx(full + s, NIN * bytes);
                bulk_g2s(staged_smem + (size_t)s * NIN * stage_bytes, abase + first * len, bytes, full + s);
                if (MODE == 0) bulk_g2s(staged_smem + (size_t)s * NIN * stage_bytes + stage_bytes, b + first * len, bytes, full + s);
            }
        }
        return;
    }
    uint32_t it = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        const size_t first = ch * G;
        const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
        const T *sa = reinterpret_cast<const T *>(staged_smem + (size_t)s * NIN * stage_bytes);
        const T *sb = MODE == 0 ? reinterpret_cast<const T *>(staged_smem + (size_t)s * NIN * stage_bytes + stage_bytes) : sa;
        T *cst = reinterpret_cast<T *>(out_base + (size_t)(it & 1) * stage_bytes);
        mbar_wait(full + s, use & 1);
        if (MODE == 2) asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");  // the bulk store that read this tile is done
        for (uint32_t v = threadIdx.x; v < cnt; v += CONSUMERS) {
            const T *pa = sa + v * len, *pb = sb + v * len;
            T acc = T(0);
            if constexpr (IN_REGS) {
                // the vector lives in registers between the norm and the division
                T x[CLEN ? CLEN : 1];
#pragma unroll
                for (unsigned i = 0; i < CLEN; ++i) x[i] = pa[i];
#pragma unroll
                for (unsigned i = 0; i < CLEN; ++i) acc = fma(x[i], MODE == 0 ? pb[i] : x[i], acc);
                if (MODE == 0) {
                    out[first + v] = acc;
                } else if (MODE == 1) {
                    out[first + v] = sqrt(acc);  // correctly rounded in T (sqrt.rn): the same bits as rounding a wider sqrt
                } else if (CLEN > 8) {
                    norms[v].init(sqrt(acc));
                } else {
                    DivBy<T> by;
                    by.init(sqrt(acc));
                    T *pc = cst + v * len;
                    bool slow = false;
#pragma unroll
                    for (unsigned i = 0; i < CLEN; ++i) {
                        pc[i] = by.fast(x[i]);
                        slow |= !by.ok(x[i]);
                    }
                    if (!DivBy<T>::kAlwaysOk && slow) {  // zero / subnormal / huge operands: the plain division
#pragma unroll
                        for (unsigned i = 0; i < CLEN; ++i)
                            if (!by.ok(x[i])) pc[i] = x[i] / by.n;
                    }
                }
            } else {
                if constexpr (CLEN != 0) {
                    // eight loads in flight ahead of the FMA chain (with a run-time length the same pragma costs more in
                    // remainder handling than it gains: LEN 38 dot 0.99 -> 0.73)
#pragma unroll 8
                    for (unsigned i = 0; i < CLEN; ++i) acc = fma(pa[i], pb[i], acc);
                } else {
                    for (unsigned i = 0; i < len; ++i) acc = fma(pa[i], pb[i], acc);
                }
                if (MODE == 0) {
                    out[first + v] = acc;
                } else if (MODE == 1) {
                    out[first + v] = sqrt(acc);
                } else if (coop_scale) {
                    norms[v].init(sqrt(acc));
                } else {
                    DivBy<T> by;
                    by.init(sqrt(acc));
                    T *pc = cst + v * len;
                    bool slow = false;
                    for (unsigned i = 0; i < len; ++i) {
                        pc[i] = by.fast(pa[i]);
                        slow |= !by.ok(pa[i]);
                    }
                    if (!DivBy<T>::kAlwaysOk && slow) {
                        for (unsigned i = 0; i < len; ++i)
                            if (!by.ok(pa[i])) pc[i] = pa[i] / by.n;
                    }
                }
            }
        }
        if (MODE == 2 && coop_scale) {
            // scale pass over the chunk's elements, all threads, consecutive elements (no bank conflicts, no idle lanes);
            // (vector, element) of the linear index kept up to date with additions
            asm volatile("bar.sync 3, %0;" ::"n"(CONSUMERS) : "memory");
            const unsigned total = cnt * len, step_v = CONSUMERS / len, step_i = CONSUMERS % len;
            unsigned v = threadIdx.x / len, i = threadIdx.x % len;
            const unsigned v0 = v, i0 = i;
            bool slow = false;
#pragma unroll 4
            for (unsigned e = threadIdx.x; e < total; e += CONSUMERS) {
                const DivBy<T> by = norms[v];
                const T x = sa[e];
                cst[e] = by.fast(x);
                slow |= !by.ok(x);
                v += step_v; i += step_i;
                if (i >= len) { i -= len; ++v; }
            }
            if (!DivBy<T>::kAlwaysOk && slow) {  // zero / subnormal / huge operands (rare): the plain division for those
                v = v0; i = i0;
                for (unsigned e = threadIdx.x; e < total; e += CONSUMERS) {
                    const DivBy<T> by = norms[v];
                    const T x = sa[e];
                    if (!by.ok(x)) cst[e] = x / by.n;
                    v += step_v; i += step_i;
                    if (i >= len) { i -= len; ++v; }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (MODE == 2) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 2, %0;" ::"n"(CONSUMERS) : "memory");
            if (threadIdx.x == 0) {
                const size_t total = (size_t)cnt * len, cbytes = total * sizeof(T);
                const uint32_t whole = (uint32_t)(cbytes & ~(size_t)15);
                if (whole)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a_mut + first * len),
                                 "r"(smem_u32(cst)), "r"(whole)
                                 : "memory");
                for (size_t e = whole / sizeof(T); e < total; ++e) a_mut[first * len + e] = cst[e];
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
        }
    }
    if (MODE == 2 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- batched transposes of small objects of any shape ------------------------------------------------------------------
// buffer[columns*row + column] = a[rows*column + row] per object (rust.rs:165-176); objects with len == rows * columns
// only (a bulk store would overwrite a never-written tail).  Works in place too: a chunk is read completely before it is
// stored.  Two ways through a chunk:
//   PER_OBJECT (compile-time shape, an object fits a thread's registers, object stride in 32-bit words shares at most a
//     factor 2 with the 32 banks): a thread reads its whole object and writes it back transposed -- two shared-memory
//     instructions per element, all indices constants;
//   otherwise the 256 consumers walk the OUTPUT elements of the chunk linearly -- thread t takes elements t, t + 256,
//     ... and keeps (object, row, column) up to date with additions only -- and gather each from its transposed position
//     in the stage (consecutive threads write consecutive elements: conflict-free).  When that gather would walk a
//     single bank (an input row of a multiple of 16 words) the producer copies row by row into a padded stage
//     (pitch = rows + 16 bytes: a 4-way conflict instead of a 32-way one).
template <typename E, unsigned CLEN, unsigned CROWS, unsigned CCOLS, bool PER_OBJECT, bool PADDED = false>
__global__ void __launch_bounds__(288)
transpose_staged_kernel(const E *in, E *out, size_t batch, size_t nchunks, unsigned len_arg, unsigned rows_arg,
                        unsigned columns_arg, unsigned G, unsigned stage_bytes, unsigned pitch) {
    constexpr int CONSUMERS = 256, STAGES = 3;
    static_assert(!PER_OBJECT || CLEN != 0, "a thread per object needs the shape at compile time");
    const unsigned len = CLEN ? CLEN : len_arg, rows = CLEN ? CROWS : rows_arg, columns = CLEN ? CCOLS : columns_arg;
    extern __shared__ __align__(128) unsigned char staged_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(staged_smem + (size_t)(STAGES + 2) * stage_bytes);
    uint64_t *empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, CONSUMERS / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == CONSUMERS / 32) {
        if (lane == 0) {
            uint32_t it = 0;
            for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
                const uint32_t s = it % STAGES, use = it / STAGES;
                if (use > 0) mbar_wait_backoff(empty + s, (use - 1) & 1);
                const size_t first = ch * G;
                const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
                if (!PADDED) {
                    // ragged final chunk: whole 16-byte words (the array ends on one: checked by the launcher)
                    const uint32_t bytes = (cnt * len * (uint32_t)sizeof(E) + 15u) & ~15u;
                    mbar_expect_tx(full + s, bytes);
                    bulk_g2s(staged_smem + (size_t)s * stage_bytes, in + first * len, bytes, full + s);
                } else {
                    // padded stage: every input row (rows elements, whole 16-byte words) lands `pitch` elements after
                    // the previous one, so that the gather below does not walk one bank
                    const uint32_t row_bytes = rows * (uint32_t)sizeof(E);
                    mbar_expect_tx(full + s, cnt * columns * row_bytes);
                    // (one copy per loop iteration: a straight-line run of four or more cp.async.bulk, which is what the
                    // unrolled loop becomes, faults with "illegal memory access" on this toolchain / part -- reproduced in
                    // isolation; two or three in a row, as everywhere else in this library, are fine)
#pragma unroll 1
                    for (uint32_t r = 0; r < cnt * columns; ++r)
                        bulk_g2s(staged_smem + (size_t)s * stage_bytes + (size_t)r * pitch * sizeof(E), in + first * len + (size_t)r * rows,
                                 row_bytes, full + s);
                }
            }
        }
        return;
    }
    // how (object, row, column) of the output element move when the linear index advances by 256
    const unsigned step_obj = CONSUMERS / len, step_o = CONSUMERS % len;
    const unsigned step_row = step_o / columns, step_col = step_o % columns;
    uint32_t it = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        const size_t first = ch * G;
        const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
        const E *src = reinterpret_cast<const E *>(staged_smem + (size_t)s * stage_bytes);
        E *cst = reinterpret_cast<E *>(staged_smem + (size_t)(STAGES + (it & 1)) * stage_bytes);
        mbar_wait(full + s, use & 1);
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");  // the bulk store that read this output tile is done
        const unsigned total = cnt * len;
        if constexpr (PER_OBJECT) {
            for (unsigned obj = threadIdx.x; obj < cnt; obj += CONSUMERS) {
                E x[CLEN ? CLEN : 1];
                const E *p = src + obj * CLEN;
                E *q = cst + obj * CLEN;
#pragma unroll
                for (unsigned i = 0; i < CLEN; ++i) x[i] = p[i];
#pragma unroll
                for (unsigned r = 0; r < CROWS; ++r)
#pragma unroll
                    for (unsigned c = 0; c < CCOLS; ++c) q[CCOLS * r + c] = x[CROWS * c + r];
            }
        } else {
            unsigned obj = threadIdx.x / len, o = threadIdx.x % len;
            unsigned orow = o / columns, ocol = o % columns;
            for (unsigned e = threadIdx.x; e < total; e += CONSUMERS) {
                // buffer[columns*row + column] = a[rows*column + row]
                cst[e] = PADDED ? src[(obj * columns + ocol) * pitch + orow] : src[obj * len + ocol * rows + orow];
                obj += step_obj; o += step_o; orow += step_row; ocol += step_col;
                if (ocol >= columns) { ocol -= columns; ++orow; }
                if (o >= len) { o -= len; ++obj; orow -= rows; }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 2, %0;" ::"n"(CONSUMERS) : "memory");
        if (threadIdx.x == 0) {
            const size_t cbytes = (size_t)total * sizeof(E);
            const uint32_t whole = (uint32_t)(cbytes & ~(size_t)15);
            if (whole)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + first * len), "r"(smem_u32(cst)),
                             "r"(whole)
                             : "memory");
            for (size_t e = whole / sizeof(E); e < total; ++e) out[first * len + e] = cst[e];
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace slas
