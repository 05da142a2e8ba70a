// vector_ops.cu -- the bandwidth-bound vector half of the hot path:
//   element-wise add/sub/mul/div   (replaces src/backends/rust.rs:45-73)
//   the normalize scale pass       (replaces src/backends/rust.rs:122-127, src/backends/blas.rs:163-169)
//   batched dot / norm / normalize over [[T; LEN]; batch]
// (the grid-wide dot / norm reductions live in reduce.cu, the fused chains in fused_chain.cu).
//
// Design (B200): 128-bit coalesced streaming loads (ld.global.cs), several independent loads in flight per
// thread.  Element-wise ops and the normalize divide are single IEEE operations (no fast-math, true
// division) => bit-exact against the reference.
#include <algorithm>

#include "async.cuh"
#include "common.cuh"
#include "divby.cuh"
#include "jit.cuh"
#include "kernels.cuh"
#include "p2p.cuh"
#include "reduce.cuh"
#include "staged_objects_kernel.cuh"

namespace slas {

template <int OP, typename T> __device__ __forceinline__ T apply_op(T a, T b) {
    if (OP == 0) return a + b;
    if (OP == 1) return a - b;
    if (OP == 2) return a * b;
    return a / b;  // IEEE div.rn (no --use_fast_math): rust.rs:45-73 divides, it does not multiply by 1/b
}
template <int OP> __device__ __forceinline__ float4 apply_vec(float4 a, float4 b) {
    return make_float4(apply_op<OP>(a.x, b.x), apply_op<OP>(a.y, b.y), apply_op<OP>(a.z, b.z),
                       apply_op<OP>(a.w, b.w));
}
template <int OP> __device__ __forceinline__ double2 apply_vec(double2 a, double2 b) {
    return make_double2(apply_op<OP>(a.x, b.x), apply_op<OP>(a.y, b.y));
}

// ---- element-wise ------------------------------------------------------------------------
constexpr int kEwThreads = 256;

template <typename T, int OP, int UNROLL>
__global__ void __launch_bounds__(kEwThreads)
ewise_vec_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t nvec, size_t len) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    const V *av = reinterpret_cast<const V *>(a);
    const V *bv = reinterpret_cast<const V *>(b);
    V *cv = reinterpret_cast<V *>(c);
    const size_t base = (size_t)blockIdx.x * (kEwThreads * UNROLL) + threadIdx.x;
    V ra[UNROLL], rb[UNROLL];
    if (base + (size_t)(UNROLL - 1) * kEwThreads < nvec) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) ra[u] = ld_stream(av + base + (size_t)u * kEwThreads);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) rb[u] = ld_stream(bv + base + (size_t)u * kEwThreads);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st_stream(cv + base + (size_t)u * kEwThreads, apply_vec<OP>(ra[u], rb[u]));
    } else {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = base + (size_t)u * kEwThreads;
            if (i < nvec) st_stream(cv + i, apply_vec<OP>(ld_stream(av + i), ld_stream(bv + i)));
        }
    }
    // scalar tail (len % N elements), done by the first threads of block 0
    if (blockIdx.x == 0) {
        const size_t i = nvec * N + threadIdx.x;
        if (i < len) c[i] = apply_op<OP>(a[i], b[i]);
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(kEwThreads)
ewise_scalar_kernel(const T *a, const T *b, T *c, size_t len) {
    for (size_t i = (size_t)blockIdx.x * kEwThreads + threadIdx.x; i < len; i += (size_t)gridDim.x * kEwThreads)
        c[i] = apply_op<OP>(a[i], b[i]);
}

// Short vectors (LEN <= ~2^22: one call is a few microseconds of HBM time): ONE wave of CTAs, a multiple of the SM
// count, block c owns the contiguous vectors [c*per, (c+1)*per) so that every SM moves the same number of bytes
// (256 CTAs of 1024 vectors on 148 SMs left 40 SMs with half the work of the others); all loads of a thread are
// issued before its first store.
template <typename T, int OP, int THREADS>
__global__ void __launch_bounds__(THREADS)
ewise_balanced_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t nvec, size_t len, size_t per,
                      int early) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    constexpr int U = 4;
    // programmatic dependent launch (reduce.cuh): let the successor be scheduled; wait for the predecessor first unless
    // the host found that nothing of ours overlaps its operands
    if (!early) pdl_wait();  // a barrier kernel: its successor starts only after everything before this kernel is done
    pdl_trigger();
    const V *av = reinterpret_cast<const V *>(a);
    const V *bv = reinterpret_cast<const V *>(b);
    V *cv = reinterpret_cast<V *>(c);
    size_t lo = (size_t)blockIdx.x * per;
    if (lo > nvec) lo = nvec;
    const size_t end = lo + per < nvec ? lo + per : nvec;
    for (size_t i = lo + threadIdx.x; i < end; i += (size_t)THREADS * U) {
        V ra[U], rb[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t j = i + (size_t)u * THREADS;
            ok[u] = j < end;
            if (ok[u]) ra[u] = ld_stream(av + j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (ok[u]) rb[u] = ld_stream(bv + i + (size_t)u * THREADS);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (ok[u]) st_stream(cv + i + (size_t)u * THREADS, apply_vec<OP>(ra[u], rb[u]));
    }
    if (blockIdx.x == 0) {
        const size_t i = nvec * N + threadIdx.x;
        if (i < len) c[i] = apply_op<OP>(a[i], b[i]);
    }
    if (early) pdl_wait();  // no block leaves before the predecessor is done: completion order stays the stream order
}

template <typename T, int OP>
static int launch_ewise_op(Context *ctx, cudaStream_t st, size_t len, const T *a, const T *b, T *c) {
    constexpr int N = Vec16<T>::N;
    constexpr int UNROLL = 4;
    if (len == 0) return SLAS_OK;
    // views at an odd offset: when all three operands share one misalignment the few elements up to the first 16-byte
    // boundary go through the scalar kernel and the rest through the vector kernel (0.74 -> 1.0 of HBM peak)
    const uintptr_t mis = (uintptr_t)a % 16;
    if (mis && mis == (uintptr_t)b % 16 && mis == (uintptr_t)c % 16 && mis % sizeof(T) == 0) {
        const size_t head = std::min<size_t>((16 - mis) / sizeof(T), len);
        ewise_scalar_kernel<T, OP><<<1, kEwThreads, 0, st>>>(a, b, c, head);
        SLAS_LAUNCH_CHECK();
        return launch_ewise_op<T, OP>(ctx, st, len - head, a + head, b + head, c + head);
    }
    if (aligned16(a) && aligned16(b) && aligned16(c)) {
        const size_t nvec = len / N;
        const size_t sms = (size_t)ctx->sm_count;
        const long small = tunable("ewise_small");  // 0 default, 1 = the tiled kernel everywhere, 2 / 3 = 512 / 1024-thread CTAs
        if (small != 1 && nvec < sms * 8 * (size_t)kEwThreads * UNROLL) {
            const int threads = small == 3 ? 1024 : (small == 2 ? 512 : 256);
            const size_t wave = sms * (size_t)(2048 / threads);
            size_t blocks = (nvec + (size_t)threads * 2 - 1) / ((size_t)threads * 2);
            if (blocks > wave) blocks = wave;
            if (blocks == 0) blocks = 1;
            size_t per = (nvec + blocks - 1) / blocks;
            per = (per + 7) / 8 * 8;  // block ranges start on 128-byte lines
            blocks = per ? (nvec + per - 1) / per : 1;
            if (blocks == 0) blocks = 1;
            const Context::Span reads[2] = {{(const char *)a, (const char *)(a + len)}, {(const char *)b, (const char *)(b + len)}};
            const Context::Span writes[1] = {{(const char *)c, (const char *)(c + len)}};
            const bool own = st == ctx->own_stream || tunable("pdl") == 2;
            const int early = own && hazard_early_ok(ctx, st, reads, 2, writes, 1) ? 1 : 0;
            cudaError_t le;
            if (tunable("pdl") == 1) {
                le = cudaSuccess;
                if (threads == 1024) ewise_balanced_kernel<T, OP, 1024><<<(unsigned)blocks, 1024, 0, st>>>(a, b, c, nvec, len, per, 0);
                else if (threads == 512) ewise_balanced_kernel<T, OP, 512><<<(unsigned)blocks, 512, 0, st>>>(a, b, c, nvec, len, per, 0);
                else ewise_balanced_kernel<T, OP, 256><<<(unsigned)blocks, 256, 0, st>>>(a, b, c, nvec, len, per, 0);
            } else if (threads == 1024) le = launch_pdl(ewise_balanced_kernel<T, OP, 1024>, (unsigned)blocks, 1024u, st, a, b, c, nvec, len, per, early);
            else if (threads == 512) le = launch_pdl(ewise_balanced_kernel<T, OP, 512>, (unsigned)blocks, 512u, st, a, b, c, nvec, len, per, early);
            else le = launch_pdl(ewise_balanced_kernel<T, OP, 256>, (unsigned)blocks, 256u, st, a, b, c, nvec, len, per, early);
            SLAS_CUDA_TRY(le);
            SLAS_LAUNCH_CHECK();
            hazard_record(ctx, st, reads, 2, writes, 1, early);
            return SLAS_OK;
        }
        size_t blocks = (nvec + (size_t)kEwThreads * UNROLL - 1) / ((size_t)kEwThreads * UNROLL);
        if (blocks == 0) blocks = 1;
        SLAS_REQUIRE(blocks <= 0x7fffffffu, "ewise: vector too long (%zu elements)", len);
        ewise_vec_kernel<T, OP, UNROLL><<<(unsigned)blocks, kEwThreads, 0, st>>>(a, b, c, nvec, len);
    } else {
        size_t blocks = (len + kEwThreads - 1) / kEwThreads;
        const size_t cap = (size_t)ctx->sm_count * 32;
        if (blocks > cap) blocks = cap;
        ewise_scalar_kernel<T, OP><<<(unsigned)blocks, kEwThreads, 0, st>>>(a, b, c, len);
    }
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

template <typename T>
int launch_ewise(Context *ctx, cudaStream_t st, int op, size_t len, const T *a, const T *b, T *c) {
    switch (op) {
    case 0: return launch_ewise_op<T, 0>(ctx, st, len, a, b, c);
    case 1: return launch_ewise_op<T, 1>(ctx, st, len, a, b, c);
    case 2: return launch_ewise_op<T, 2>(ctx, st, len, a, b, c);
    case 3: return launch_ewise_op<T, 3>(ctx, st, len, a, b, c);
    }
    return fail(SLAS_ERR_INVALID, "unknown element-wise op %d", op);
}
template int launch_ewise<float>(Context *, cudaStream_t, int, size_t, const float *, const float *, float *);
template int launch_ewise<double>(Context *, cudaStream_t, int, size_t, const double *, const double *, double *);

// ---- normalize scale pass: a[i] = a[i] / *norm (true division, rust.rs:125) --------------------
template <typename T, int UNROLL>
__global__ void __launch_bounds__(kEwThreads)
scale_div_vec_kernel(const T *in, T *a, size_t nvec, size_t len, const T *__restrict__ norm_ptr) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    const T nrm = *norm_ptr;
    DivBy<T> by;
    by.init(nrm);
    V *av = reinterpret_cast<V *>(a);
    const V *iv = reinterpret_cast<const V *>(in);  // == av for the in-place normalize
    const size_t base = (size_t)blockIdx.x * (kEwThreads * UNROLL) + threadIdx.x;
    V r[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const size_t i = base + (size_t)u * kEwThreads;
        if (i < nvec) r[u] = iv[i];  // second touch of the vector: let it hit L2
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const size_t i = base + (size_t)u * kEwThreads;
        if (i < nvec) {
            T *e = reinterpret_cast<T *>(&r[u]);
#pragma unroll
            for (int j = 0; j < N; ++j) e[j] = by.div(e[j]);  // the bits of IEEE div.rn (divby.cuh): not a reciprocal multiply
            st_stream(av + i, r[u]);
        }
    }
    if (blockIdx.x == 0) {
        const size_t i = nvec * N + threadIdx.x;
        if (i < len) a[i] = by.div(in[i]);
    }
}
template <typename T>
__global__ void __launch_bounds__(kEwThreads) scale_div_scalar_kernel(const T *in, T *a, size_t len, const T *norm_ptr) {
    DivBy<T> by;
    by.init(*norm_ptr);
    for (size_t i = (size_t)blockIdx.x * kEwThreads + threadIdx.x; i < len; i += (size_t)gridDim.x * kEwThreads)
        a[i] = by.div(in[i]);
}

template <typename T>
int launch_scale_div(Context *ctx, cudaStream_t st, size_t len, const T *in, T *a, const T *norm_dev) {
    constexpr int N = Vec16<T>::N;
    constexpr int UNROLL = 4;
    if (len == 0) return SLAS_OK;
    const uintptr_t mis = (uintptr_t)a % 16;
    if (mis && mis == (uintptr_t)in % 16 && mis % sizeof(T) == 0) {  // a view at an odd offset: scalar head, vector body
        const size_t head = std::min<size_t>((16 - mis) / sizeof(T), len);
        scale_div_scalar_kernel<T><<<1, kEwThreads, 0, st>>>(in, a, head, norm_dev);
        SLAS_LAUNCH_CHECK();
        return launch_scale_div<T>(ctx, st, len - head, in + head, a + head, norm_dev);
    }
    if (aligned16(a) && aligned16(in)) {
        const size_t nvec = len / N;
        size_t blocks = (nvec + (size_t)kEwThreads * UNROLL - 1) / ((size_t)kEwThreads * UNROLL);
        if (blocks == 0) blocks = 1;
        SLAS_REQUIRE(blocks <= 0x7fffffffu, "normalize: vector too long (%zu elements)", len);
        scale_div_vec_kernel<T, UNROLL><<<(unsigned)blocks, kEwThreads, 0, st>>>(in, a, nvec, len, norm_dev);
    } else {
        size_t blocks = (len + kEwThreads - 1) / kEwThreads;
        const size_t cap = (size_t)ctx->sm_count * 32;
        if (blocks > cap) blocks = cap;
        scale_div_scalar_kernel<T><<<(unsigned)blocks, kEwThreads, 0, st>>>(in, a, len, norm_dev);
    }
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}
template int launch_scale_div<float>(Context *, cudaStream_t, size_t, const float *, float *, const float *);
template int launch_scale_div<double>(Context *, cudaStream_t, size_t, const double *, double *, const double *);

// ---- batched dot / norm / normalize over [[T; len]; batch] -----------------------------------
// One group of GROUP threads per vector (GROUP = 32: a warp per vector, several vectors per
// CTA; GROUP = 256: a CTA per vector).  MODE 0: out[v] = dot(a_v, b_v); 1: out[v] = norm(a_v);
// 2: a_v /= norm(a_v) in place (second pass re-reads the vector from L1/L2, so HBM sees one
// read and one write).
template <typename T, int GROUP, int MODE>
__global__ void __launch_bounds__(256)
batched_reduce_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ out, T *__restrict__ a_mut,
                      size_t batch, size_t len, int vectorised) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    constexpr int GROUPS_PER_CTA = 256 / GROUP;
    __shared__ double red[8];
    const int g = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
    for (size_t v0 = (size_t)blockIdx.x * GROUPS_PER_CTA; v0 < batch; v0 += (size_t)gridDim.x * GROUPS_PER_CTA) {
        const size_t v = v0 + g;
        const bool live = v < batch;
        const T *pa = (MODE == 2 ? a_mut : a) + (live ? v : 0) * len;
        const T *pb = MODE == 0 ? b + (live ? v : 0) * len : pa;
        double s = 0.0;
        // vectorised == 2: vectors of an odd length start at a different misalignment each, shared by both operands (the
        // bases are aligned alike): `head` scalar elements up to the vector's first 16-byte boundary, a 128-bit body, a
        // scalar tail.  vectorised == 1: every vector is whole aligned words (head = 0).
        size_t head = 0;
        if (vectorised == 2) {
            const uintptr_t mis = (uintptr_t)pa % 16;
            head = mis ? (16 - mis) / sizeof(T) : 0;
            if (head > len) head = len;
        }
        if (live) {
            if (vectorised) {
                const V *av = reinterpret_cast<const V *>(pa + head);
                const V *bv = reinterpret_cast<const V *>(pb + head);
                const size_t nvec = (len - head) / N;
                T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
                size_t i = t;
                // four independent 16-byte loads per operand in flight per thread
                for (; i + 3 * GROUP < nvec; i += 4 * GROUP) {
                    V x[4], y[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)  // normalize re-reads its vector: keep it cached (no evict-first hint)
                        x[u] = MODE == 2 ? av[i + u * GROUP] : ld_stream(av + i + u * GROUP);
#pragma unroll
                    for (int u = 0; u < 4; ++u) y[u] = MODE == 0 ? ld_stream(bv + i + u * GROUP) : x[u];
                    const T *px0 = reinterpret_cast<const T *>(&x[0]), *py0 = reinterpret_cast<const T *>(&y[0]);
                    const T *px1 = reinterpret_cast<const T *>(&x[1]), *py1 = reinterpret_cast<const T *>(&y[1]);
                    const T *px2 = reinterpret_cast<const T *>(&x[2]), *py2 = reinterpret_cast<const T *>(&y[2]);
                    const T *px3 = reinterpret_cast<const T *>(&x[3]), *py3 = reinterpret_cast<const T *>(&y[3]);
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        acc0 = fma(px0[j], py0[j], acc0); acc1 = fma(px1[j], py1[j], acc1);
                        acc2 = fma(px2[j], py2[j], acc2); acc3 = fma(px3[j], py3[j], acc3);
                    }
                }
                for (; i < nvec; i += GROUP) {
                    V x0 = av[i];
                    V y0 = MODE == 0 ? bv[i] : x0;
                    const T *px0 = reinterpret_cast<const T *>(&x0), *py0 = reinterpret_cast<const T *>(&y0);
#pragma unroll
                    for (int j = 0; j < N; ++j) acc0 = fma(px0[j], py0[j], acc0);
                }
                acc0 += acc2;
                acc1 += acc3;
                s = (double)acc0 + (double)acc1;
                for (size_t j = head + nvec * N + t; j < len; j += GROUP) s += (double)pa[j] * (double)pb[j];
                if ((size_t)t < head) s += (double)pa[t] * (double)pb[t];
            } else {
                for (size_t j = t; j < len; j += GROUP) s += (double)pa[j] * (double)pb[j];
            }
        }
        if (GROUP == 32) {
            s = warp_sum(s);
        } else {
            s = block_sum<256>(s, red);
            if (threadIdx.x == 0) red[0] = s;
            __syncthreads();
            s = red[0];
            __syncthreads();
        }
        if (!live) continue;
        if (MODE == 0) {
            if (t == 0) out[v] = (T)s;
        } else if (MODE == 1) {
            if (t == 0) out[v] = (T)sqrt(s);
        } else {
            DivBy<T> by;
            by.init((T)sqrt(s));
            T *pm = a_mut + v * len;
            if (vectorised) {
                V *mv = reinterpret_cast<V *>(pm + head);
                const size_t nvec = (len - head) / N;
                size_t i = t;
                for (; i + 3 * GROUP < nvec; i += 4 * GROUP) {  // second touch: L1/L2 hits, four in flight
                    V x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) x[u] = mv[i + u * GROUP];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        T *px = reinterpret_cast<T *>(&x[u]);
#pragma unroll
                        for (int j = 0; j < N; ++j) px[j] = by.div(px[j]);
                        st_stream(mv + i + u * GROUP, x[u]);
                    }
                }
                for (; i < nvec; i += GROUP) {
                    V x = mv[i];
                    T *px = reinterpret_cast<T *>(&x);
#pragma unroll
                    for (int j = 0; j < N; ++j) px[j] = by.div(px[j]);
                    st_stream(mv + i, x);
                }
                for (size_t j = head + nvec * N + t; j < len; j += GROUP) pm[j] = by.div(pm[j]);
                if ((size_t)t < head) pm[t] = by.div(pm[t]);
            } else {
                for (size_t j = t; j < len; j += GROUP) pm[j] = by.div(pm[j]);
            }
        }
    }
}

// Tiny vectors (at most PER 16-byte pieces per lane): LPV lanes per vector, 32 / LPV vectors per warp
// instruction, so a warp still reads 512 consecutive bytes per load; U such groups are in flight per
// warp iteration (PER * U = 8 independent 16-byte loads per operand per lane).  The vector stays in
// registers, so normalize is one read and one write.  Reduction = log2(LPV) shuffles.
template <typename T, int MODE, int PER, int U>
__global__ void __launch_bounds__(256)
batched_reduce_subwarp_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ out,
                              T *__restrict__ a_mut, size_t batch, unsigned len, unsigned lpv) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    const unsigned lane = threadIdx.x & 31, vpw = 32 / lpv;  // vectors per warp instruction
    const unsigned t = lane % lpv, vl = lane / lpv;
    const unsigned nvec = len / N;                           // 16-byte pieces per vector (<= PER * lpv)
    const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 8;
    const T *abase = MODE == 2 ? a_mut : a;
    for (size_t v0 = warp * vpw * U; v0 < batch; v0 += nwarps * vpw * U) {
        V x[U][PER], y[U][PER];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + (size_t)u * vpw + vl;
            live[u] = v < batch;
            const size_t vv = live[u] ? v : batch - 1;
            const V *av = reinterpret_cast<const V *>(abase + vv * len);
#pragma unroll
            for (int p = 0; p < PER; ++p) {
                const unsigned i = t + p * lpv;
                if (i < nvec) x[u][p] = ld_stream(av + i);
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t v = v0 + (size_t)u * vpw + vl;
                const V *bv = reinterpret_cast<const V *>(b + (live[u] ? v : batch - 1) * len);
#pragma unroll
                for (int p = 0; p < PER; ++p) {
                    const unsigned i = t + p * lpv;
                    if (i < nvec) y[u][p] = ld_stream(bv + i);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            T acc = T(0);
#pragma unroll
            for (int p = 0; p < PER; ++p) {
                const unsigned i = t + p * lpv;
                if (i < nvec) {
                    const T *px = reinterpret_cast<const T *>(&x[u][p]);
                    const T *py = MODE == 0 ? reinterpret_cast<const T *>(&y[u][p]) : px;
#pragma unroll
                    for (int j = 0; j < N; ++j) acc = fma(px[j], py[j], acc);
                }
            }
            double s = (double)acc;
            for (unsigned o = 1; o < lpv; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const size_t v = v0 + (size_t)u * vpw + vl;
            if (!live[u]) continue;
            if (MODE == 0) {
                if (t == 0) out[v] = (T)s;
            } else if (MODE == 1) {
                if (t == 0) out[v] = (T)sqrt(s);
            } else {
                DivBy<T> by;
                by.init((T)sqrt(s));
                V *mv = reinterpret_cast<V *>(a_mut + v * len);
#pragma unroll
                for (int p = 0; p < PER; ++p) {
                    const unsigned i = t + p * lpv;
                    if (i < nvec) {
                        T *px = reinterpret_cast<T *>(&x[u][p]);
#pragma unroll
                        for (int j = 0; j < N; ++j) px[j] = by.div(px[j]);
                        st_stream(mv + i, x[u][p]);
                    }
                }
            }
        }
    }
}

// Vectors of at most 256 bytes (a multiple of 32): ONE LANE PER VECTOR.  A lane reads its whole vector with 256-bit
// loads (LDG.E.256: whole 32-byte sectors per request), reduces it in registers -- no shuffles -- and the warp stores
// 32 consecutive results (normalize: scales in registers and writes the vector back with 256-bit stores).
// Measured (fraction of HBM peak, lanes-per-vector kernel above in brackets): LEN 16 f32 norm 0.97 (0.73), normalize
// 0.98 (0.78); LEN 64 f32 norm 0.95 (0.60), normalize 0.81 (0.70); LEN 8 f64 norm 0.91 (0.73), normalize 0.96 (0.87).
__device__ __forceinline__ void ldg256(const void *p, uint32_t (&w)[8]) {
    asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(void *p, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
template <typename T, int MODE, int CHMAX>
__global__ void __launch_bounds__(256)
batched_reduce_lane_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ out, T *__restrict__ a_mut,
                           size_t batch, unsigned chunks /* 32-byte chunks per vector, <= CHMAX */) {
    constexpr int E32 = 32 / (int)sizeof(T);
    const T *abase = MODE == 2 ? a_mut : a;
    const size_t len = (size_t)chunks * E32;
    for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < batch; v += (size_t)gridDim.x * 256) {
        uint32_t xa[CHMAX][8], xb[MODE == 0 ? CHMAX : 1][8];
#pragma unroll
        for (int c = 0; c < CHMAX; ++c)
            if (c < (int)chunks) ldg256(abase + v * len + c * E32, xa[c]);
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < CHMAX; ++c)
                if (c < (int)chunks) ldg256(b + v * len + c * E32, xb[MODE == 0 ? c : 0]);
        }
        T acc0 = T(0), acc1 = T(0);
#pragma unroll
        for (int c = 0; c < CHMAX; ++c)
            if (c < (int)chunks) {
                const T *px = reinterpret_cast<const T *>(xa[c]);
                const T *py = MODE == 0 ? reinterpret_cast<const T *>(xb[MODE == 0 ? c : 0]) : px;
#pragma unroll
                for (int j = 0; j < E32 / 2; ++j) {
                    acc0 = fma(px[j], py[j], acc0);
                    acc1 = fma(px[E32 / 2 + j], py[E32 / 2 + j], acc1);
                }
            }
        const double s = (double)acc0 + (double)acc1;
        if (MODE == 0) {
            out[v] = (T)s;
        } else if (MODE == 1) {
            out[v] = (T)sqrt(s);
        } else {
            DivBy<T> by;
            by.init((T)sqrt(s));
#pragma unroll
            for (int c = 0; c < CHMAX; ++c)
                if (c < (int)chunks) {
                    T *px = reinterpret_cast<T *>(xa[c]);
#pragma unroll
                    for (int j = 0; j < E32; ++j) px[j] = by.div(px[j]);
                    stg256(a_mut + v * len + c * E32, xa[c]);
                }
        }
    }
}

// Short vectors that are NOT whole 16-byte words (3-vectors, LEN 5, 7, 30 ...): the vector kernels above cannot load
// them with aligned vectors and a warp per vector idles on three elements (LEN 3 f32: 0.02 of HBM peak).  Same cure as
// for the small transposes and products: HBM only sees 1-D bulk copies, a thread per vector works in shared memory --
// batched_reduce_staged_kernel, staged_objects_kernel.cuh.  LEN is a template constant for the lengths built in below
// (up to 8 elements) and for whatever else NVRTC specialises at run time (staged_jit.cu); CLEN = 0 is the run-time
// fallback.  All forms run the same operations in the same order: the same bits.
template <typename T, int MODE>
using StagedReduceKernel = void (*)(const T *, const T *, T *, T *, size_t, size_t, unsigned, unsigned, unsigned);
template <typename T, int MODE>
static StagedReduceKernel<T, MODE> staged_reduce_builtin(size_t len) {
    switch (len) {
    case 1: return batched_reduce_staged_kernel<T, MODE, 1>;
    case 3: return batched_reduce_staged_kernel<T, MODE, 3>;
    case 5: return batched_reduce_staged_kernel<T, MODE, 5>;
    case 7: return batched_reduce_staged_kernel<T, MODE, 7>;
    }
    if constexpr (sizeof(T) == 4) {  // (even lengths of 8-byte elements are whole 16-byte words: only LEN 2 is staged)
        if (len == 2) return batched_reduce_staged_kernel<T, MODE, 2>;
        if (len == 4) return batched_reduce_staged_kernel<T, MODE, 4>;
        if (len == 6) return batched_reduce_staged_kernel<T, MODE, 6>;
    } else {
        if (len == 2) return batched_reduce_staged_kernel<T, MODE, 2>;
    }
    return nullptr;
}

template <typename T, int MODE>
static int launch_batched_reduce_staged(Context *ctx, cudaStream_t st, size_t batch, size_t len, const T *a, const T *b, T *out,
                                        T *a_mut, bool *handled) {
    *handled = false;
    const T *base = MODE == 2 ? a_mut : a;
    const size_t vbytes = len * sizeof(T);
    // longer vectors leave too few of them per chunk to keep a thread-per-vector CTA busy: measured crossover against
    // the warp-per-vector kernel (profiles/r1e_tune_short_vectors.log) ~800 bytes for dot / norm (LEN 127 f32: 0.88 / 0.76
    // against 0.58 / 0.32), ~600 bytes for normalize
    // (norm and normalize must switch at the same length: normalize divides by the bits norm returns)
    const size_t limit = tunable("bdot_limit") > 0 ? (size_t)tunable("bdot_limit") : (MODE == 0 ? 832 : 640);
    if (len == 0 || vbytes > limit || batch < 256 || !aligned16(base) || (MODE == 0 && !aligned16(b)) || tunable("bdot_variant") == 2)
        return SLAS_OK;
    size_t galign = 16;
    while (galign > 1 && vbytes % (16 / (galign / 2)) == 0) galign /= 2;  // 16 / gcd(vbytes, 16): chunks start on 16-byte words
    // the kernel streams whole 16-byte words: a batch that does not end on one (1001 3-vectors ...) keeps its last
    // < galign vectors for the warp-per-vector kernel instead of falling off the fast path altogether
    const size_t rem = batch % galign;
    if (rem) {
        const size_t main = batch - rem;
        if (main < 256) return SLAS_OK;
        SLAS_TRY((launch_batched_reduce_staged<T, MODE>(ctx, st, main, len, a, b, out, a_mut, handled)));
        if (!*handled) return SLAS_OK;
        batched_reduce_kernel<T, 32, MODE><<<(unsigned)((rem + 7) / 8), 256, 0, st>>>(
            a ? a + main * len : nullptr, b ? b + main * len : nullptr, out ? out + main : nullptr, a_mut ? a_mut + main * len : nullptr,
            rem, len, 0);
        SLAS_LAUNCH_CHECK();
        return SLAS_OK;
    }
    // measured (profiles/r2w_tune_oddsweep3.log): the cooperative scale pass of normalize (LEN > 8, two more shared tiles
    // than dot) prefers three CTAs per SM -- 12 KB stages: LEN 13 / 30 / 38 f32 0.84 / 0.85 / 0.91 against 0.80 / 0.81 / 0.84
    // at 16 KB, LEN 19 f64 0.92 against 0.75
    // ... and for vectors of up to 160 bytes dot prefers 4 KB stages per operand and norm 8 KB (four CTAs per SM either
    // way: profiles/r3e_tune_oddsweep4.log -- LEN 2 / 3 / 30 f32 norm 0.87 / 0.90 / 0.91 at 16 KB, 0.96 / 0.96 / 0.95 at 8 KB;
    // dot 0.93 / 0.98 / 0.98 -> 0.98 / 1.01 / 1.00); longer ones lose at small stages (LEN 127: 0.95 -> 0.61)
    const long small_kb = MODE == 0 ? 4 : 8;
    const long stage_kb = tunable("bdot_stage_kb") > 0 ? tunable("bdot_stage_kb")
                          : (MODE == 2 ? (len > 8 ? 12 : 16) : (vbytes <= 160 ? small_kb : 16));
    size_t G = ((size_t)stage_kb * 1024 / vbytes) / galign * galign;
    const size_t gcap = tunable("bdot_gcap") > 0 ? (size_t)tunable("bdot_gcap") : 4096;
    if (G > gcap) G = gcap / galign * galign;
    if (G < galign) G = galign;
    size_t stage_bytes = 0, smem = 0;
    for (;; G = (G / 2 + galign - 1) / galign * galign) {  // (a stage-size tunable beyond what fits: halve)
        stage_bytes = (G * vbytes + 127) / 128 * 128;
        smem = (3 * (MODE == 0 ? 2 : 1) + (MODE == 2 ? 2 : 0)) * stage_bytes + 6 * sizeof(uint64_t) + (MODE == 2 && len > 8 ? G * sizeof(DivBy<T>) : 0);  // (the divisors of the cooperative scale pass)
        if (smem <= 200 * 1024 || G <= galign) break;
    }
    const size_t nchunks = (batch + G - 1) / G;
    int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
    const int cta_cap = tunable("staged_cta_cap") > 0 ? (int)tunable("staged_cta_cap") : 4;
    if (per_sm > cta_cap) per_sm = cta_cap;
    if (per_sm < 1) per_sm = 1;
    size_t grid = (size_t)ctx->sm_count * per_sm;
    if (grid > nchunks) grid = nchunks;
    // shape as a compile-time constant: built in, else specialised by NVRTC (lengths that fit a thread's unrolled loop),
    // else the run-time-shape instantiation.  tunable staged_jit: 1 = no NVRTC, 2 = run-time shape only.
    const long jit_mode = tunable("staged_jit");
    StagedReduceKernel<T, MODE> kern = jit_mode == 2 ? nullptr : staged_reduce_builtin<T, MODE>(len);
    if (!kern && jit_mode == 0 && smem <= kStagedJitSmem) {
        if (CUfunction fn = staged_reduce_jit(ctx, sizeof(T), MODE, (unsigned)len)) {
            size_t batch_arg = batch, nchunks_arg = nchunks;
            unsigned len_arg = (unsigned)len, g_arg = (unsigned)G, stage_arg = (unsigned)stage_bytes;
            void *params[] = {(void *)&a, (void *)&b, (void *)&out, (void *)&a_mut, (void *)&batch_arg, (void *)&nchunks_arg,
                              (void *)&len_arg, (void *)&g_arg, (void *)&stage_arg};
            SLAS_TRY(jit_launch(fn, (unsigned)grid, 288, smem, st, params));
            *handled = true;
            return SLAS_OK;
        }
    }
    if (!kern) kern = batched_reduce_staged_kernel<T, MODE, 0>;
    SLAS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, 288, smem, st>>>(a, b, out, a_mut, batch, nchunks, (unsigned)len, (unsigned)G, (unsigned)stage_bytes);
    SLAS_LAUNCH_CHECK();
    *handled = true;
    return SLAS_OK;
}

// Long vectors: every vector is cut into segments of kSegElems elements, one CTA per segment, so the
// grid fills the chip even for a handful of vectors.  partial[v][s] (f64) -> batched_finish_kernel adds
// the segments of a vector in index order (deterministic) -> out / norms.
constexpr size_t kSegBytes = 256 * 1024;

template <typename T, int MODE /* 0 dot, 1 squared norm */>
__global__ void __launch_bounds__(256)
batched_segment_kernel(const T *__restrict__ a, const T *__restrict__ b, double *__restrict__ partial, size_t len,
                       unsigned segs, size_t seg_elems) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    __shared__ double red[8];
    const size_t v = blockIdx.x / segs;
    const unsigned sg = blockIdx.x % segs;
    const size_t lo = (size_t)sg * seg_elems, hi = lo + seg_elems < len ? lo + seg_elems : len;
    const V *av = reinterpret_cast<const V *>(a + v * len + lo);
    const V *bv = reinterpret_cast<const V *>((MODE == 0 ? b : a) + v * len + lo);
    const size_t nvec = (hi - lo) / N;
    T acc[4] = {T(0), T(0), T(0), T(0)};
    size_t i = threadIdx.x;
    for (; i + 3 * 256 < nvec; i += 4 * 256) {
        V x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = ld_stream(av + i + u * 256);
#pragma unroll
        for (int u = 0; u < 4; ++u) y[u] = MODE == 0 ? ld_stream(bv + i + u * 256) : x[u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const T *px = reinterpret_cast<const T *>(&x[u]), *py = reinterpret_cast<const T *>(&y[u]);
#pragma unroll
            for (int j = 0; j < N; ++j) acc[u] = fma(px[j], py[j], acc[u]);
        }
    }
    for (; i < nvec; i += 256) {
        const V x = ld_stream(av + i);
        const V y = MODE == 0 ? ld_stream(bv + i) : x;
        const T *px = reinterpret_cast<const T *>(&x), *py = reinterpret_cast<const T *>(&y);
#pragma unroll
        for (int j = 0; j < N; ++j) acc[0] = fma(px[j], py[j], acc[0]);
    }
    double s = ((double)acc[0] + (double)acc[1]) + ((double)acc[2] + (double)acc[3]);
    // ragged end of the vector (len % N elements) belongs to the last segment
    if (sg == segs - 1) {
        const T *pa = a + v * len, *pb = (MODE == 0 ? b : a) + v * len;
        for (size_t j = lo + nvec * N + threadIdx.x; j < len; j += 256) s += (double)pa[j] * (double)pb[j];
    }
    s = block_sum<256>(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

template <typename T>
__global__ void __launch_bounds__(256)
batched_finish_kernel(const double *__restrict__ partial, T *__restrict__ out, size_t batch, unsigned segs, int take_sqrt) {
    for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < batch; v += (size_t)gridDim.x * 256) {
        double s = 0.0;
        for (unsigned j = 0; j < segs; ++j) s += partial[v * segs + j];
        out[v] = take_sqrt ? (T)sqrt(s) : (T)s;
    }
}

// a[v][i] /= norms[v] over the whole batch (true division)
template <typename T>
__global__ void __launch_bounds__(256)
batched_scale_kernel(T *__restrict__ a, const T *__restrict__ norms, size_t len, size_t nvec_per, size_t total_vec) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total_vec; i += (size_t)gridDim.x * 256) {
        const size_t v = i / nvec_per, w = i - v * nvec_per;
        V *p = reinterpret_cast<V *>(a + v * len) + w;
        V x = *p;
        DivBy<T> by;
        by.init(norms[v]);
        T *px = reinterpret_cast<T *>(&x);
#pragma unroll
        for (int j = 0; j < N; ++j) px[j] = by.div(px[j]);
        st_stream(p, x);
    }
}

template <typename T, int MODE>
static int launch_batched_segmented(Context *ctx, cudaStream_t st, size_t batch, size_t len, const T *a, const T *b,
                                    T *out, T *a_mut) {
    constexpr int N = Vec16<T>::N;
    const size_t seg_elems = kSegBytes / sizeof(T);
    const size_t segs = (len + seg_elems - 1) / seg_elems;
    SLAS_REQUIRE(batch * segs <= 0x7fffffffu && segs <= 0xffffffffu, "batched reduction: too many segments");
    void *ws = nullptr;
    SLAS_TRY(ensure_workspace(ctx, batch * segs * sizeof(double) + batch * sizeof(T) + 256, &ws));
    double *partial = (double *)ws;
    T *norms = (T *)((char *)ws + ((batch * segs * sizeof(double) + 255) & ~(size_t)255));
    const T *src = MODE == 2 ? a_mut : a;
    if (MODE == 0) batched_segment_kernel<T, 0><<<(unsigned)(batch * segs), 256, 0, st>>>(src, b, partial, len, (unsigned)segs, seg_elems);
    else batched_segment_kernel<T, 1><<<(unsigned)(batch * segs), 256, 0, st>>>(src, src, partial, len, (unsigned)segs, seg_elems);
    SLAS_LAUNCH_CHECK();
    size_t fb = (batch + 255) / 256;
    if (fb > (size_t)ctx->sm_count * 8) fb = (size_t)ctx->sm_count * 8;
    batched_finish_kernel<T><<<(unsigned)fb, 256, 0, st>>>(partial, MODE == 2 ? norms : out, batch, (unsigned)segs, MODE != 0);
    SLAS_LAUNCH_CHECK();
    if (MODE == 2) {
        const size_t nvec_per = len / N, total = batch * nvec_per;
        size_t sb = (total + 255) / 256;
        if (sb > (size_t)ctx->sm_count * 32) sb = (size_t)ctx->sm_count * 32;
        batched_scale_kernel<T><<<(unsigned)sb, 256, 0, st>>>(a_mut, norms, len, nvec_per, total);
        SLAS_LAUNCH_CHECK();
    }
    return SLAS_OK;
}

template <typename T, int MODE>
static int launch_batched_reduce_mode(Context *ctx, cudaStream_t st, size_t batch, size_t len, const T *a,
                                      const T *b, T *out, T *a_mut) {
    if (batch == 0) return SLAS_OK;
    const T *base = MODE == 2 ? a_mut : a;
    const bool vec = aligned16(base) && (MODE != 0 || aligned16(b)) && (len * sizeof(T)) % 16 == 0;
    const size_t cap = (size_t)ctx->sm_count * 8;
    constexpr int N = Vec16<T>::N;
    // a lane per vector, 256-bit accesses: up to 256 bytes for norm / normalize (which must share a kernel: normalize
    // divides by the bits norm returns), up to 128 bytes for dot (two operands: 256-byte vectors cost it its occupancy;
    // measured, LEN 32 f64: 0.70 against 1.02 for the lanes-per-vector kernel below)
    if (vec && len > 0 && (len * sizeof(T)) % 32 == 0 && len * sizeof(T) <= (MODE == 0 ? 128 : 256) && (uintptr_t)base % 32 == 0 &&
        (MODE != 0 || (uintptr_t)b % 32 == 0) && tunable("bdot_variant") != 1 && tunable("bdot_variant") != 4) {
        const unsigned chunks = (unsigned)(len * sizeof(T) / 32);
        size_t blocks = (batch + 255) / 256;
        const size_t lcap = (size_t)ctx->sm_count * 16;
        if (blocks > lcap) blocks = lcap;
        if (chunks <= 1) batched_reduce_lane_kernel<T, MODE, 1><<<(unsigned)blocks, 256, 0, st>>>(a, b, out, a_mut, batch, chunks);
        else if (chunks <= 2) batched_reduce_lane_kernel<T, MODE, 2><<<(unsigned)blocks, 256, 0, st>>>(a, b, out, a_mut, batch, chunks);
        else if (chunks <= 4) batched_reduce_lane_kernel<T, MODE, 4><<<(unsigned)blocks, 256, 0, st>>>(a, b, out, a_mut, batch, chunks);
        else batched_reduce_lane_kernel<T, MODE, 8><<<(unsigned)blocks, 256, 0, st>>>(a, b, out, a_mut, batch, chunks);
        SLAS_LAUNCH_CHECK();
        return SLAS_OK;
    }
    if (!vec || len <= (size_t)N || tunable("bdot_variant") == 4) {
        // not whole 16-byte words (3-vectors, LEN 5, 30 ...), and ONE 16-byte word (4-vectors f32, 2-vectors f64: normalize
        // 0.84 -> 0.90, norm 0.94 -> 0.99 against the sub-warp kernel; longer power-of-two vectors would walk a single
        // shared-memory bank there: LEN 16 f32 0.34, profiles/r3r_tune_lanevsstaged.log): bulk-TMA staged, a thread per vector
        bool staged = false;
        SLAS_TRY((launch_batched_reduce_staged<T, MODE>(ctx, st, batch, len, a, b, out, a_mut, &staged)));
        if (staged) return SLAS_OK;
    }
    if constexpr (MODE != 0) {
        // 8 KB .. 512 KB per vector: a thread-block cluster per vector, the vector in registers between norm and division
        // (norm switches with normalize: normalize divides by the bits norm returns)
        bool clustered = false;
        SLAS_TRY((launch_batched_norm_cluster<T, MODE>(ctx, st, batch, len, a, out, a_mut, &clustered)));
        if (clustered) return SLAS_OK;
    }
    if (vec && len >= (size_t)N && len / N <= 32 * 8) {
        // tiny: sub-warp groups, registers only
        const size_t pieces = len / N;
        unsigned lpv = 1;
        while (lpv < 32 && (size_t)lpv * 2 <= pieces) lpv *= 2;  // largest power of two <= pieces: every lane has work
        const size_t per = (pieces + lpv - 1) / lpv;             // 1 .. PER
        const size_t vec_per_cta = (size_t)8 * (32 / lpv);
        auto blocks_for = [&](int u) {
            size_t blocks = (batch + vec_per_cta * u - 1) / (vec_per_cta * u);
            return (unsigned)(blocks > cap ? cap : blocks);
        };
        if (per <= 1) batched_reduce_subwarp_kernel<T, MODE, 1, 8><<<blocks_for(8), 256, 0, st>>>(a, b, out, a_mut, batch, (unsigned)len, lpv);
        else if (per <= 2) batched_reduce_subwarp_kernel<T, MODE, 2, 4><<<blocks_for(4), 256, 0, st>>>(a, b, out, a_mut, batch, (unsigned)len, lpv);
        else if (per <= 4) batched_reduce_subwarp_kernel<T, MODE, 4, 2><<<blocks_for(2), 256, 0, st>>>(a, b, out, a_mut, batch, (unsigned)len, lpv);
        else batched_reduce_subwarp_kernel<T, MODE, 8, 1><<<blocks_for(1), 256, 0, st>>>(a, b, out, a_mut, batch, (unsigned)len, lpv);
        SLAS_LAUNCH_CHECK();
        return SLAS_OK;
    }
    if (vec && len * sizeof(T) >= 2 * kSegBytes) return launch_batched_segmented<T, MODE>(ctx, st, batch, len, a, b, out, a_mut);
    // (normalize: a register-resident variant -- vector held in registers between the reduction and the scale pass --
    // measured SLOWER than re-reading from L2, 0.50-0.81 against 0.79-0.98 of HBM peak for 1.5K..32K elements:
    // 160+ registers per thread leave one CTA per SM and its load / reduce / store phases do not overlap)
    // odd lengths: every vector peels its own head when both operands' bases are misaligned alike (element-aligned)
    const bool peel = !vec && len >= 64 && (uintptr_t)base % sizeof(T) == 0 &&
                      (MODE != 0 || (uintptr_t)base % 16 == (uintptr_t)b % 16) && tunable("bdot_variant") != 3;
    const int vmode = vec ? 1 : (peel ? 2 : 0);
    if (len <= 8192) {
        size_t blocks = (batch + 7) / 8;
        if (blocks > cap) blocks = cap;
        batched_reduce_kernel<T, 32, MODE><<<(unsigned)blocks, 256, 0, st>>>(a, b, out, a_mut, batch, len, vmode);
    } else {
        size_t blocks = batch;
        if (blocks > cap) blocks = cap;
        batched_reduce_kernel<T, 256, MODE><<<(unsigned)blocks, 256, 0, st>>>(a, b, out, a_mut, batch, len, vmode);
    }
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

template <typename T>
int launch_batched_reduce(Context *ctx, cudaStream_t st, int mode, size_t batch, size_t len, const T *a,
                          const T *b, T *out, T *a_mut) {
    switch (mode) {
    case 0: return launch_batched_reduce_mode<T, 0>(ctx, st, batch, len, a, b, out, a_mut);
    case 1: return launch_batched_reduce_mode<T, 1>(ctx, st, batch, len, a, b, out, a_mut);
    case 2: return launch_batched_reduce_mode<T, 2>(ctx, st, batch, len, a, b, out, a_mut);
    }
    return fail(SLAS_ERR_INVALID, "unknown batched reduction mode %d", mode);
}
template int launch_batched_reduce<float>(Context *, cudaStream_t, int, size_t, size_t, const float *,
                                          const float *, float *, float *);
template int launch_batched_reduce<double>(Context *, cudaStream_t, int, size_t, size_t, const double *,
                                           const double *, double *, double *);

// ---- synthetic fill (counter-based; bench/test plumbing for vectors too big to stage) ------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
template <typename T>
__global__ void __launch_bounds__(256) fill_uniform_kernel(T *x, size_t len, uint64_t seed, uint64_t offset, T lo, T hi) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < len; i += (size_t)gridDim.x * 256) {
        const uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + offset + i);
        const T u = sizeof(T) == 4 ? (T)((float)(h >> 40) * (1.0f / 16777216.0f))
                                   : (T)((double)(h >> 11) * (1.0 / 9007199254740992.0));
        x[i] = lo + (hi - lo) * u;
    }
}
template <typename T>
int launch_fill_uniform(Context *ctx, cudaStream_t st, size_t len, T *x, uint64_t seed, uint64_t offset, T lo, T hi) {
    if (len == 0) return SLAS_OK;
    size_t blocks = (len + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    fill_uniform_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(x, len, seed, offset, lo, hi);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}
template int launch_fill_uniform<float>(Context *, cudaStream_t, size_t, float *, uint64_t, uint64_t, float, float);
template int launch_fill_uniform<double>(Context *, cudaStream_t, size_t, double *, uint64_t, uint64_t, double, double);

}  // namespace slas
