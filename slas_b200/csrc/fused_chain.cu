// fused_chain.cu -- fused chains: sequences a slas user writes as separate ops, in one pass over HBM.
// slas leaves lazy execution to its users (CONTRIBUTING.md:7-22), so these are extra entry points.  Same launch
// geometry, vector tiling, accumulator layout and combine order as reduce_kernel (reduce.cuh), and the element-wise
// step is the same single IEEE operation as ewise_vec_kernel => the results carry the same bits as the unfused
// chain on this backend.
//
// CHAIN 0..3: x = a (+,-,*,/) b  [written to c when WRITE_C]; result = sum x*d      (add -> dot; rust.rs:45-73, 20-41)
// CHAIN 4:    results = sum a*b, sqrt(sum a*a), sqrt(sum b*b)                      (dot + both norms; rust.rs:116-121)
#include "common.cuh"
#include "kernels.cuh"
#include "reduce.cuh"

namespace slas {

template <int OP, typename T> __device__ __forceinline__ T chain_op(T a, T b) {
    if (OP == 0) return a + b;
    if (OP == 1) return a - b;
    if (OP == 2) return a * b;
    return a / b;  // IEEE div.rn (no --use_fast_math)
}
template <int OP> __device__ __forceinline__ float4 chain_vec(float4 a, float4 b) {
    return make_float4(chain_op<OP>(a.x, b.x), chain_op<OP>(a.y, b.y), chain_op<OP>(a.z, b.z), chain_op<OP>(a.w, b.w));
}
template <int OP> __device__ __forceinline__ double2 chain_vec(double2 a, double2 b) {
    return make_double2(chain_op<OP>(a.x, b.x), chain_op<OP>(a.y, b.y));
}

template <typename T, int CHAIN, bool WRITE_C, int THREADS>
__global__ void __launch_bounds__(THREADS)
fused_chain_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, const T *__restrict__ d, size_t len,
                   size_t nvec, int vectorised, const RedGeom g, const RedScratch sc, T *__restrict__ typed_out) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    constexpr int NRES = CHAIN == 4 ? 3 : 1;
    constexpr int OP = CHAIN < 4 ? CHAIN : 0;
    constexpr int U = RedTile<THREADS>::U;  // reduce_kernel's tiling: same element-to-accumulator map, same bits
    __shared__ double red[THREADS / 32];
    __shared__ bool is_last;
    Reducer<T, 0> r0, r1, r2;
    r0.clear(); r1.clear(); r2.clear();
    const size_t tid = (size_t)blockIdx.x * THREADS + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * THREADS;
    if (vectorised) {
        const V *av = reinterpret_cast<const V *>(a);
        const V *bv = reinterpret_cast<const V *>(b);
        const V *dv = reinterpret_cast<const V *>(d);
        V *cv = reinterpret_cast<V *>(c);
        size_t i, end, stride;
        red_range<THREADS>(g, nvec, i, end, stride);
        int since_flush = 0;
        auto use = [&](int u, size_t j, V x, V y, V z) {
            if (CHAIN == 4) {
                r0.add(u, x, y); r1.add(u, x, x); r2.add(u, y, y);
            } else {
                const V e = chain_vec<OP>(x, y);
                if (WRITE_C) st_stream(cv + j, e);
                r0.add(u, e, z);
            }
        };
        for (; i + (size_t)(U - 1) * THREADS < end; i += stride) {
            V x[U], y[U], z[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t j = i + (size_t)u * THREADS;
                x[u] = ld_stream(av + j);
                y[u] = ld_stream(bv + j);
                z[u] = CHAIN == 4 ? x[u] : ld_stream(dv + j);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) use(u, i + (size_t)u * THREADS, x[u], y[u], z[u]);
            if (++since_flush == kFlushEvery) {
                r0.flush();
                if (CHAIN == 4) { r1.flush(); r2.flush(); }
                since_flush = 0;
            }
        }
        {
            V x[U], y[U], z[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t j = i + (size_t)u * THREADS;
                ok[u] = j < end;
                x[u] = ok[u] ? ld_stream(av + j) : vzero((V *)nullptr);
                y[u] = ok[u] ? ld_stream(bv + j) : vzero((V *)nullptr);
                z[u] = (CHAIN != 4 && ok[u]) ? ld_stream(dv + j) : x[u];
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) use(u, i + (size_t)u * THREADS, x[u], y[u], z[u]);
        }
        r0.flush();
        if (CHAIN == 4) { r1.flush(); r2.flush(); }
        const size_t j = nvec * N + tid;
        if (j < len) {
            if (CHAIN == 4) {
                r0.add_scalar(a[j], b[j]); r1.add_scalar(a[j], a[j]); r2.add_scalar(b[j], b[j]);
            } else {
                const T e = chain_op<OP>(a[j], b[j]);
                if (WRITE_C) c[j] = e;
                r0.add_scalar(e, d[j]);
            }
        }
    } else {
        for (size_t j = tid; j < len; j += nthreads) {
            if (CHAIN == 4) {
                r0.add_scalar(a[j], b[j]); r1.add_scalar(a[j], a[j]); r2.add_scalar(b[j], b[j]);
            } else {
                const T e = chain_op<OP>(a[j], b[j]);
                if (WRITE_C) c[j] = e;
                r0.add_scalar(e, d[j]);
            }
        }
    }
    double s[NRES], p[NRES];
    s[0] = block_sum<THREADS>(r0.sum, red);
    if (CHAIN == 4) { s[NRES > 1 ? 1 : 0] = block_sum<THREADS>(r1.sum, red); s[NRES > 2 ? 2 : 0] = block_sum<THREADS>(r2.sum, red); }
    if (!grid_combine<THREADS, NRES>(g, sc, s, p, red, &is_last)) return;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NRES; ++q) typed_out[q] = (CHAIN == 4 && q > 0) ? (T)sqrt(p[q]) : (T)p[q];
    }
}

template <typename T, int CHAIN, bool WRITE_C>
static void launch_fused_threads(const RedGeom &g, cudaStream_t st, const T *a, const T *b, T *c, const T *d, size_t len,
                                 size_t nvec, int vec, const RedScratch &sc, T *typed_out) {
    if (g.threads == 1024)
        fused_chain_kernel<T, CHAIN, WRITE_C, 1024><<<g.blocks, 1024, 0, st>>>(a, b, c, d, len, nvec, vec, g, sc, typed_out);
    else if (g.threads == 512)
        fused_chain_kernel<T, CHAIN, WRITE_C, 512><<<g.blocks, 512, 0, st>>>(a, b, c, d, len, nvec, vec, g, sc, typed_out);
    else
        fused_chain_kernel<T, CHAIN, WRITE_C, 256><<<g.blocks, 256, 0, st>>>(a, b, c, d, len, nvec, vec, g, sc, typed_out);
}

template <typename T, int CHAIN>
static int launch_fused_chain_kind(Context *ctx, cudaStream_t st, size_t len, const T *a, const T *b, T *c, const T *d,
                                   T *typed_out) {
    constexpr int N = Vec16<T>::N;
    const bool vec = aligned16(a) && aligned16(b) && (CHAIN == 4 || aligned16(d)) && (c == nullptr || aligned16(c));
    const size_t nvec = vec ? len / N : 0;
    const RedGeom g = reduce_geometry(ctx, vec ? nvec : len, vec);  // reduce_kernel's geometry: same bits as the unfused chain
    const RedScratch sc = reduce_scratch(ctx);
    if (c) launch_fused_threads<T, CHAIN, true>(g, st, a, b, c, d, len, nvec, vec ? 1 : 0, sc, typed_out);
    else launch_fused_threads<T, CHAIN, false>(g, st, a, b, c, d, len, nvec, vec ? 1 : 0, sc, typed_out);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

template <typename T>
int launch_fused_chain(Context *ctx, cudaStream_t st, int chain, size_t len, const T *a, const T *b, T *c, const T *d,
                       T *typed_out) {
    switch (chain) {
    case 0: return launch_fused_chain_kind<T, 0>(ctx, st, len, a, b, c, d, typed_out);
    case 1: return launch_fused_chain_kind<T, 1>(ctx, st, len, a, b, c, d, typed_out);
    case 2: return launch_fused_chain_kind<T, 2>(ctx, st, len, a, b, c, d, typed_out);
    case 3: return launch_fused_chain_kind<T, 3>(ctx, st, len, a, b, c, d, typed_out);
    case 4: return launch_fused_chain_kind<T, 4>(ctx, st, len, a, b, nullptr, nullptr, typed_out);
    }
    return fail(SLAS_ERR_INVALID, "unknown fused chain %d", chain);
}
template int launch_fused_chain<float>(Context *, cudaStream_t, int, size_t, const float *, const float *, float *, const float *,
                                       float *);
template int launch_fused_chain<double>(Context *, cudaStream_t, int, size_t, const double *, const double *, double *,
                                        const double *, double *);

}  // namespace slas
