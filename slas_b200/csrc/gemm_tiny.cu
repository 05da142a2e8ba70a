// gemm_tiny.cu -- batched matrix_mul for static shapes whose every dimension is in 1..4 (2x3 . 3x2 of the
// reference's own tests, tests/src/main.rs:289-352; 3x3 . 3x3; 4x4 . 4x1 ...): ONE THREAD PER MATRIX.
// Replaces the same loop of Matrix::matrix_mul calls as gemm_small.cu (src/tensor.rs:441-454 ->
// cblas_{s,d}gemm, src/backends/blas.rs:94-109); 4x4x4 keeps its own configuration there.
//
// These products are pure HBM streaming (<= 0.67 flop/byte), so the memory side is the one of gemm_small.cu:
// persistent CTAs, a producer warp that streams chunks of 256 consecutive matrices of A and of B into a
// three-stage shared-memory ring with 1-D bulk TMA copies completing on mbarriers, and C leaving through a
// double-buffered shared stage with one bulk store per chunk (full-line writes).  The compute side is
// register-only: a thread reads its A and B with the widest shared-memory vector loads the object size allows
// (a scalar read pattern with a stride of 8 or 16 words would be an 8- / 16-way bank conflict), forms the
// M x N results as ascending-k FMA chains -- the same order as every other GEMM kernel here -- and writes them
// to the C stage as vectors.  a_trans / b_trans select one of four fully unrolled bodies (uniform branch),
// so one kernel per (T, M, N, K) serves all flag combinations.
#include "async.cuh"
#include "common.cuh"
#include "kernels.cuh"

namespace slas {

// CNT contiguous elements between shared memory and registers with the widest vectors that divide the object
template <typename T, int CNT>
__device__ __forceinline__ void lds_linear(T (&dst)[CNT], const T *p) {
    constexpr int BYTES = CNT * (int)sizeof(T);
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(p)[i];
    } else if constexpr (BYTES % 8 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 8; ++i) reinterpret_cast<uint2 *>(dst)[i] = reinterpret_cast<const uint2 *>(p)[i];
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) dst[i] = p[i];
    }
}
template <typename T, int CNT>
__device__ __forceinline__ void sts_linear(T *p, const T (&src)[CNT]) {
    constexpr int BYTES = CNT * (int)sizeof(T);
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) reinterpret_cast<uint4 *>(p)[i] = reinterpret_cast<const uint4 *>(src)[i];
    } else if constexpr (BYTES % 8 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 8; ++i) reinterpret_cast<uint2 *>(p)[i] = reinterpret_cast<const uint2 *>(src)[i];
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) p[i] = src[i];
    }
}

template <typename T, int M, int N, int K, bool TA, bool TB>
__device__ __forceinline__ void tiny_product(const T (&la)[M * K], const T (&lb)[K * N], T (&lc)[M * N]) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            T acc = T(0);
#pragma unroll
            for (int p = 0; p < K; ++p) acc = fma(TA ? la[p * M + i] : la[i * K + p], TB ? lb[j * K + p] : lb[p * N + j], acc);
            lc[i * N + j] = acc;
        }
}

template <typename T, int M, int N, int K>
__global__ void __launch_bounds__(288)
gemm_tiny_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t batch, size_t nchunks, int ta,
                 int tb, unsigned stage_elems, unsigned cstage_elems) {
    constexpr int G = 256, CONSUMERS = 256, STAGES = 3, MK = M * K, KN = K * N, MN = M * N;
    extern __shared__ __align__(128) unsigned char smem[];
    T *stage_base = reinterpret_cast<T *>(smem);
    T *cstage_base = stage_base + (size_t)STAGES * stage_elems;
    uint64_t *full = reinterpret_cast<uint64_t *>(cstage_base + 2 * (size_t)cstage_elems);
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
                if (use > 0) mbar_wait(empty + s, (use - 1) & 1);
                const size_t first = ch * G;
                const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
                T *sa = stage_base + (size_t)s * stage_elems;
                T *sb = sa + (size_t)G * MK;
                // ragged final chunk: whole 16-byte words (the arrays end on one: checked by the launcher)
                const uint32_t bytes_a = (cnt * MK * (uint32_t)sizeof(T) + 15u) & ~15u;
                const uint32_t bytes_b = (cnt * KN * (uint32_t)sizeof(T) + 15u) & ~15u;
                mbar_expect_tx(full + s, bytes_a + bytes_b);
                bulk_g2s(sa, a + first * MK, bytes_a, full + s);
                bulk_g2s(sb, b + first * KN, bytes_b, full + s);
            }
        }
        return;
    }
    uint32_t it = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        const size_t first = ch * G;
        const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
        const T *ma = stage_base + (size_t)s * stage_elems + (size_t)threadIdx.x * MK;
        const T *mb = stage_base + (size_t)s * stage_elems + (size_t)G * MK + (size_t)threadIdx.x * KN;
        T *cst = cstage_base + (size_t)(it & 1) * cstage_elems;
        mbar_wait(full + s, use & 1);
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");  // the bulk store that read this C stage is done
        if (threadIdx.x < cnt) {
            alignas(16) T la[MK];
            alignas(16) T lb[KN];
            alignas(16) T lc[MN];
            lds_linear<T, MK>(la, ma);
            lds_linear<T, KN>(lb, mb);
            if (!ta && !tb) tiny_product<T, M, N, K, false, false>(la, lb, lc);
            else if (ta && !tb) tiny_product<T, M, N, K, true, false>(la, lb, lc);
            else if (!ta && tb) tiny_product<T, M, N, K, false, true>(la, lb, lc);
            else tiny_product<T, M, N, K, true, true>(la, lb, lc);
            sts_linear<T, MN>(cst + (size_t)threadIdx.x * MN, lc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 2, %0;" ::"n"(CONSUMERS) : "memory");
        if (threadIdx.x == 0) {
            const size_t cbytes = (size_t)cnt * MN * sizeof(T);
            const uint32_t whole = (uint32_t)(cbytes & ~(size_t)15);
            if (whole)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(c + first * MN), "r"(smem_u32(cst)),
                             "r"(whole)
                             : "memory");
            // (a ragged last chunk may end inside a 16-byte word: those < 16 bytes go out with plain stores)
            for (size_t e = whole / sizeof(T); e < (size_t)cnt * MN; ++e) c[first * MN + e] = cst[e];
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename T, int M, int N, int K>
static int launch_tiny(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, int ta, int tb) {
    constexpr int G = 256, MK = M * K, KN = K * N, MN = M * N;
    const size_t stage_elems = ((size_t)G * (MK + KN) * sizeof(T) + 127) / 128 * 128 / sizeof(T);
    const size_t cstage_elems = ((size_t)G * MN * sizeof(T) + 127) / 128 * 128 / sizeof(T);
    const size_t smem = (3 * stage_elems + 2 * cstage_elems) * sizeof(T) + 6 * sizeof(uint64_t);
    const size_t nchunks = (batch + G - 1) / G;
    int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    size_t grid = (size_t)ctx->sm_count * per_sm;
    if (grid > nchunks) grid = nchunks;
    auto kern = gemm_tiny_kernel<T, M, N, K>;
    SLAS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, 288, smem, st>>>(a, b, c, batch, nchunks, ta, tb, (unsigned)stage_elems, (unsigned)cstage_elems);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

// (m, n, k) in 1..4 -> one of the 64 instantiations per element type
template <typename T, int M, int N>
static int launch_tiny_k(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t k, int ta, int tb) {
    switch (k) {
    case 1: return launch_tiny<T, M, N, 1>(ctx, st, batch, a, b, c, ta, tb);
    case 2: return launch_tiny<T, M, N, 2>(ctx, st, batch, a, b, c, ta, tb);
    case 3: return launch_tiny<T, M, N, 3>(ctx, st, batch, a, b, c, ta, tb);
    case 4: return launch_tiny<T, M, N, 4>(ctx, st, batch, a, b, c, ta, tb);
    }
    return SLAS_ERR_UNSUPPORTED;
}
template <typename T, int M>
static int launch_tiny_n(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t n, size_t k, int ta,
                         int tb) {
    switch (n) {
    case 1: return launch_tiny_k<T, M, 1>(ctx, st, batch, a, b, c, k, ta, tb);
    case 2: return launch_tiny_k<T, M, 2>(ctx, st, batch, a, b, c, k, ta, tb);
    case 3: return launch_tiny_k<T, M, 3>(ctx, st, batch, a, b, c, k, ta, tb);
    case 4: return launch_tiny_k<T, M, 4>(ctx, st, batch, a, b, c, k, ta, tb);
    }
    return SLAS_ERR_UNSUPPORTED;
}

template <typename T>
int launch_gemm_tiny_batched(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m, size_t n,
                             size_t k, int ta, int tb) {
    if (m < 1 || m > 4 || n < 1 || n > 4 || k < 1 || k > 4 || batch < 1024 || tunable("gemm_tiny_variant") == 1)
        return SLAS_ERR_UNSUPPORTED;
    // the final (ragged) chunk is copied in whole 16-byte words: the input arrays themselves must end on one
    if ((batch * m * k * sizeof(T)) % 16 || (batch * k * n * sizeof(T)) % 16) return SLAS_ERR_UNSUPPORTED;
    switch (m) {
    case 1: return launch_tiny_n<T, 1>(ctx, st, batch, a, b, c, n, k, ta, tb);
    case 2: return launch_tiny_n<T, 2>(ctx, st, batch, a, b, c, n, k, ta, tb);
    case 3: return launch_tiny_n<T, 3>(ctx, st, batch, a, b, c, n, k, ta, tb);
    case 4: return launch_tiny_n<T, 4>(ctx, st, batch, a, b, c, n, k, ta, tb);
    }
    return SLAS_ERR_UNSUPPORTED;
}
// built twice (Makefile): -DSLAS_TINY_F64 selects the double instantiations, so the two halves compile in parallel
#ifndef SLAS_TINY_F64
template int launch_gemm_tiny_batched<float>(Context *, cudaStream_t, size_t, const float *, const float *, float *, size_t,
                                             size_t, size_t, int, int);
#else
template int launch_gemm_tiny_batched<double>(Context *, cudaStream_t, size_t, const double *, const double *, double *, size_t,
                                              size_t, size_t, int, int);
#endif

}  // namespace slas
