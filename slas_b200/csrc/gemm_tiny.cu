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
#include "common.cuh"
#include "gemm_tiny_kernel.cuh"
#include "kernels.cuh"

namespace slas {

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
