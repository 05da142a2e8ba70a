// kernels.cuh -- internal launchers.  Every launcher enqueues on the stream it is given and
// works on DEVICE pointers; api.cu adds the host-pointer staging and the C ABI on top.
#pragma once
#include "common.cuh"
#include "p2p.cuh"

namespace slas {

// vector_ops.cu
template <typename T>
int launch_ewise(Context *ctx, cudaStream_t st, int op, size_t len, const T *a, const T *b, T *c);
// kind 0 dot, 1 squared-norm, 2 complex dotu (len counts REAL scalars); mode 0 cast, 1 sqrt
template <typename T>
int launch_reduce(Context *ctx, cudaStream_t st, int kind, size_t len, const T *a, const T *b, double *result,
                  T *typed_out, int mode, int accumulate, const PeerExchange *peers = nullptr);
template <typename T> int launch_finish_scalar(cudaStream_t st, const double *in, T *out, int mode);
// normalize of a short vector in one launch (registers hold the vector between the two phases); *handled = false: not
// applicable (too long for one wave's registers, misaligned), use launch_reduce + launch_scale_div
template <typename T> int launch_normalize_fused(Context *ctx, cudaStream_t st, size_t len, T *a, T *norm_out, bool *handled);
template <typename T> int launch_scale_div(Context *ctx, cudaStream_t st, size_t len, const T *in, T *a, const T *norm_dev);
// fused chains: 0..3 = (add, sub, mul, div) -> dot with d (c optional), 4 = dot(a,b), |a|, |b|; results to typed_out (device)
template <typename T>
int launch_fused_chain(Context *ctx, cudaStream_t st, int chain, size_t len, const T *a, const T *b, T *c, const T *d,
                       T *typed_out);
// mode 0 dot, 1 norm, 2 normalize in place
template <typename T>
int launch_batched_reduce(Context *ctx, cudaStream_t st, int mode, size_t batch, size_t len, const T *a,
                          const T *b, T *out, T *a_mut);
// batched_cluster.cu -- norm (MODE 1) / normalize (MODE 2) of medium vectors, a thread-block cluster per vector;
// *handled = false: not a length for it
template <typename T, int MODE>
int launch_batched_norm_cluster(Context *ctx, cudaStream_t st, size_t batch, size_t len, const T *a, T *out, T *a_mut, bool *handled);
template <typename T>
int launch_fill_uniform(Context *ctx, cudaStream_t st, size_t len, T *x, uint64_t seed, uint64_t offset, T lo, T hi);

// transpose.cu -- batch objects of `len` elements each
int launch_transpose(Context *ctx, cudaStream_t st, size_t batch, size_t len, size_t elem_size, const void *a,
                     void *buffer, size_t columns);
// in place; *handled == false: no in-place kernel for this shape, use a temporary
int launch_transpose_inplace(Context *ctx, cudaStream_t st, size_t batch, size_t len, size_t elem_size, void *a,
                             size_t columns, bool *handled);

// gemm_jit.cu -- the gemm_small kernel specialised at run time (NVRTC) for shapes without a built-in instantiation;
// SLAS_ERR_UNSUPPORTED when the shape does not fit its tiling rules or NVRTC is not available
template <typename T>
int launch_gemm_jit_batched(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m, size_t n,
                            size_t k, int ta, int tb);

// gemm_tiny.cu -- every dimension in 1..4: one thread per matrix; SLAS_ERR_UNSUPPORTED otherwise
template <typename T>
int launch_gemm_tiny_batched(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m, size_t n,
                             size_t k, int ta, int tb);

// gemm_small.cu -- batched dense small GEMM; returns SLAS_ERR_UNSUPPORTED when no
// shape-specialised kernel exists (caller falls back to the generic kernel).
template <typename T>
int launch_gemm_small_batched(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c,
                              size_t m, size_t n, size_t k, int ta, int tb);

// gemm_generic.cu -- any shape / leading dimensions, optional batch with element strides
template <typename T>
int launch_gemm_generic(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m,
                        size_t n, size_t k, size_t lda, size_t ldb, size_t ldc, int ta, int tb, size_t stride_a,
                        size_t stride_b, size_t stride_c);
template <typename T>
int launch_gemv(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *x, T *y, size_t m, size_t n,
                size_t lda, int ta, size_t stride_a, size_t stride_x, size_t stride_y);

// gemm_large.cu -- FFMA/DFMA tiled kernel for big single products
template <typename T>
int launch_gemm_large(Context *ctx, cudaStream_t st, const T *a, const T *b, T *c, size_t m, size_t n, size_t k,
                      size_t lda, size_t ldb, size_t ldc, int ta, int tb);

// gemm_tcgen05.cu -- 3xTF32 on the 5th-gen tensor cores (TMA + TMEM); SLAS_ERR_UNSUPPORTED
// when the shape does not fit its tiles.
// gemm_large.cu -- a batch of medium-sized products on the register-tiled FFMA / DFMA kernel, one grid layer per object
template <typename T>
int launch_gemm_large_batched(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m, size_t n,
                              size_t k, size_t lda, size_t ldb, size_t ldc, int ta, int tb, size_t stride_a, size_t stride_b,
                              size_t stride_c);
// gemm_large_f64.cu -- 128 x 128 tiles, 8 x 8 per thread
int launch_dgemm_big(Context *ctx, cudaStream_t st, const double *a, const double *b, double *c, size_t m, size_t n, size_t k,
                     size_t lda, size_t ldb, size_t ldc, int ta, int tb);
// the zero-padded problem size the tcgen05 kernels would run (m, n, k) on, and whether the CTA-pair kernel is used
void sgemm_3xtf32_padded(size_t m, size_t n, size_t k, size_t *mp, size_t *np, size_t *kp, bool *pair);
int launch_sgemm_3xtf32(Context *ctx, cudaStream_t st, const float *a, const float *b, float *c, size_t m,
                        size_t n, size_t k, size_t lda, size_t ldb, size_t ldc, int ta, int tb);

}  // namespace slas
