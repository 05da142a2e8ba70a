/*
 * slas_oracle.h -- CPU restatement of the slas hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under slas_b200/ may include, link or call this.  Only tests/, the
 * smoke() check in __graft_entry__.py and the cpu_baseline / --impl reference
 * legs of bench.py use it, and only as the checker / the timed CPU baseline.
 *
 * Every function restates one reference function; the reference file:line it
 * follows is cited next to each prototype (paths relative to /root/reference).
 * The reference itself (nightly-only Rust, no rustc in this image) cannot be
 * compiled here, so this restatement is pinned against the reference's own
 * golden vectors (tests/golden/reference_vectors.json, transcribed from
 * tests/src/main.rs and the doc-tests).  Real-valued snrm2/dnrm2, dgemm, dgemv,
 * ddot, zdotu, dznrm2 have no golden vector in the reference: for those the
 * parity is UNPINNED (convention pinned by the f32 twins, rounding is not).
 */
#ifndef SLAS_ORACLE_H
#define SLAS_ORACLE_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/backends/rust.rs:20-41 -- LANES-wide accumulators, mul then add (no FMA),
 * ordered horizontal sum (lane 0..LANES-1), scalar tail.  lanes = the value of
 * simd_lanes::max_for_type (src/simd_lanes.rs:24-37): 4/8 for f32, 2/4 for f64. */
float  oracle_sdot(size_t len, const float *a, const float *b, int lanes);
double oracle_ddot(size_t len, const double *a, const double *b, int lanes);

/* src/backends/blas.rs:31-52 -- cblas_{c,z}dotu_sub: UNconjugated complex dot,
 * interleaved (re,im) pairs, result through out[2]. */
void oracle_cdotu(size_t len, const float *a, const float *b, float *out);
void oracle_zdotu(size_t len, const double *a, const double *b, double *out);

/* src/backends/rust.rs:45-73 -- c[i] = a[i] op b[i]; one IEEE op per element.
 * op: 0 add, 1 sub, 2 mul, 3 div. */
void oracle_sewise(int op, size_t len, const float *a, const float *b, float *c);
void oracle_dewise(int op, size_t len, const double *a, const double *b, double *c);

/* src/backends/rust.rs:116-127 -- sqrt of the strictly sequential sum of a[i]*a[i];
 * normalize divides every element by that norm (true division). */
float  oracle_snorm(size_t len, const float *a);
double oracle_dnorm(size_t len, const double *a);
void   oracle_snormalize(size_t len, float *a);
void   oracle_dnormalize(size_t len, double *a);

/* src/backends/rust.rs:129-147 -- complex norm: two running sums (re^2, im^2),
 * added at the end, sqrt; normalize divides re and im by the norm. */
float  oracle_cnorm(size_t len, const float *a);
double oracle_znorm(size_t len, const double *a);
void   oracle_cnormalize(size_t len, float *a);
void   oracle_znormalize(size_t len, double *a);

/* src/backends/rust.rs:151-177 -- buffer[columns*row + column] = a[(len/columns)*column + row].
 * elem = element size in bytes (any Copy type; bit-exact data movement). */
void oracle_transpose(size_t len, size_t elem, const void *a, void *buffer, size_t columns);
void oracle_transpose_inplace(size_t len, size_t elem, void *a, size_t columns);

/* src/backends/blas.rs:94-109 -- cblas_{s,d}gemm(RowMajor, opA, opB, m, n, k, 1, A, lda, B, ldb, 0, C, ldc).
 * *_f64acc accumulate each output in double (the "truth" variant used for tolerances);
 * the plain variants accumulate sequentially over k in T with separate mul and add. */
void oracle_sgemm(const float *a, const float *b, float *c, size_t m, size_t n, size_t k,
                  size_t lda, size_t ldb, size_t ldc, int a_trans, int b_trans);
void oracle_sgemm_f64acc(const float *a, const float *b, float *c, size_t m, size_t n, size_t k,
                         size_t lda, size_t ldb, size_t ldc, int a_trans, int b_trans);
void oracle_dgemm(const double *a, const double *b, double *c, size_t m, size_t n, size_t k,
                  size_t lda, size_t ldb, size_t ldc, int a_trans, int b_trans);

/* src/backends/blas.rs:114-151 -- matrix_vector_mul with the TRAIT's argument
 * meaning: m = underlying width (columns), n = underlying height (rows);
 * Blas passes cblas_gemv(RowMajor, op, M=n, N=m, 1, A, lda, x, 1, 0, y, 1). */
void oracle_sgemv(const float *a, const float *x, float *y, size_t m, size_t n, size_t lda, int a_trans);
void oracle_dgemv(const double *a, const double *x, double *y, size_t m, size_t n, size_t lda, int a_trans);

/* src/tensor.rs:202-219 -- linear index, axis 0 fastest; returns (size_t)-1 when
 * any index is out of bounds (the reference asserts). */
size_t oracle_tensor_index(size_t ndim, const size_t *shape, const size_t *index);
/* src/tensor.rs:296-301 -- Matrix[(r,c)] on a shape [axis0, axis1] with the lazy-transpose flag. */
size_t oracle_matrix_index(const size_t shape[2], int is_trans, size_t r, size_t c);

/* src/tensor.rs:376-391 -- derive (m,n,k,lda,ldb,ldc) from two matrix views.
 * out = {m, n, k, lda, ldb, ldc}. */
void oracle_matmul_args(const size_t shape_a[2], int a_trans, const size_t shape_b[2], int b_trans,
                        size_t out[6]);

/* Batched CPU baselines: what a slas user writes today -- a loop of single calls
 * over [[T; LEN]; batch] contiguous objects. */
void oracle_sgemm_batched(size_t batch, const float *a, const float *b, float *c, size_t m, size_t n,
                          size_t k, int a_trans, int b_trans);
void oracle_dgemm_batched(size_t batch, const double *a, const double *b, double *c, size_t m, size_t n,
                          size_t k, int a_trans, int b_trans);

#ifdef __cplusplus
}
#endif
#endif
