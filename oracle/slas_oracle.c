/*
 * slas_oracle.c -- CPU restatement of the slas hot path.  TEST INFRASTRUCTURE ONLY
 * (see slas_oracle.h).  Build with -ffp-contract=off -fno-fast-math: the Rust
 * reference never contracts a*b+c into an FMA and never reassociates.
 */
#include "slas_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAX_LANES 16

/* ---- dot: src/backends/rust.rs:20-41 ------------------------------------ */
#define DEFINE_DOT(NAME, T)                                                              \
    T NAME(size_t len, const T *a, const T *b, int lanes) {                              \
        T acc[MAX_LANES];                                                                \
        if (lanes < 1) lanes = 1;                                                        \
        if (lanes > MAX_LANES) lanes = MAX_LANES;                                        \
        for (int l = 0; l < lanes; ++l) acc[l] = (T)0;                                   \
        size_t body = len / (size_t)lanes;                                               \
        for (size_t n = 0; n < body; ++n)         /* rust.rs:29-34 */                    \
            for (int l = 0; l < lanes; ++l) {                                            \
                T p = a[n * lanes + l] * b[n * lanes + l];                               \
                acc[l] = acc[l] + p;                                                     \
            }                                                                            \
        T sum = (T)0;                             /* rust.rs:35 reduce_sum, ordered */   \
        for (int l = 0; l < lanes; ++l) sum = sum + acc[l];                              \
        for (size_t n = len - (len % (size_t)lanes); n < len; ++n) { /* rust.rs:36-38 */ \
            T p = a[n] * b[n];                                                           \
            sum = sum + p;                                                               \
        }                                                                                \
        return sum;                                                                      \
    }
DEFINE_DOT(oracle_sdot, float)
DEFINE_DOT(oracle_ddot, double)

/* ---- complex unconjugated dot: src/backends/blas.rs:31-52 ---------------- */
#define DEFINE_DOTU(NAME, T)                                                 \
    void NAME(size_t len, const T *a, const T *b, T *out) {                  \
        T re = (T)0, im = (T)0;                                              \
        for (size_t i = 0; i < len; ++i) {                                   \
            T ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1]; \
            T t0 = ar * br, t1 = ai * bi, t2 = ar * bi, t3 = ai * br;        \
            re = re + (t0 - t1);                                             \
            im = im + (t2 + t3);                                             \
        }                                                                    \
        out[0] = re;                                                         \
        out[1] = im;                                                         \
    }
DEFINE_DOTU(oracle_cdotu, float)
DEFINE_DOTU(oracle_zdotu, double)

/* ---- element-wise: src/backends/rust.rs:45-73 ---------------------------- */
#define DEFINE_EWISE(NAME, T)                                                 \
    void NAME(int op, size_t len, const T *a, const T *b, T *c) {             \
        switch (op) {                                                         \
        case 0: for (size_t i = 0; i < len; ++i) c[i] = a[i] + b[i]; break;   \
        case 1: for (size_t i = 0; i < len; ++i) c[i] = a[i] - b[i]; break;   \
        case 2: for (size_t i = 0; i < len; ++i) c[i] = a[i] * b[i]; break;   \
        default: for (size_t i = 0; i < len; ++i) c[i] = a[i] / b[i]; break;  \
        }                                                                     \
    }
DEFINE_EWISE(oracle_sewise, float)
DEFINE_EWISE(oracle_dewise, double)

/* ---- norm / normalize (real): src/backends/rust.rs:116-127 --------------- */
float oracle_snorm(size_t len, const float *a) {
    float s = 0.0f;
    for (size_t i = 0; i < len; ++i) { float p = a[i] * a[i]; s = s + p; }
    return sqrtf(s);
}
double oracle_dnorm(size_t len, const double *a) {
    double s = 0.0;
    for (size_t i = 0; i < len; ++i) { double p = a[i] * a[i]; s = s + p; }
    return sqrt(s);
}
void oracle_snormalize(size_t len, float *a) {
    float nrm = oracle_snorm(len, a);
    for (size_t i = 0; i < len; ++i) a[i] = a[i] / nrm;
}
void oracle_dnormalize(size_t len, double *a) {
    double nrm = oracle_dnorm(len, a);
    for (size_t i = 0; i < len; ++i) a[i] = a[i] / nrm;
}

/* ---- norm / normalize (complex): src/backends/rust.rs:129-147 ------------ */
float oracle_cnorm(size_t len, const float *a) {
    float re = 0.0f, im = 0.0f;
    for (size_t i = 0; i < len; ++i) {
        float r2 = a[2 * i] * a[2 * i], i2 = a[2 * i + 1] * a[2 * i + 1];
        re = re + r2; im = im + i2;
    }
    return sqrtf(re + im);
}
double oracle_znorm(size_t len, const double *a) {
    double re = 0.0, im = 0.0;
    for (size_t i = 0; i < len; ++i) {
        double r2 = a[2 * i] * a[2 * i], i2 = a[2 * i + 1] * a[2 * i + 1];
        re = re + r2; im = im + i2;
    }
    return sqrt(re + im);
}
void oracle_cnormalize(size_t len, float *a) {
    float nrm = oracle_cnorm(len, a);
    for (size_t i = 0; i < 2 * len; ++i) a[i] = a[i] / nrm;
}
void oracle_znormalize(size_t len, double *a) {
    double nrm = oracle_znorm(len, a);
    for (size_t i = 0; i < 2 * len; ++i) a[i] = a[i] / nrm;
}

/* ---- transpose: src/backends/rust.rs:151-177 ----------------------------- */
void oracle_transpose(size_t len, size_t elem, const void *a, void *buffer, size_t columns) {
    const unsigned char *src = (const unsigned char *)a;
    unsigned char *dst = (unsigned char *)buffer;
    if (columns == 0) return;
    size_t rows = len / columns;
    for (size_t column = 0; column < columns; ++column)       /* rust.rs:168 */
        for (size_t row = 0; row < rows; ++row)               /* rust.rs:169 */
            memcpy(dst + (columns * row + column) * elem, src + (rows * column + row) * elem, elem);
}
void oracle_transpose_inplace(size_t len, size_t elem, void *a, size_t columns) {
    void *tmp = calloc(len ? len : 1, elem);                  /* rust.rs:157 zeroed temp */
    oracle_transpose(len, elem, a, tmp, columns);
    memcpy(a, tmp, len * elem);                               /* rust.rs:159 */
    free(tmp);
}

/* ---- gemm: src/backends/blas.rs:94-109 ----------------------------------- */
#define OPA(i, p) (a_trans ? a[(p) * lda + (i)] : a[(i) * lda + (p)])
#define OPB(p, j) (b_trans ? b[(j) * ldb + (p)] : b[(p) * ldb + (j)])
void oracle_sgemm(const float *a, const float *b, float *c, size_t m, size_t n, size_t k, size_t lda,
                  size_t ldb, size_t ldc, int a_trans, int b_trans) {
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) {
            float s = 0.0f;
            for (size_t p = 0; p < k; ++p) { float t = OPA(i, p) * OPB(p, j); s = s + t; }
            c[i * ldc + j] = s;
        }
}
void oracle_sgemm_f64acc(const float *a, const float *b, float *c, size_t m, size_t n, size_t k,
                         size_t lda, size_t ldb, size_t ldc, int a_trans, int b_trans) {
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) {
            double s = 0.0;
            for (size_t p = 0; p < k; ++p) s += (double)OPA(i, p) * (double)OPB(p, j);
            c[i * ldc + j] = (float)s;
        }
}
void oracle_dgemm(const double *a, const double *b, double *c, size_t m, size_t n, size_t k, size_t lda,
                  size_t ldb, size_t ldc, int a_trans, int b_trans) {
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) {
            double s = 0.0;
            for (size_t p = 0; p < k; ++p) { double t = OPA(i, p) * OPB(p, j); s = s + t; }
            c[i * ldc + j] = s;
        }
}

/* ---- gemv: src/backends/blas.rs:114-151 (m = width, n = height) ---------- */
#define DEFINE_GEMV(NAME, T)                                                                     \
    void NAME(const T *a, const T *x, T *y, size_t m, size_t n, size_t lda, int a_trans) {       \
        if (!a_trans) {                                                                          \
            for (size_t i = 0; i < n; ++i) {                                                     \
                T s = (T)0;                                                                      \
                for (size_t j = 0; j < m; ++j) { T t = a[i * lda + j] * x[j]; s = s + t; }       \
                y[i] = s;                                                                        \
            }                                                                                    \
        } else {                                                                                 \
            for (size_t j = 0; j < m; ++j) {                                                     \
                T s = (T)0;                                                                      \
                for (size_t i = 0; i < n; ++i) { T t = a[i * lda + j] * x[i]; s = s + t; }       \
                y[j] = s;                                                                        \
            }                                                                                    \
        }                                                                                        \
    }
DEFINE_GEMV(oracle_sgemv, float)
DEFINE_GEMV(oracle_dgemv, double)

/* ---- index math: src/tensor.rs:202-219, 296-301, 376-391, 485-510 -------- */
size_t oracle_tensor_index(size_t ndim, const size_t *shape, const size_t *index) {
    size_t sum = 0, product = 1;
    for (size_t n = 0; n < ndim; ++n) {
        if (!(index[n] < shape[n])) return (size_t)-1;        /* tensor.rs:209-214 assert */
        sum += index[n] * product;
        product *= shape[n];
    }
    return sum;
}
size_t oracle_matrix_index(const size_t shape[2], int is_trans, size_t r, size_t c) {
    size_t idx[2];
    if (is_trans) { idx[0] = r; idx[1] = c; } else { idx[0] = c; idx[1] = r; }  /* tensor.rs:298 */
    return oracle_tensor_index(2, shape, idx);
}
void oracle_matmul_args(const size_t sa[2], int a_trans, const size_t sb[2], int b_trans, size_t out[6]) {
    size_t m = a_trans ? sa[0] : sa[1];                       /* self.rows(),   tensor.rs:495-501 */
    size_t k = b_trans ? sb[0] : sb[1];                       /* other.rows()                     */
    size_t n = b_trans ? sb[1] : sb[0];                       /* other.columns(), tensor.rs:504-510 */
    out[0] = m; out[1] = n; out[2] = k;
    out[3] = sa[0];                                           /* lda = axis_len(0), tensor.rs:380 */
    out[4] = sb[0];                                           /* ldb                              */
    out[5] = n;                                               /* ldc                              */
}

/* ---- batched loops (CPU baseline shape: one call per object) ------------- */
void oracle_sgemm_batched(size_t batch, const float *a, const float *b, float *c, size_t m, size_t n,
                          size_t k, int a_trans, int b_trans) {
    size_t lda = a_trans ? m : k, ldb = b_trans ? k : n;
    for (size_t i = 0; i < batch; ++i)
        oracle_sgemm(a + i * m * k, b + i * k * n, c + i * m * n, m, n, k, lda, ldb, n, a_trans, b_trans);
}
void oracle_dgemm_batched(size_t batch, const double *a, const double *b, double *c, size_t m, size_t n,
                          size_t k, int a_trans, int b_trans) {
    size_t lda = a_trans ? m : k, ldb = b_trans ? k : n;
    for (size_t i = 0; i < batch; ++i)
        oracle_dgemm(a + i * m * k, b + i * k * n, c + i * m * n, m, n, k, lda, ldb, n, a_trans, b_trans);
}
