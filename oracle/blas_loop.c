/*
 * blas_loop.c -- CPU-baseline driver for the reference's `Blas` backend.  TEST /
 * BENCH INFRASTRUCTURE ONLY (see slas_oracle.h).
 *
 * The reference's Blas backend is a thin FFI to cblas (src/backends/blas.rs:21,
 * 94-109, 136-149, 161); a slas user who wants a batch of small products writes
 * a loop of Matrix::matrix_mul calls (src/tensor.rs:441-454), i.e. a loop of
 * cblas_sgemm(RowMajor, opA, opB, m, n, k, 1, A, lda, B, ldb, 0, C, ldc).  This
 * file is that loop, in C so that the loop overhead is not Python's, with the
 * cblas entry points passed in as function pointers (they come from scipy's
 * bundled OpenBLAS, the stand-in for the un-vendored system BLAS).
 * `threads` > 1 splits the batch over pthreads: the reference itself is single
 * threaded, this is the most generous reading of "all the host cores".
 */
#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>

typedef void (*sgemm_fn)(int, int, int, int, int, int, float, const float *, int, const float *, int,
                         float, float *, int);
typedef void (*dgemm_fn)(int, int, int, int, int, int, double, const double *, int, const double *, int,
                         double, double *, int);
typedef float (*sdot_fn)(int, const float *, int, const float *, int);

enum { RowMajor = 101, NoTrans = 111, Trans = 112 };

typedef struct {
    void *fn; int is_double;
    size_t lo, hi, m, n, k; int ta, tb;
    const void *a, *b; void *c;
} job_t;

static void *run(void *p) {
    job_t *j = (job_t *)p;
    size_t m = j->m, n = j->n, k = j->k;
    int lda = (int)(j->ta ? m : k), ldb = (int)(j->tb ? k : n), ldc = (int)n;
    int opa = j->ta ? Trans : NoTrans, opb = j->tb ? Trans : NoTrans;
    if (j->is_double) {
        const double *a = (const double *)j->a, *b = (const double *)j->b; double *c = (double *)j->c;
        dgemm_fn f = (dgemm_fn)j->fn;
        for (size_t i = j->lo; i < j->hi; ++i)
            f(RowMajor, opa, opb, (int)m, (int)n, (int)k, 1.0, a + i * m * k, lda, b + i * k * n, ldb, 0.0,
              c + i * m * n, ldc);
    } else {
        const float *a = (const float *)j->a, *b = (const float *)j->b; float *c = (float *)j->c;
        sgemm_fn f = (sgemm_fn)j->fn;
        for (size_t i = j->lo; i < j->hi; ++i)
            f(RowMajor, opa, opb, (int)m, (int)n, (int)k, 1.0f, a + i * m * k, lda, b + i * k * n, ldb, 0.0f,
              c + i * m * n, ldc);
    }
    return NULL;
}

/* batch x (m,n,k) products over contiguous [[T; LEN]; batch] operands. */
int blas_gemm_loop(void *gemm, int is_double, size_t batch, const void *a, const void *b, void *c,
                   size_t m, size_t n, size_t k, int ta, int tb, int threads) {
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)threads);
    if (!th || !jobs) return 1;
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (job_t){gemm, is_double, batch * (size_t)t / (size_t)threads,
                          batch * (size_t)(t + 1) / (size_t)threads, m, n, k, ta, tb, a, b, c};
        if (t + 1 < threads) pthread_create(&th[t], NULL, run, &jobs[t]);
    }
    run(&jobs[threads - 1]);
    for (int t = 0; t + 1 < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}

/* dot over len >= 2^31: the reference's `LEN as i32` (blas.rs:21) would truncate, so chunk. */
double blas_sdot_chunked(void *sdot, size_t len, const float *a, const float *b) {
    sdot_fn f = (sdot_fn)sdot;
    const size_t chunk = (size_t)1 << 30;
    double acc = 0.0;
    for (size_t o = 0; o < len; o += chunk) {
        size_t c = len - o < chunk ? len - o : chunk;
        acc += (double)f((int)c, a + o, 1, b + o, 1);
    }
    return acc;
}
