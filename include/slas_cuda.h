/*
 * slas_cuda.h -- C ABI of libslas_cuda.so, the B200 (sm_100a) compute backend that
 * stands behind `slas_backend::Cuda`.
 *
 * This is the drop-in boundary: every op entry point takes exactly the arguments
 * the matching method of the reference's backend-operation traits takes
 * (/root/reference/src/backends.rs:121-183), with `impl StaticVec<T, LEN>`
 * lowered to (len, pointer) -- the only thing a StaticVec gives a backend is
 * `as_ptr()/as_mut_ptr()` to LEN contiguous elements (src/static_vec.rs:121-134).
 * Each declaration cites the reference function it replaces; INTEGRATION.md holds
 * the `extern "C"` block and trait impls a slas maintainer adds on the Rust side.
 *
 * Conventions
 *  - Prefix s = f32, d = f64, c = Complex<f32>, z = Complex<f64> (interleaved re,im).
 *  - Every function returns an int status (0 = SLAS_OK); the reference's ops are
 *    infallible by type, so the Rust shim panics with slas_cuda_last_error().
 *  - Sizes are size_t, never int: the reference's `LEN as i32` (src/backends/blas.rs:21)
 *    truncates at 2^31 elements, this ABI does not.
 *  - Data pointers may be DEVICE pointers (the fast, device-resident path) or HOST
 *    pointers (staged through pinned buffers, copies overlapped with the kernels);
 *    all data operands of one call must live on the same side.  Scalar results
 *    (`out`) may independently be host or device pointers.
 *  - Ops are enqueued on the context's stream (slas_cuda_set_stream); a call
 *    returns after enqueueing unless it has to hand a scalar or host data back.
 *  - There is no CPU fallback: without a CUDA device every op fails with
 *    SLAS_ERR_CUDA.
 */
#ifndef SLAS_CUDA_H
#define SLAS_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    SLAS_OK = 0,
    SLAS_ERR_INVALID = 1,     /* bad argument (null pointer, shape mismatch, mixed host/device operands) */
    SLAS_ERR_CUDA = 2,        /* CUDA runtime error, or no device */
    SLAS_ERR_NCCL = 3,        /* NCCL missing or a collective failed */
    SLAS_ERR_UNSUPPORTED = 4  /* valid in the reference but not built here */
};

/* ---- context -------------------------------------------------------------------
 * A reference backend is a zero-sized type made by `B::default()` at every call
 * site (src/backends/rust.rs:2-3, src/static_vec.rs:76,219), so it can hold no
 * state: device, stream and scratch live in a process-global, lazily initialised
 * context per device.  Calls are thread-safe (libtest runs tests concurrently). */
int slas_cuda_init(int device);              /* device < 0: keep the current device */
int slas_cuda_shutdown(void);
const char *slas_cuda_last_error(void);      /* thread-local message of the last failure */
int slas_cuda_device_count(int *count);
int slas_cuda_device_info(int *sm_count, size_t *l2_bytes, size_t *hbm_bytes, int *cc_major, int *cc_minor);
int slas_cuda_set_stream(void *cuda_stream); /* a cudaStream_t; NULL restores the library's own stream */
int slas_cuda_synchronize(void);
/* Number of kernels this library launched on the calling process so far (bench.py's gpu_launches). */
uint64_t slas_cuda_launch_count(void);

/* CUDA graphs for launch-bound sequences (BASELINE config 1: dot / add at LEN 2^20 are a few microseconds of HBM
 * time each).  Between graph_begin and graph_end every op called with DEVICE pointers (results included: a host
 * result pointer needs a synchronising copy and fails the capture) is recorded on the backend's stream instead
 * of executed; graph_launch replays the whole sequence with one launch.  Scratch buffers that grow during the
 * capture are allocated then and kept (relaxed capture mode).  Calls from other threads between begin and end are
 * recorded too; replays are stream-ordered with every other op. */
int slas_cuda_graph_begin(void);
int slas_cuda_graph_end(void **graph_exec);
int slas_cuda_graph_launch(void *graph_exec);
int slas_cuda_graph_destroy(void *graph_exec);

/* Kernel-selection knobs for benchmarking and profiling (not part of the reference interface): a
 * named integer, e.g. "transpose_variant", "gemm32d_variant".  Unknown names are rejected. */
int slas_cuda_set_tunable(const char *name, long value);
int slas_cuda_get_tunable(const char *name, long *value);
/* Roofline denominator measured on the box: sustained FMA TFLOP/s of a register-only 8x8
 * outer-product loop (the GEMM kernels' instruction mix), f32 (is_double = 0) or f64. */
int slas_cuda_bench_fma_peak(int is_double, double *tflops);
/* Tensor-pipe denominator measured on the box: back-to-back tcgen05.mma.kind::tf32 (M128 N256 K8) from resident
 * shared-memory tiles, one CTA per SM, no memory traffic -> dense TF32 TFLOP/s at the clocks the chip sustains.
 * The 3xTF32 product executes three such MMAs per useful one. */
int slas_cuda_bench_tf32_peak(double *tflops);
/* BASELINE configs[0] (the reference's own dot_product bench, tests/src/main.rs:589-601, is a loop of single calls):
 * `calls` back-to-back calls of slas_cuda_sdot (op 0, results to out[i]), slas_cuda_sadd (op 1, results to
 * out + i * stride) or slas_cuda_sdot with its scalar returned to the HOST after every call (op 2: what `a.dot(&b)` does
 * in the reference) or slas_cuda_snormalize in place on out + i * stride (op 3), on operands `stride` elements apart,
 * issued from C -- the launch-bound loop without an interpreter between the calls.  a, b device pointers. */
int slas_cuda_bench_call_loop(int op, size_t len, const float *a, const float *b, float *out, int calls, size_t stride);
/* Self-check of the hoisted division normalize uses (csrc/divby.cuh: the reciprocal refinement of div.rn.f32 runs once per
 * divisor, each numerator costs three FMAs; operands outside [2^-60, 2^60] take the plain division): for each of the
 * `count` divisors (host array) EVERY one of the 2^32 float bit patterns is divided both ways on the device;
 * *mismatches = how many quotients differ in their bits from `x / n` (NaN == NaN).  0 is the only acceptable answer. */
int slas_cuda_bench_divby_mismatches(const float *divisors, size_t count, uint64_t *mismatches);
/* With the tunable "reduce_phases" = 1 the dot / norm kernel stamps %globaltimer (ns): [0] start, [1] block 0's
 * loads + FMAs done, [2] its block sum done, [3] grid combine done, [4] result written. */
int slas_cuda_get_reduce_phases(uint64_t *ns5);

/* ---- device-resident storage (what `CudaVec<T, LEN>` / `CudaBatch` own) --------- */
int slas_cuda_malloc(void **dptr, size_t bytes);
int slas_cuda_free(void *dptr);
int slas_cuda_malloc_host(void **hptr, size_t bytes);   /* pinned */
int slas_cuda_free_host(void *hptr);
/* Host operands.  What a slas StaticVec hands over is PAGEABLE memory ([T; LEN] / Vec, src/static_vec.rs:126): the
 * library stages it through its own pinned bounce ring with a pool of copy threads (streaming stores), overlapped with
 * the DMA engines -- three passes over host memory per byte instead of one, so it runs at ~0.73 of the pinned rate on
 * a host whose DRAM is the limit.  Memory from slas_cuda_malloc_host, or a caller's own long-lived allocation
 * registered in place with slas_cuda_host_register (unregister before freeing it), is read / written by the DMA
 * engines directly. */
int slas_cuda_host_register(void *hptr, size_t bytes);
int slas_cuda_host_unregister(void *hptr);
int slas_cuda_memcpy_h2d(void *dst, const void *src, size_t bytes);
int slas_cuda_memcpy_d2h(void *dst, const void *src, size_t bytes);
int slas_cuda_memcpy_d2d(void *dst, const void *src, size_t bytes);
int slas_cuda_memset(void *dst, int byte, size_t bytes);
/* 1 if p is a device pointer, 0 if host. */
int slas_cuda_is_device_ptr(const void *p, int *is_device);
/* Counter-based synthetic data straight into HBM (bench/test plumbing for vectors that
 * cannot be staged through the host): x[i] = lo + (hi-lo)*u(seed, offset+i). */
int slas_cuda_sfill_uniform(size_t len, float *x, uint64_t seed, uint64_t offset, float lo, float hi);
int slas_cuda_dfill_uniform(size_t len, double *x, uint64_t seed, uint64_t offset, double lo, double hi);

/* ---- DotProduct::dot  (src/backends.rs:122-126; Rust: src/backends/rust.rs:20-41;
 *      Blas: src/backends/blas.rs:15-23, complex dotu :31-52) ---------------------- */
int slas_cuda_sdot(size_t len, const float *a, const float *b, float *out);
int slas_cuda_ddot(size_t len, const double *a, const double *b, double *out);
int slas_cuda_cdotu(size_t len, const float *a, const float *b, float *out /* [re, im] */);
int slas_cuda_zdotu(size_t len, const double *a, const double *b, double *out /* [re, im] */);

/* ---- Normalize::{norm, normalize}  (src/backends.rs:128-130; Rust real: rust.rs:116-127,
 *      complex: rust.rs:129-147; Blas: blas.rs:156-170).  NormOutput is the real type. --- */
int slas_cuda_snrm2(size_t len, const float *a, float *out);
int slas_cuda_dnrm2(size_t len, const double *a, double *out);
int slas_cuda_scnrm2(size_t len, const float *a, float *out);
int slas_cuda_dznrm2(size_t len, const double *a, double *out);
int slas_cuda_snormalize(size_t len, float *a);    /* a[i] /= norm (true division, rust.rs:125) */
int slas_cuda_dnormalize(size_t len, double *a);
int slas_cuda_cnormalize(size_t len, float *a);    /* len complex elements */
int slas_cuda_znormalize(size_t len, double *a);

/* ---- Addition/Subtraction/Multiplication/Divition  (src/backends.rs:156-182;
 *      Rust: src/backends/rust.rs:45-73).  c[i] = a[i] op b[i], one IEEE op, bit-exact. --- */
int slas_cuda_sadd(size_t len, const float *a, const float *b, float *c);
int slas_cuda_ssub(size_t len, const float *a, const float *b, float *c);
int slas_cuda_smul(size_t len, const float *a, const float *b, float *c);
int slas_cuda_sdiv(size_t len, const float *a, const float *b, float *c);
int slas_cuda_dadd(size_t len, const double *a, const double *b, double *c);
int slas_cuda_dsub(size_t len, const double *a, const double *b, double *c);
int slas_cuda_dmul(size_t len, const double *a, const double *b, double *c);
int slas_cuda_ddiv(size_t len, const double *a, const double *b, double *c);

/* ---- Fused chains: op sequences a slas user issues back to back today, in one pass over
 *      HBM.  slas leaves lazy execution to its users (CONTRIBUTING.md:7-22), so these are extra entry points,
 *      not a change of the trait methods; each gives the same bits as the unfused chain on this backend.
 *      x_dot: c = a (+,-,*,/) b as rust.rs:45-73 (c may be NULL: the intermediate is never materialised), then
 *             *out = dot(c, d) as rust.rs:20-41 -- 12 (+4 with c) bytes per f32 element instead of 20.
 *      dot_norms: out3 = { dot(a,b), norm(a), norm(b) } (rust.rs:20-41, 116-121) in one read of a and b
 *             (cosine similarity = out3[0] / (out3[1] * out3[2])): 8 bytes per f32 element instead of 16.
 *      normalize_into: out = a / norm(a), a untouched -- the copy-on-write (static_vec.rs:304-318) +
 *             normalize (rust.rs:122-127) chain without the memcpy pass. ------------------------------------ */
int slas_cuda_sadd_dot(size_t len, const float *a, const float *b, float *c, const float *d, float *out);
int slas_cuda_ssub_dot(size_t len, const float *a, const float *b, float *c, const float *d, float *out);
int slas_cuda_smul_dot(size_t len, const float *a, const float *b, float *c, const float *d, float *out);
int slas_cuda_sdiv_dot(size_t len, const float *a, const float *b, float *c, const float *d, float *out);
int slas_cuda_dadd_dot(size_t len, const double *a, const double *b, double *c, const double *d, double *out);
int slas_cuda_dsub_dot(size_t len, const double *a, const double *b, double *c, const double *d, double *out);
int slas_cuda_dmul_dot(size_t len, const double *a, const double *b, double *c, const double *d, double *out);
int slas_cuda_ddiv_dot(size_t len, const double *a, const double *b, double *c, const double *d, double *out);
int slas_cuda_sdot_norms(size_t len, const float *a, const float *b, float *out3);
int slas_cuda_ddot_norms(size_t len, const double *a, const double *b, double *out3);
int slas_cuda_snormalize_into(size_t len, const float *a, float *out);
int slas_cuda_dnormalize_into(size_t len, const double *a, double *out);

/* ---- Transpose::{transpose, transpose_inplace}  (src/backends.rs:152-154;
 *      Rust: src/backends/rust.rs:151-177).  buffer[columns*row + column] =
 *      a[(len/columns)*column + row]; `columns` is the column count of the OUTPUT.
 *      Generic over any Copy element like the reference (`Transpose<T: Copy>`): any elem_size (1/2/4/8/16 bytes on
 *      the vector kernels, anything else -- a 12-byte [f32; 3] ... -- on a granule kernel).  Bit-exact. --- */
int slas_cuda_transpose(size_t len, size_t elem_size, const void *a, void *buffer, size_t columns);
int slas_cuda_transpose_inplace(size_t len, size_t elem_size, void *a, size_t columns);

/* ---- MatrixMul::matrix_mul  (src/backends.rs:132-141; Blas: src/backends/blas.rs:67-111)
 *      C[m x n] = op(A) . op(B), row-major, alpha = 1, beta = 0, exactly the
 *      cblas_{s,d}gemm(RowMajor, opA, opB, m, n, k, 1, A, lda, B, ldb, 0, C, ldc) call. ---- */
int slas_cuda_sgemm(const float *a, const float *b, float *c, size_t m, size_t n, size_t k,
                    size_t lda, size_t ldb, size_t ldc, int a_trans, int b_trans);
int slas_cuda_dgemm(const double *a, const double *b, double *c, size_t m, size_t n, size_t k,
                    size_t lda, size_t ldb, size_t ldc, int a_trans, int b_trans);
/* Which kernel a large f32 product runs on: 0 = FFMA only (every C element an ascending-k fp32 FMA chain),
 * 1 = tcgen05/TMEM 3xTF32 for every contraction its tiling covers (m % 128, n % 256, k % 32 == 0, dense),
 * 2 = auto, the default: 3xTF32 when the product also has >= 2^27 multiply-adds, FFMA otherwise.  3xTF32 is an
 * fp32-class result (measured <= 2e-6 * sum|a||b|, inside the 1e-5 * sqrt(K) parity bar).  Infinities and NaNs
 * propagate as in the fp32 chain (the split pass keeps an infinity out of the cross terms). */
int slas_cuda_set_sgemm_path(int path);
int slas_cuda_get_sgemm_path(int *path);

/* ---- MatrixMul::matrix_vector_mul  (src/backends.rs:142-150; Blas: blas.rs:114-151)
 *      TRAIT meaning of the arguments: m = underlying width (columns), n = underlying
 *      height (rows), lda = width; y = op(A) . x, i.e. cblas_gemv(RowMajor, op, M=n, N=m, ...). */
int slas_cuda_sgemv(const float *a, const float *x, float *y, size_t m, size_t n, size_t lda, int a_trans);
int slas_cuda_dgemv(const double *a, const double *x, double *y, size_t m, size_t n, size_t lda, int a_trans);

/* ---- batched views: `batch` same-shape objects, contiguous [[T; LEN]; batch] ------------
 * (north_star: "batched views over many same-shape objects, so compile-time shapes become
 * template parameters of the kernels").  Semantics = the single-object op applied to every
 * object; matrices are dense (lda = a_trans ? m : k, ldb = b_trans ? k : n, ldc = n as
 * Matrix::matrix_mul_buffer derives them, src/tensor.rs:376-382). */
int slas_cuda_sgemm_batched(size_t batch, const float *a, const float *b, float *c, size_t m, size_t n,
                            size_t k, int a_trans, int b_trans);
int slas_cuda_dgemm_batched(size_t batch, const double *a, const double *b, double *c, size_t m, size_t n,
                            size_t k, int a_trans, int b_trans);
int slas_cuda_sgemv_batched(size_t batch, const float *a, const float *x, float *y, size_t m, size_t n,
                            int a_trans);
int slas_cuda_dgemv_batched(size_t batch, const double *a, const double *x, double *y, size_t m, size_t n,
                            int a_trans);
int slas_cuda_transpose_batched(size_t batch, size_t len, size_t elem_size, const void *a, void *buffer,
                                size_t columns);
/* Matrix::deref_mut over a batch (src/tensor.rs:546-564 per object): every object transposed in place */
int slas_cuda_transpose_inplace_batched(size_t batch, size_t len, size_t elem_size, void *a, size_t columns);
int slas_cuda_sdot_batched(size_t batch, size_t len, const float *a, const float *b, float *out);
int slas_cuda_ddot_batched(size_t batch, size_t len, const double *a, const double *b, double *out);
int slas_cuda_snrm2_batched(size_t batch, size_t len, const float *a, float *out);
int slas_cuda_dnrm2_batched(size_t batch, size_t len, const double *a, double *out);
int slas_cuda_snormalize_batched(size_t batch, size_t len, float *a);
int slas_cuda_dnormalize_batched(size_t batch, size_t len, double *a);

/* ---- several GPUs inside ONE process ---------------------------------------------------------
 * A slas backend is a zero-sized type made by `B::default()` in an ordinary program (src/static_vec.rs:76,219,
 * src/backends/rust.rs:2-3): no launcher, no ranks.  slas_cuda_init_devices(n) (n <= 0: every visible device; the
 * environment variable SLAS_CUDA_DEVICES=n does the same at the first op) initialises devices 0..n-1 with peer
 * access between them.  From then on
 *  - an op on DEVICE operands runs on the device that owns them (the pointer says which);
 *  - element-wise ops and every *_batched entry on HOST operands of >= 64 MB are split over the devices by element
 *    range / batch index, each device running its own copy pipeline on its own worker thread;
 *  - dot / norm / normalize on HOST operands are split by vector slice; the f64 partials are added in device order;
 *  - the *_multi entries below take ONE DEVICE-RESIDENT SLICE PER DEVICE (slice i on its own GPU, any lengths) and
 *    do the exchange inside the reduction kernels over NVLink peer memory: the in-process form of the *_sharded ops
 *    further down (same kernels, same rank-ordered f64 sum, peer access instead of CUDA IPC). */
int slas_cuda_init_devices(int ndev);
int slas_cuda_devices(int *ndev, int *peer_exchange /* 1: the *_multi ops exchange over peer memory */);
int slas_cuda_malloc_on(int device, void **dptr, size_t bytes);
int slas_cuda_sdot_multi(int nslices, const size_t *lens, const float *const *a, const float *const *b, float *out);
int slas_cuda_ddot_multi(int nslices, const size_t *lens, const double *const *a, const double *const *b, double *out);
int slas_cuda_snrm2_multi(int nslices, const size_t *lens, const float *const *a, float *out);
int slas_cuda_dnrm2_multi(int nslices, const size_t *lens, const double *const *a, double *out);
int slas_cuda_snormalize_multi(int nslices, const size_t *lens, float *const *a);
int slas_cuda_dnormalize_multi(int nslices, const size_t *lens, double *const *a);

/* ---- multi-GPU: one process per GPU, this rank owns one contiguous slice -----------------
 * Batched ops shard by batch index and need no collective (call the *_batched entry on the
 * local slice).  A single huge dot / norm / normalize shards by vector slice: local
 * deterministic partial on this rank's slice, then ONE ncclAllReduce(sum) of the f64
 * partial on the context's stream. */
int slas_cuda_nccl_unique_id(void *id128);                      /* rank 0: 128-byte ncclUniqueId */
int slas_cuda_comm_init(int nranks, int rank, const void *id128);
int slas_cuda_comm_destroy(void);
int slas_cuda_comm_info(int *nranks, int *rank);                /* nranks = 1 when no communicator */
/* Peer-memory exchange: the fused form of the collective.  Every rank exports a small mailbox
 * (slas_cuda_p2p_handle: 64-byte CUDA IPC handle), the harness all-gathers the handles, and
 * slas_cuda_p2p_init maps the peers' mailboxes over NVLink.  From then on the *_sharded ops do their
 * exchange INSIDE the reduction kernel (peer stores + rank-ordered sum, bit-identical on every rank)
 * instead of calling NCCL.  nranks <= 8 (one NVSwitch domain). */
int slas_cuda_p2p_handle(void *handle64);
int slas_cuda_p2p_init(int nranks, int rank, const void *handles /* nranks x 64 bytes, rank order */);
int slas_cuda_p2p_destroy(void);
/* Slice [begin, end) of a len-element vector owned by `rank` of `nranks`; boundaries are
 * multiples of 16 bytes' worth of elements (elem_size <= 16). */
int slas_cuda_shard_range(size_t len, size_t elem_size, int nranks, int rank, size_t *begin, size_t *end);
int slas_cuda_sdot_sharded(size_t local_len, const float *a, const float *b, float *out);
int slas_cuda_ddot_sharded(size_t local_len, const double *a, const double *b, double *out);
int slas_cuda_snrm2_sharded(size_t local_len, const float *a, float *out);
int slas_cuda_dnrm2_sharded(size_t local_len, const double *a, double *out);
int slas_cuda_snormalize_sharded(size_t local_len, float *a);
int slas_cuda_dnormalize_sharded(size_t local_len, double *a);
/* Local f64 partial sums (sum a[i]*b[i]) without the collective: what the sharded ops reduce. */
int slas_cuda_sdot_partial(size_t len, const float *a, const float *b, double *partial);
int slas_cuda_ddot_partial(size_t len, const double *a, const double *b, double *partial);

#ifdef __cplusplus
}
#endif
#endif /* SLAS_CUDA_H */
