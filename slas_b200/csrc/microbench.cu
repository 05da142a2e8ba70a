// microbench.cu -- roofline denominators measured on the box itself ("measure an
// FFMA-peak microbench on the box and record it").  Not part of the reference interface.
//   slas_cuda_bench_fma_peak: register-only 8x8 outer-product FMA loop (the instruction mix of the
//   GEMM inner loops, no memory traffic) -> sustained FP32 / FP64 FMA TFLOP/s at the clocks the
//   chip actually holds under that load.  (The tensor-pipe counterpart, slas_cuda_bench_tf32_peak, lives in
//   gemm_tcgen05.cu next to the PTX it shares.)
#include "common.cuh"
#include "divby.cuh"

namespace slas {

template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *out, int iters, T seed) {
    T a[8], b[8], acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed + (T)(threadIdx.x + i) * (T)1e-3;
        b[i] = seed - (T)(threadIdx.x * 2 + i) * (T)1e-3;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = (T)0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            // keep the operands moving so nothing is hoisted or folded
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = a[i] * (T)0.999 + (T)1e-4; }
        }
    }
    T s = (T)0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    if (s == (T)123456789) out[0] = s;  // never true: defeats dead-code elimination
}

template <typename T>
static int bench_fma(Context *ctx, double *tflops, double *ms_out) {
    T *dummy = reinterpret_cast<T *>(ctx->scalars + 8);
    const int iters = sizeof(T) == 4 ? 4096 : 1024;
    const unsigned blocks = (unsigned)ctx->sm_count * 8;
    cudaEvent_t e0, e1;
    SLAS_CUDA_TRY(cudaEventCreate(&e0));
    SLAS_CUDA_TRY(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) fma_peak_kernel<T><<<blocks, 256, 0, ctx->stream>>>(dummy, iters, (T)1);
    SLAS_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) fma_peak_kernel<T><<<blocks, 256, 0, ctx->stream>>>(dummy, iters, (T)1);
    SLAS_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    SLAS_CUDA_TRY(cudaEventSynchronize(e1));
    count_launch(reps + 2);
    float ms = 0;
    SLAS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // per round: 64 FMAs of the outer product + the 8 FMAs that keep the operands moving (a*0.999 + 1e-4) = 72 issued
    const double flop = 2.0 * 72 * 4 * (double)iters * 256.0 * blocks * reps;
    *tflops = flop / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms / reps;
    return SLAS_OK;
}

}  // namespace slas

using namespace slas;

extern "C" int slas_cuda_bench_fma_peak(int is_double, double *tflops) {
    SLAS_REQUIRE(tflops, "null output");
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    return is_double ? bench_fma<double>(ctx, tflops, nullptr) : bench_fma<float>(ctx, tflops, nullptr);
}

// The launch-bound loop of BASELINE configs[0] driven from C (no interpreter between the calls): `calls` back-to-back
// calls of the public entry points on consecutive operands `stride` elements apart.  op 0: slas_cuda_sdot (results to
// out[i]); op 1: slas_cuda_sadd (results to out + i * stride); op 2: slas_cuda_sdot returning its scalar to the HOST
// after every call, the way `a.dot(&b)` does in the reference (out may be NULL; the last value is stored to out[0] when it
// is a host pointer); op 3: slas_cuda_snormalize in place on out + i * stride.  a, b device pointers.
extern "C" int slas_cuda_bench_call_loop(int op, size_t len, const float *a, const float *b, float *out, int calls,
                                         size_t stride) {
    SLAS_REQUIRE(op >= 0 && op <= 3, "call_loop: op must be 0 (dot), 1 (add), 2 (dot, scalar to the host) or 3 (normalize)");
    float host_result = 0.f;
    for (int i = 0; i < calls; ++i) {
        const size_t off = (size_t)i * stride;
        if (op == 0) SLAS_TRY(slas_cuda_sdot(len, a + off, b + off, out + i));
        else if (op == 1) SLAS_TRY(slas_cuda_sadd(len, a + off, b + off, out + off));
        else if (op == 3) SLAS_TRY(slas_cuda_snormalize(len, out + off));  // in place on out + i * stride
        else SLAS_TRY(slas_cuda_sdot(len, a + off, b + off, &host_result));
    }
    if (op == 2 && out) {
        bool dev = true;
        if (pointer_is_device(out, &dev) == SLAS_OK && !dev) out[0] = host_result;
    }
    return SLAS_OK;
}

// %globaltimer stamps of the last dot / norm kernel when the tunable "reduce_phases" is set: ns[0] kernel start
// (block 0), ns[1] block 0's loads and FMAs done, ns[2] its block sum done, ns[3] grid combine done (all partials
// in), ns[4] result written.  Synchronises the stream.
extern "C" int slas_cuda_get_reduce_phases(uint64_t *ns5) {
    SLAS_REQUIRE(ns5, "null output");
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    SLAS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    SLAS_CUDA_TRY(cudaMemcpy(ns5, ctx->phases, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return SLAS_OK;
}

// ---- exhaustive self-check of divby.cuh (see include/slas_cuda.h) ----------------------------------------------------
namespace slas {
__global__ void __launch_bounds__(256) divby_check_kernel(const float *divisors, unsigned long long *mismatches) {
    const float n = divisors[blockIdx.y];
    DivBy<float> by;
    by.init(n);
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < (1ull << 32); i += (unsigned long long)gridDim.x * 256) {
        const float x = __uint_as_float((unsigned)i);
        const float q = by.div(x), r = x / n;
        if (__float_as_uint(q) != __float_as_uint(r) && !(q != q && r != r)) ++bad;
        // the branch-free form the kernels use: fast() where ok(), else the plain division
        const float f = by.ok(x) ? by.fast(x) : r;
        if (__float_as_uint(f) != __float_as_uint(r) && !(f != f && r != r)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}
}  // namespace slas

extern "C" int slas_cuda_bench_divby_mismatches(const float *divisors, size_t count, uint64_t *mismatches) {
    using namespace slas;
    SLAS_REQUIRE(divisors && mismatches && count > 0 && count <= 65535, "divby self-check: 1 .. 65535 divisors");
    Context *ctx;
    SLAS_TRY(get_context(&ctx));
    float *d = nullptr;
    unsigned long long *m = nullptr;
    SLAS_CUDA_TRY(cudaMalloc(&d, count * sizeof(float)));
    cudaError_t e = cudaMalloc(&m, sizeof(unsigned long long));
    if (e != cudaSuccess) { cudaFree(d); return fail(SLAS_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cudaMemcpy(d, divisors, count * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemset(m, 0, sizeof(unsigned long long));
    divby_check_kernel<<<dim3((unsigned)ctx->sm_count * 8, (unsigned)count), 256, 0, ctx->stream>>>(d, m);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy(mismatches, m, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaFree(m);
    if (e != cudaSuccess) return fail(SLAS_ERR_CUDA, "divby self-check: %s", cudaGetErrorString(e));
    return SLAS_OK;
}
