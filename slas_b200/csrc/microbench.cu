// microbench.cu -- roofline denominators measured on the box itself ("measure an
// FFMA-peak microbench on the box and record it").  Not part of the reference interface.
//   slas_cuda_bench_fma_peak: register-only 8x8 outer-product FMA loop (the instruction mix of the
//   GEMM inner loops, no memory traffic) -> sustained FP32 / FP64 FMA TFLOP/s at the clocks the
//   chip actually holds under that load.
#include "common.cuh"

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
    const double flop = 2.0 * 64 * 4 * (double)iters * 256.0 * blocks * reps;  // the 8 bookkeeping FMAs per round are not counted
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
