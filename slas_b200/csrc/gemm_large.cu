// gemm_large.cu -- one big MatrixMul::matrix_mul (replaces the cblas_{s,d}gemm call of
// src/backends/blas.rs:94-109 for products too large for the batched static-shape kernels,
// e.g. the single 4096 x 4096 f32 contraction of BASELINE config 4).
//
// This is the REFERENCE-FAITHFUL path: fp32 (fp64) multiply-add on the FMA pipe, every C element
// accumulated in T in ascending k order inside one thread.  Classic register-tiled design:
//   CTA tile (16*TMICRO)^2 x 8, 256 threads, each thread a TMICRO x TMICRO block of C held as
//   2 x 2 sub-blocks one 16-byte vector wide (f32: 8x8, f64: 4x4) so every shared-memory
//   read is a conflict-free 128-bit load; A and B tiles are staged k-major in padded shared
//   memory, double buffered, with the next tile's global loads in flight during the FMAs.
// Any m, n, k, leading dimensions and op(A)/op(B) are accepted: interior tiles use 128-bit
// loads, edge tiles fall back to guarded scalar loads with zero fill.
//
// Where the time goes at 4096^3 (ncu --set full, profiles/r2i_ncu_summary.csv, r2i_ffma_stalls.txt): FMA pipe 69 %
// busy, 84 % of the loop's instructions are FFMA, the schedulers issue on ~90 % of the cycles (stall_dispatch 10 %,
// barrier 7 %); shared memory is conflict-free (wavefronts == ideal).  Two rewrites were measured and dropped
// (profiles/r2j_tune_large.log, r1b_tune.log): an interior-only kernel with 16 k-steps per stage and no guards (90.6 %
// FFMA in the loop) ran 48.0 TFLOP/s against this kernel's 49.6, a cp.async multi-stage ring 41.8 -- the kernel is not
// bound by its bookkeeping instructions; it sits at 0.80 of what a register-only FFMA loop reaches on this chip
// (60.9 TFLOP/s, microbench.cu).
#include "common.cuh"
#include "kernels.cuh"

namespace slas {

template <typename T> struct LargeCfg;
template <> struct LargeCfg<float> {
    static constexpr int VW = 4, TMICRO = 8, BM = 128, BN = 128, BK = 8, PAD = 4;
    using V = float4;
};
template <> struct LargeCfg<double> {
    static constexpr int VW = 2, TMICRO = 4, BM = 64, BN = 64, BK = 8, PAD = 2;
    using V = double2;
};

// Fetch this thread's one vector of a (BM|BN) x BK operand tile into registers.
// KCONTIG: the source is contiguous along k (A not transposed / B transposed); otherwise it is
// contiguous along the m (or n) index.
template <typename T, bool KCONTIG>
__device__ __forceinline__ void tile_fetch(T (&reg)[LargeCfg<T>::VW], const T *__restrict__ src, size_t ld,
                                           size_t row0, size_t rows, size_t k0, size_t kdim, bool vec_ok) {
    using C = LargeCfg<T>;
    constexpr int VW = C::VW;
    const int t = threadIdx.x;
    if (KCONTIG) {
        constexpr int VPR = C::BK / VW;  // vectors per row
        const size_t row = row0 + t / VPR, kk = k0 + (size_t)(t % VPR) * VW;
        if (vec_ok && row < rows && kk + VW <= kdim) {
            *reinterpret_cast<typename C::V *>(reg) = *reinterpret_cast<const typename C::V *>(src + row * ld + kk);
        } else {
#pragma unroll
            for (int i = 0; i < VW; ++i) reg[i] = (row < rows && kk + i < kdim) ? src[row * ld + kk + i] : T(0);
        }
    } else {
        constexpr int VPK = C::BM / VW;  // vectors per k row (BM == BN)
        const size_t kk = k0 + t / VPK, row = row0 + (size_t)(t % VPK) * VW;
        if (vec_ok && kk < kdim && row + VW <= rows) {
            *reinterpret_cast<typename C::V *>(reg) = *reinterpret_cast<const typename C::V *>(src + kk * ld + row);
        } else {
#pragma unroll
            for (int i = 0; i < VW; ++i) reg[i] = (kk < kdim && row + i < rows) ? src[kk * ld + row + i] : T(0);
        }
    }
}

// Store the fetched vector k-major into shared memory: S[k][row].
template <typename T, bool KCONTIG>
__device__ __forceinline__ void tile_stash(T (*S)[LargeCfg<T>::BM + LargeCfg<T>::PAD], const T (&reg)[LargeCfg<T>::VW]) {
    using C = LargeCfg<T>;
    constexpr int VW = C::VW;
    const int t = threadIdx.x;
    if (KCONTIG) {
        constexpr int VPR = C::BK / VW;
        const int row = t / VPR, kq = (t % VPR) * VW;
#pragma unroll
        for (int i = 0; i < VW; ++i) S[kq + i][row] = reg[i];
    } else {
        constexpr int VPK = C::BM / VW;
        const int kk = t / VPK, row = (t % VPK) * VW;
        *reinterpret_cast<typename C::V *>(&S[kk][row]) = *reinterpret_cast<const typename C::V *>(reg);
    }
}

template <typename T, bool TA, bool TB>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 2 : 1)
gemm_large_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t m, size_t n, size_t k,
                  size_t lda, size_t ldb, size_t ldc, int vec_a, int vec_b, int vec_c, size_t stride_a, size_t stride_b,
                  size_t stride_c) {
    using C = LargeCfg<T>;
    // blockIdx.z = object of a batch of medium-sized products (strides in elements; 0 for a single product)
    a += blockIdx.z * stride_a;
    b += blockIdx.z * stride_b;
    c += blockIdx.z * stride_c;
    constexpr int VW = C::VW, BM = C::BM, BN = C::BN, BK = C::BK, PAD = C::PAD;
    __shared__ __align__(16) T As[2][BK][BM + PAD];
    __shared__ __align__(16) T Bs[2][BK][BN + PAD];
    const size_t m0 = (size_t)blockIdx.y * BM, n0 = (size_t)blockIdx.x * BN;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;

    T acc[2 * VW][2 * VW];
#pragma unroll
    for (int i = 0; i < 2 * VW; ++i)
#pragma unroll
        for (int j = 0; j < 2 * VW; ++j) acc[i][j] = T(0);

    T ra[VW], rb[VW];
    // op(A)[i][p]: !TA -> a[i*lda + p] (k contiguous); TA -> a[p*lda + i] (m contiguous)
    tile_fetch<T, !TA>(ra, a, lda, m0, m, 0, k, vec_a);
    // op(B)[p][j]: !TB -> b[p*ldb + j] (n contiguous); TB -> b[j*ldb + p] (k contiguous)
    tile_fetch<T, TB>(rb, b, ldb, n0, n, 0, k, vec_b);
    tile_stash<T, !TA>(As[0], ra);
    tile_stash<T, TB>(Bs[0], rb);
    __syncthreads();

    const size_t ktiles = (k + BK - 1) / BK;
    for (size_t kt = 0; kt < ktiles; ++kt) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < ktiles) {
            tile_fetch<T, !TA>(ra, a, lda, m0, m, (kt + 1) * BK, k, vec_a);
            tile_fetch<T, TB>(rb, b, ldb, n0, n, (kt + 1) * BK, k, vec_b);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T av[2 * VW], bv[2 * VW];
            *reinterpret_cast<typename C::V *>(&av[0]) = *reinterpret_cast<const typename C::V *>(&As[cur][kk][ty * VW]);
            *reinterpret_cast<typename C::V *>(&av[VW]) = *reinterpret_cast<const typename C::V *>(&As[cur][kk][BM / 2 + ty * VW]);
            *reinterpret_cast<typename C::V *>(&bv[0]) = *reinterpret_cast<const typename C::V *>(&Bs[cur][kk][tx * VW]);
            *reinterpret_cast<typename C::V *>(&bv[VW]) = *reinterpret_cast<const typename C::V *>(&Bs[cur][kk][BN / 2 + tx * VW]);
#pragma unroll
            for (int i = 0; i < 2 * VW; ++i)
#pragma unroll
                for (int j = 0; j < 2 * VW; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) {
            tile_stash<T, !TA>(As[cur ^ 1], ra);
            tile_stash<T, TB>(Bs[cur ^ 1], rb);
        }
        __syncthreads();
    }

    // C rows: ty*VW + i (i < VW) and BM/2 + ty*VW + i; columns likewise with tx
#pragma unroll
    for (int hi = 0; hi < 2; ++hi)
#pragma unroll
        for (int i = 0; i < VW; ++i) {
            const size_t row = m0 + hi * (BM / 2) + ty * VW + i;
            if (row >= m) continue;
#pragma unroll
            for (int hj = 0; hj < 2; ++hj) {
                const size_t col = n0 + hj * (BN / 2) + tx * VW;
                T *dst = c + row * ldc + col;
                if (vec_c && col + VW <= n) {
                    *reinterpret_cast<typename C::V *>(dst) = *reinterpret_cast<const typename C::V *>(&acc[hi * VW + i][hj * VW]);
                } else {
#pragma unroll
                    for (int j = 0; j < VW; ++j)
                        if (col + j < n) dst[j] = acc[hi * VW + i][hj * VW + j];
                }
            }
        }
}

template <typename T>
static int launch_gemm_large_z(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m, size_t n,
                               size_t k, size_t lda, size_t ldb, size_t ldc, int ta, int tb, size_t stride_a, size_t stride_b,
                               size_t stride_c) {
    using C = LargeCfg<T>;
    (void)ctx;
    const size_t gx = (n + C::BN - 1) / C::BN, gy = (m + C::BM - 1) / C::BM;
    SLAS_REQUIRE(gx <= 0x7fffffffu && gy <= 65535, "gemm: grid too large (%zu x %zu tiles)", gx, gy);
    const int vec_a = aligned16(a) && (lda * sizeof(T)) % 16 == 0 && (stride_a * sizeof(T)) % 16 == 0;
    const int vec_b = aligned16(b) && (ldb * sizeof(T)) % 16 == 0 && (stride_b * sizeof(T)) % 16 == 0;
    const int vec_c = aligned16(c) && (ldc * sizeof(T)) % 16 == 0 && (stride_c * sizeof(T)) % 16 == 0;
    for (size_t done = 0; done < batch; done += 65535) {
        const size_t cnt = batch - done < 65535 ? batch - done : 65535;
        dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)cnt);
        const T *pa = a + done * stride_a, *pb = b + done * stride_b;
        T *pc = c + done * stride_c;
#define SLAS_LARGE_LAUNCH(TA_, TB_)                                                                                        \
        gemm_large_kernel<T, TA_, TB_><<<grid, 256, 0, st>>>(pa, pb, pc, m, n, k, lda, ldb, ldc, vec_a, vec_b, vec_c, stride_a, \
                                                             stride_b, stride_c)
        if (!ta && !tb) SLAS_LARGE_LAUNCH(false, false);
        else if (ta && !tb) SLAS_LARGE_LAUNCH(true, false);
        else if (!ta && tb) SLAS_LARGE_LAUNCH(false, true);
        else SLAS_LARGE_LAUNCH(true, true);
#undef SLAS_LARGE_LAUNCH
        SLAS_LAUNCH_CHECK();
    }
    return SLAS_OK;
}

template <typename T>
int launch_gemm_large(Context *ctx, cudaStream_t st, const T *a, const T *b, T *c, size_t m, size_t n, size_t k,
                      size_t lda, size_t ldb, size_t ldc, int ta, int tb) {
    if (m == 0 || n == 0) return SLAS_OK;
    if constexpr (sizeof(T) == 8) {
        // at least one full wave of 128 x 128 tiles: the big-tile DFMA kernel (gemm_large_f64.cu)
        if (((m + 127) / 128) * ((n + 127) / 128) >= (size_t)ctx->sm_count && tunable("gemm_large_variant") != 3)
            return launch_dgemm_big(ctx, st, a, b, c, m, n, k, lda, ldb, ldc, ta, tb);
    }
    return launch_gemm_large_z<T>(ctx, st, 1, a, b, c, m, n, k, lda, ldb, ldc, ta, tb, 0, 0, 0);
}

// A batch of medium-sized products (too big for the shared-memory staged batched kernels, 48x48 f32 / 40x40 f64 and up):
// the same register-tiled FFMA / DFMA kernel, one grid layer per object.
template <typename T>
int launch_gemm_large_batched(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m, size_t n,
                              size_t k, size_t lda, size_t ldb, size_t ldc, int ta, int tb, size_t stride_a, size_t stride_b,
                              size_t stride_c) {
    if (m == 0 || n == 0 || batch == 0) return SLAS_OK;
    return launch_gemm_large_z<T>(ctx, st, batch, a, b, c, m, n, k, lda, ldb, ldc, ta, tb, stride_a, stride_b, stride_c);
}
template int launch_gemm_large_batched<float>(Context *, cudaStream_t, size_t, const float *, const float *, float *, size_t,
                                              size_t, size_t, size_t, size_t, size_t, int, int, size_t, size_t, size_t);
template int launch_gemm_large_batched<double>(Context *, cudaStream_t, size_t, const double *, const double *, double *, size_t,
                                               size_t, size_t, size_t, size_t, size_t, int, int, size_t, size_t, size_t);
template int launch_gemm_large<float>(Context *, cudaStream_t, const float *, const float *, float *, size_t, size_t,
                                      size_t, size_t, size_t, size_t, int, int);
template int launch_gemm_large<double>(Context *, cudaStream_t, const double *, const double *, double *, size_t,
                                       size_t, size_t, size_t, size_t, size_t, int, int);

}  // namespace slas
