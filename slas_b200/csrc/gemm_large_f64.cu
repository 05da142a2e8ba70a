// gemm_large_f64.cu -- the big-tile DFMA kernel for one large f64 MatrixMul::matrix_mul (the cblas_dgemm call of
// src/backends/blas.rs:94-109): same design as gemm_large.cu (k-major padded shared tiles, double buffering,
// register-staged global loads, ascending-k FMA chain per C element), with a 128 x 128 x 8 CTA tile and an 8 x 8
// register tile per thread held as 4 x 4 sub-blocks of one double2 -- one 128-bit shared-memory load per 8 DFMAs
// instead of one per 4 with the 64 x 64 / 4 x 4 configuration of gemm_large.cu.  Measured at 4096^3: 22.3-23.8
// TFLOP/s (0.84-0.89 of the 26.6 TFLOP/s DFMA-peak microbench) against 18.3; used when the product has at least
// one full wave of 128 x 128 tiles, the 64 x 64 kernel keeps the smaller ones (more CTAs).
#include "common.cuh"
#include "kernels.cuh"

namespace slas {

template <typename T> struct BigCfg;
// VW = elements per 16-byte vector, SUB = vector sub-blocks per thread and dimension (micro tile = SUB*VW = 8),
// FV = 16-byte vectors of an operand tile each thread fetches per k-tile (BM * BK / VW / 256)
template <> struct BigCfg<double> {
    static constexpr int VW = 2, SUB = 4, BM = 128, BN = 128, BK = 8, PAD = 2, FV = 2;
    using V = double2;
};

// Fetch this thread's one vector of a (BM|BN) x BK operand tile into registers.
// KCONTIG: the source is contiguous along k (A not transposed / B transposed); otherwise it is
// contiguous along the m (or n) index.
template <typename T, bool KCONTIG>
__device__ __forceinline__ void big_tile_fetch(T (&reg)[BigCfg<T>::FV][BigCfg<T>::VW], const T *__restrict__ src, size_t ld,
                                           size_t row0, size_t rows, size_t k0, size_t kdim, bool vec_ok) {
    using C = BigCfg<T>;
    constexpr int VW = C::VW;
#pragma unroll
    for (int f = 0; f < C::FV; ++f) {
        const int t = threadIdx.x + f * 256;
        if (KCONTIG) {
            constexpr int VPR = C::BK / VW;  // vectors per row
            const size_t row = row0 + t / VPR, kk = k0 + (size_t)(t % VPR) * VW;
            if (vec_ok && row < rows && kk + VW <= kdim) {
                *reinterpret_cast<typename C::V *>(reg[f]) = *reinterpret_cast<const typename C::V *>(src + row * ld + kk);
            } else {
#pragma unroll
                for (int i = 0; i < VW; ++i) reg[f][i] = (row < rows && kk + i < kdim) ? src[row * ld + kk + i] : T(0);
            }
        } else {
            constexpr int VPK = C::BM / VW;  // vectors per k row (BM == BN)
            const size_t kk = k0 + t / VPK, row = row0 + (size_t)(t % VPK) * VW;
            if (vec_ok && kk < kdim && row + VW <= rows) {
                *reinterpret_cast<typename C::V *>(reg[f]) = *reinterpret_cast<const typename C::V *>(src + kk * ld + row);
            } else {
#pragma unroll
                for (int i = 0; i < VW; ++i) reg[f][i] = (kk < kdim && row + i < rows) ? src[kk * ld + row + i] : T(0);
            }
        }
    }
}

// Store the fetched vectors k-major into shared memory: S[k][row].
template <typename T, bool KCONTIG>
__device__ __forceinline__ void big_tile_stash(T (*S)[BigCfg<T>::BM + BigCfg<T>::PAD],
                                           const T (&reg)[BigCfg<T>::FV][BigCfg<T>::VW]) {
    using C = BigCfg<T>;
    constexpr int VW = C::VW;
#pragma unroll
    for (int f = 0; f < C::FV; ++f) {
        const int t = threadIdx.x + f * 256;
        if (KCONTIG) {
            constexpr int VPR = C::BK / VW;
            const int row = t / VPR, kq = (t % VPR) * VW;
#pragma unroll
            for (int i = 0; i < VW; ++i) S[kq + i][row] = reg[f][i];
        } else {
            constexpr int VPK = C::BM / VW;
            const int kk = t / VPK, row = (t % VPK) * VW;
            *reinterpret_cast<typename C::V *>(&S[kk][row]) = *reinterpret_cast<const typename C::V *>(reg[f]);
        }
    }
}

template <typename T, bool TA, bool TB>
__global__ void __launch_bounds__(256, 1)
gemm_big_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t m, size_t n, size_t k,
                  size_t lda, size_t ldb, size_t ldc, int vec_a, int vec_b, int vec_c) {
    using C = BigCfg<T>;
    constexpr int VW = C::VW, BM = C::BM, BN = C::BN, BK = C::BK, PAD = C::PAD, SUB = C::SUB, MT = C::SUB * C::VW;
    __shared__ __align__(16) T As[2][BK][BM + PAD];
    __shared__ __align__(16) T Bs[2][BK][BN + PAD];
    const size_t m0 = (size_t)blockIdx.y * BM, n0 = (size_t)blockIdx.x * BN;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;

    T acc[MT][MT];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) acc[i][j] = T(0);

    T ra[C::FV][VW], rb[C::FV][VW];
    // op(A)[i][p]: !TA -> a[i*lda + p] (k contiguous); TA -> a[p*lda + i] (m contiguous)
    big_tile_fetch<T, !TA>(ra, a, lda, m0, m, 0, k, vec_a);
    // op(B)[p][j]: !TB -> b[p*ldb + j] (n contiguous); TB -> b[j*ldb + p] (k contiguous)
    big_tile_fetch<T, TB>(rb, b, ldb, n0, n, 0, k, vec_b);
    big_tile_stash<T, !TA>(As[0], ra);
    big_tile_stash<T, TB>(Bs[0], rb);
    __syncthreads();

    const size_t ktiles = (k + BK - 1) / BK;
    for (size_t kt = 0; kt < ktiles; ++kt) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < ktiles) {
            big_tile_fetch<T, !TA>(ra, a, lda, m0, m, (kt + 1) * BK, k, vec_a);
            big_tile_fetch<T, TB>(rb, b, ldb, n0, n, (kt + 1) * BK, k, vec_b);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T av[MT], bv[MT];
#pragma unroll
            for (int sb = 0; sb < SUB; ++sb)
                *reinterpret_cast<typename C::V *>(&av[sb * VW]) =
                    *reinterpret_cast<const typename C::V *>(&As[cur][kk][sb * (BM / SUB) + ty * VW]);
#pragma unroll
            for (int sb = 0; sb < SUB; ++sb)
                *reinterpret_cast<typename C::V *>(&bv[sb * VW]) =
                    *reinterpret_cast<const typename C::V *>(&Bs[cur][kk][sb * (BN / SUB) + tx * VW]);
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < MT; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) {
            big_tile_stash<T, !TA>(As[cur ^ 1], ra);
            big_tile_stash<T, TB>(Bs[cur ^ 1], rb);
        }
        __syncthreads();
    }

    // C rows: hi * (BM / SUB) + ty*VW + i (hi < SUB, i < VW); columns likewise with tx
#pragma unroll
    for (int hi = 0; hi < SUB; ++hi)
#pragma unroll
        for (int i = 0; i < VW; ++i) {
            const size_t row = m0 + hi * (BM / SUB) + ty * VW + i;
            if (row >= m) continue;
#pragma unroll
            for (int hj = 0; hj < SUB; ++hj) {
                const size_t col = n0 + hj * (BN / SUB) + tx * VW;
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

int launch_dgemm_big(Context *ctx, cudaStream_t st, const double *a, const double *b, double *c, size_t m, size_t n, size_t k,
                     size_t lda, size_t ldb, size_t ldc, int ta, int tb) {
    using T = double;
    using C = BigCfg<T>;
    (void)ctx;
    if (m == 0 || n == 0) return SLAS_OK;
    const size_t gx = (n + C::BN - 1) / C::BN, gy = (m + C::BM - 1) / C::BM;
    SLAS_REQUIRE(gx <= 0x7fffffffu && gy <= 65535, "gemm: grid too large (%zu x %zu tiles)", gx, gy);
    const int vec_a = aligned16(a) && (lda * sizeof(T)) % 16 == 0;
    const int vec_b = aligned16(b) && (ldb * sizeof(T)) % 16 == 0;
    const int vec_c = aligned16(c) && (ldc * sizeof(T)) % 16 == 0;
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (!ta && !tb) gemm_big_kernel<T, false, false><<<grid, 256, 0, st>>>(a, b, c, m, n, k, lda, ldb, ldc, vec_a, vec_b, vec_c);
    else if (ta && !tb) gemm_big_kernel<T, true, false><<<grid, 256, 0, st>>>(a, b, c, m, n, k, lda, ldb, ldc, vec_a, vec_b, vec_c);
    else if (!ta && tb) gemm_big_kernel<T, false, true><<<grid, 256, 0, st>>>(a, b, c, m, n, k, lda, ldb, ldc, vec_a, vec_b, vec_c);
    else gemm_big_kernel<T, true, true><<<grid, 256, 0, st>>>(a, b, c, m, n, k, lda, ldb, ldc, vec_a, vec_b, vec_c);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

}  // namespace slas
