// gemm_ffma.cu -- fast path of the reference-faithful f32 MatrixMul::matrix_mul for ONE big product
// (cblas_sgemm of src/backends/blas.rs:94-109, BASELINE config 4): fp32 FFMA, every C element an
// ascending-k fp32 accumulation inside one thread.
//
// 128 x 128 x 16 CTA tile, 256 threads, 8 x 8 register tile per thread (2 x 2 blocks of one float4).
// Both operands go global -> shared with 16-byte cp.async (LDGSTS, zero-fill at the edges) in
// whatever orientation they are stored, so no register staging and no transposing stores:
//   k-contiguous operand  (A not transposed / B transposed):  S[mn][BK + 4]  -> the compute loop
//       reads 4 consecutive k of one row with one LDS.128 (8 loads per 4 k-steps);
//   mn-contiguous operand (A transposed / B not transposed):  S[BK][128 + 4] -> 4 consecutive
//       rows/columns of one k with one LDS.128 (2 loads per k-step).
// Either way 16 LDS.128 feed 256 FFMA.  A 3-stage cp.async ring keeps two tiles in flight; one
// __syncthreads per 16-wide k-tile; 2 CTAs per SM.
// Requirements (else the caller uses gemm_large.cu): 16-byte aligned A, B and lda, ldb multiples of 4.
#include "common.cuh"
#include "kernels.cuh"

namespace slas {

namespace ffma {
constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, PAD = 4, THREADS = 256;
constexpr int KC_PITCH = BK + PAD;    // k-contiguous tile: 128 rows of 20 floats
constexpr int MN_PITCH = BM + PAD;    // mn-contiguous tile: 16 rows of 132 floats
constexpr int TILE_FLOATS = BM * KC_PITCH > BK * MN_PITCH ? BM * KC_PITCH : BK * MN_PITCH;  // 2560
constexpr int STAGE_FLOATS = 2 * TILE_FLOATS;
constexpr int SMEM_BYTES = STAGES * STAGE_FLOATS * 4;  // 61440
}  // namespace ffma

__device__ __forceinline__ void cp_async16(float *dst_smem, const float *src, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Issue this thread's two 16-byte copies of one operand tile.  KCONTIG: rows along mn, k contiguous.
template <bool KCONTIG>
__device__ __forceinline__ void tile_issue(float *S, const float *__restrict__ src, size_t ld, size_t mn0, size_t mn_dim,
                                           size_t k0, size_t kdim) {
    using namespace ffma;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = threadIdx.x + i * THREADS;
        if (KCONTIG) {
            const int row = c >> 2, kc = (c & 3) * 4;
            const size_t gr = mn0 + row, gk = k0 + kc;
            int bytes = 0;
            if (gr < mn_dim && gk < kdim) bytes = (int)((kdim - gk) >= 4 ? 16 : (kdim - gk) * 4);
            cp_async16(S + row * KC_PITCH + kc, bytes ? src + gr * ld + gk : src, bytes);
        } else {
            const int kk = c >> 5, mc = (c & 31) * 4;
            const size_t gk = k0 + kk, gm = mn0 + mc;
            int bytes = 0;
            if (gk < kdim && gm < mn_dim) bytes = (int)((mn_dim - gm) >= 4 ? 16 : (mn_dim - gm) * 4);
            cp_async16(S + kk * MN_PITCH + mc, bytes ? src + gk * ld + gm : src, bytes);
        }
    }
}

// Fetch this thread's 8 values (two float4 halves at h0 and 64 + h0) for 4 consecutive k.
// out[kk][0..7]
template <bool KCONTIG>
__device__ __forceinline__ void frag_load(float (&out)[4][8], const float *S, int h0, int k4) {
    using namespace ffma;
    if (KCONTIG) {
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4 *>(S + (half * 64 + h0 + i) * KC_PITCH + k4);
                out[0][half * 4 + i] = v.x; out[1][half * 4 + i] = v.y; out[2][half * 4 + i] = v.z; out[3][half * 4 + i] = v.w;
            }
    } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const float4 v = *reinterpret_cast<const float4 *>(S + (k4 + kk) * MN_PITCH + half * 64 + h0);
                out[kk][half * 4 + 0] = v.x; out[kk][half * 4 + 1] = v.y; out[kk][half * 4 + 2] = v.z; out[kk][half * 4 + 3] = v.w;
            }
    }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(ffma::THREADS, 2)
sgemm_ffma_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ c, size_t m, size_t n,
                  size_t k, size_t lda, size_t ldb, size_t ldc, int vec_c) {
    using namespace ffma;
    extern __shared__ __align__(16) float smem_f[];
    const size_t m0 = (size_t)blockIdx.y * BM, n0 = (size_t)blockIdx.x * BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const size_t ktiles = (k + BK - 1) / BK;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    auto issue = [&](size_t kt) {
        float *sa = smem_f + (kt % STAGES) * STAGE_FLOATS, *sb = sa + TILE_FLOATS;
        // op(A)[i][p]: !TA -> a[i*lda + p] (k contiguous); TA -> a[p*lda + i] (m contiguous)
        tile_issue<!TA>(sa, a, lda, m0, m, kt * BK, k);
        // op(B)[p][j]: !TB -> b[p*ldb + j] (n contiguous); TB -> b[j*ldb + p] (k contiguous)
        tile_issue<TB>(sb, b, ldb, n0, n, kt * BK, k);
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if ((size_t)s < ktiles) issue(s);
        cp_async_commit();
    }
    for (size_t kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();   // tile kt has landed (this thread's copies) ...
        __syncthreads();               // ... and everyone's; also: everyone is done reading tile kt-1's slot
        if (kt + STAGES - 1 < ktiles) issue(kt + STAGES - 1);
        cp_async_commit();
        const float *sa = smem_f + (kt % STAGES) * STAGE_FLOATS, *sb = sa + TILE_FLOATS;
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            float av[4][8], bv[4][8];
            frag_load<!TA>(av, sa, ty * 4, k4);
            frag_load<TB>(bv, sb, tx * 4, k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[kk][i], bv[kk][j], acc[i][j]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int hi = 0; hi < 2; ++hi)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const size_t row = m0 + hi * 64 + ty * 4 + i;
            if (row >= m) continue;
#pragma unroll
            for (int hj = 0; hj < 2; ++hj) {
                const size_t col = n0 + hj * 64 + tx * 4;
                float *dst = c + row * ldc + col;
                const float *src = &acc[hi * 4 + i][hj * 4];
                if (vec_c && col + 4 <= n) {
                    *reinterpret_cast<float4 *>(dst) = make_float4(src[0], src[1], src[2], src[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < n) dst[j] = src[j];
                }
            }
        }
}

int launch_sgemm_ffma(Context *ctx, cudaStream_t st, const float *a, const float *b, float *c, size_t m, size_t n, size_t k,
                      size_t lda, size_t ldb, size_t ldc, int ta, int tb) {
    using namespace ffma;
    (void)ctx;
    if (!aligned16(a) || !aligned16(b) || lda % 4 || ldb % 4) return SLAS_ERR_UNSUPPORTED;
    if (m == 0 || n == 0) return SLAS_OK;
    const size_t gx = (n + BN - 1) / BN, gy = (m + BM - 1) / BM;
    if (gx > 0x7fffffffu || gy > 65535) return SLAS_ERR_UNSUPPORTED;
    const int vec_c = aligned16(c) && ldc % 4 == 0;
    dim3 grid((unsigned)gx, (unsigned)gy);
#define SLAS_FFMA_LAUNCH(TA_, TB_)                                                                                   \
    do {                                                                                                             \
        auto kern = sgemm_ffma_kernel<TA_, TB_>;                                                                     \
        SLAS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));          \
        kern<<<grid, THREADS, SMEM_BYTES, st>>>(a, b, c, m, n, k, lda, ldb, ldc, vec_c);                             \
    } while (0)
    if (!ta && !tb) SLAS_FFMA_LAUNCH(false, false);
    else if (ta && !tb) SLAS_FFMA_LAUNCH(true, false);
    else if (!ta && tb) SLAS_FFMA_LAUNCH(false, true);
    else SLAS_FFMA_LAUNCH(true, true);
#undef SLAS_FFMA_LAUNCH
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

}  // namespace slas
