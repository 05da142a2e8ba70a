// gemm_small.cu -- batched static-shape matrix_mul: C_i = op(A_i) . op(B_i) for `batch` dense
// row-major objects laid out [[T; M*K]; batch], [[T; K*N]; batch] -> [[T; M*N]; batch].
// Replaces a loop of Matrix::matrix_mul calls (src/tensor.rs:441-454), each of which is one
// cblas_{s,d}gemm(RowMajor, opA, opB, m, n, k, 1, A, lda, B, ldb, 0, C, n)
// (src/backends/blas.rs:94-109).  The reference's compile-time shapes (MatrixShape<M, K>,
// src/tensor.rs:108-115) become template parameters here.
//
// These products are HBM-bound up to 32x32 (4x4 f32: 0.67 flop/byte), so the kernel is built
// around the memory system, not the FMA pipe:
//   * persistent CTAs, one producer warp + 8 consumer warps;
//   * the producer streams CHUNKS of G consecutive matrices of A and of B into a ring of
//     shared-memory stages with 1-D bulk TMA copies (cp.async.bulk -> UBLKCP) that complete on
//     an mbarrier -- a batch of static matrices is one contiguous byte range, so no tensor
//     map is needed and every HBM request is a full, aligned burst;
//   * consumers wait on the stage's "full" barrier, compute register-tiled TM x TN blocks of
//     C with FFMA/DFMA (fp32 products accumulate in fp32: the reference-faithful path) straight
//     out of shared memory, store C with 128-bit coalesced stores, and release the stage
//     through its "empty" barrier.
// Deterministic: every C element is one thread's fixed-order sum over k.
#include "common.cuh"
#include "kernels.cuh"
#include "gemm_small_kernel.cuh"

namespace slas {

template <typename T, typename Cfg>
static int launch_cfg(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, int ta, int tb) {
    const size_t nchunks = (batch + Cfg::G - 1) / Cfg::G;
    const size_t smem = Cfg::SMEM_BYTES;
    int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    size_t grid = (size_t)ctx->sm_count * per_sm;
    if (grid > nchunks) grid = nchunks;
    const unsigned threads = Cfg::CONSUMERS + 32;
#define SLAS_SMALL_LAUNCH(TA_, TB_)                                                                              \
    do {                                                                                                         \
        auto kern = gemm_small_kernel<T, Cfg, TA_, TB_>;                                                         \
        SLAS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        kern<<<(unsigned)grid, threads, smem, st>>>(a, b, c, batch, nchunks);                                    \
    } while (0)
    if (!ta && !tb) SLAS_SMALL_LAUNCH(false, false);
    else if (ta && !tb) SLAS_SMALL_LAUNCH(true, false);
    else if (!ta && tb) SLAS_SMALL_LAUNCH(false, true);
    else SLAS_SMALL_LAUNCH(true, true);
#undef SLAS_SMALL_LAUNCH
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

// ---- any small static shape: runtime (m, n, k), same staging ------------------------------------------------
// Shapes without a register-tiled specialisation (2x3 . 3x2, 8x16 . 16x4, ...) keep the memory side of the design
// -- persistent CTAs, bulk-TMA chunks of G consecutive matrices into a 3-stage ring, mbarriers -- and compute one
// C element per thread per step straight out of shared memory (A reads broadcast across a row of C, B reads
// consecutive), writing C with consecutive threads on consecutive elements.  G is a multiple of 4 so that every
// chunk is a whole number of 16-byte words whatever the shape.
template <typename T>
__global__ void __launch_bounds__(288)
gemm_small_generic_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t batch, size_t nchunks,
                          unsigned m, unsigned n, unsigned k, int ta, int tb, unsigned G, unsigned stage_elems) {
    constexpr int STAGES = 3, CONSUMERS = 256;
    extern __shared__ __align__(128) unsigned char smem[];
    T *stage_base = reinterpret_cast<T *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * stage_elems * sizeof(T));
    uint64_t *empty = full + STAGES;
    const unsigned mk = m * k, kn = k * n, mn = m * n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, CONSUMERS / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == CONSUMERS / 32) {
        if (lane == 0) {
            uint32_t it = 0;
            for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
                const uint32_t s = it % STAGES, use = it / STAGES;
                if (use > 0) mbar_wait_backoff(empty + s, (use - 1) & 1);
                const size_t first = ch * G;
                const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
                T *sa = stage_base + (size_t)s * stage_elems;
                T *sb = sa + (size_t)G * mk;
                // a ragged final chunk may end on a partial 16-byte word: round the copy up (the staging
                // buffer has room, the extra elements are never read; the arrays end on a 16-byte boundary
                // because the caller only takes this path when batch * m * k * sizeof(T) % 16 == 0)
                const uint32_t bytes_a = (cnt * mk * (uint32_t)sizeof(T) + 15u) & ~15u;
                const uint32_t bytes_b = (cnt * kn * (uint32_t)sizeof(T) + 15u) & ~15u;
                mbar_expect_tx(full + s, bytes_a + bytes_b);
                bulk_g2s(sa, a + first * mk, bytes_a, full + s);
                bulk_g2s(sb, b + first * kn, bytes_b, full + s);
            }
        }
        return;
    }
    uint32_t it = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        const size_t first = ch * G;
        const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
        const T *sa = stage_base + (size_t)s * stage_elems;
        const T *sb = sa + (size_t)G * mk;
        mbar_wait(full + s, use & 1);
        T *cc = c + first * mn;
        constexpr unsigned VW = 16 / sizeof(T);
        if (!tb && n % VW == 0) {
            // one 16-byte vector of a C row per thread: A scalar (broadcast), B and C 128-bit
            struct alignas(16) V { T v[VW]; };
            const unsigned nq = n / VW, per_mat = m * nq;
            for (unsigned e = threadIdx.x; e < cnt * per_mat; e += CONSUMERS) {
                const unsigned mat = e / per_mat, r = e - mat * per_mat, i = r / nq, jq = r - i * nq;
                const T *ma = sa + mat * mk;
                const V *mb = reinterpret_cast<const V *>(sb + mat * kn) + jq;
                V acc;
#pragma unroll
                for (unsigned q = 0; q < VW; ++q) acc.v[q] = T(0);
                for (unsigned p = 0; p < k; ++p) {
                    const T av = ta ? ma[p * m + i] : ma[i * k + p];
                    const V bv = mb[p * nq];
#pragma unroll
                    for (unsigned q = 0; q < VW; ++q) acc.v[q] = fma(av, bv.v[q], acc.v[q]);
                }
                *reinterpret_cast<V *>(cc + (size_t)mat * mn + i * n + jq * VW) = acc;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
            continue;
        }
        for (unsigned e = threadIdx.x; e < cnt * mn; e += CONSUMERS) {
            const unsigned mat = e / mn, r = e - mat * mn, i = r / n, j = r - i * n;
            const T *ma = sa + mat * mk, *mb = sb + mat * kn;
            T acc = T(0);
            if (!ta && !tb) for (unsigned p = 0; p < k; ++p) acc = fma(ma[i * k + p], mb[p * n + j], acc);
            else if (ta && !tb) for (unsigned p = 0; p < k; ++p) acc = fma(ma[p * m + i], mb[p * n + j], acc);
            else if (!ta && tb) for (unsigned p = 0; p < k; ++p) acc = fma(ma[i * k + p], mb[j * k + p], acc);
            else for (unsigned p = 0; p < k; ++p) acc = fma(ma[p * m + i], mb[j * k + p], acc);
            cc[e] = acc;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
}

template <typename T>
static int launch_generic_small(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m,
                                size_t n, size_t k, int ta, int tb) {
    const size_t mk = m * k, kn = k * n, per = (mk + kn) * sizeof(T);
    if (k == 0 || per == 0 || per > 16 * 1024) return SLAS_ERR_UNSUPPORTED;
    // the final (ragged) chunk is copied in whole 16-byte words: the arrays themselves must end on one
    if ((batch * mk * sizeof(T)) % 16 || (batch * kn * sizeof(T)) % 16) return SLAS_ERR_UNSUPPORTED;
    size_t G = (32 * 1024) / per / 4 * 4;
    if (G < 4) G = 4;
    if (G > 256) G = 256;
    const size_t stage_elems = ((G * (mk + kn) * sizeof(T) + 127) & ~(size_t)127) / sizeof(T);
    const size_t smem = 3 * stage_elems * sizeof(T) + 6 * sizeof(uint64_t);
    if (smem > 200 * 1024) return SLAS_ERR_UNSUPPORTED;
    const size_t nchunks = (batch + G - 1) / G;
    int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    size_t grid = (size_t)ctx->sm_count * per_sm;
    if (grid > nchunks) grid = nchunks;
    auto kern = gemm_small_generic_kernel<T>;
    SLAS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, 288, smem, st>>>(a, b, c, batch, nchunks, (unsigned)m, (unsigned)n, (unsigned)k, ta, tb, (unsigned)G,
                                            (unsigned)stage_elems);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

//                         T       M   N   K  TM TN CONS   G  STAGES
using CfgS4 = SmallCfg<float, 4, 4, 4, 1, 4, 256, 128, 4>;
using CfgS4t = SmallCfg<float, 4, 4, 4, 1, 4, 256, 128, 3, true>;   // bulk-store variants (benchmark knob gemm4s_variant)
using CfgS4u = SmallCfg<float, 4, 4, 4, 1, 4, 256, 256, 3, true>;
using CfgS4v3 = SmallCfg<float, 4, 4, 4, 1, 4, 256, 256, 4, true>;
using CfgS4v4 = SmallCfg<float, 4, 4, 4, 1, 4, 256, 512, 2, true>;
using CfgS4v5 = SmallCfg<float, 4, 4, 4, 1, 4, 256, 256, 4, false>;
using CfgS4v6 = SmallCfg<float, 4, 4, 4, 1, 4, 256, 256, 3, false>;
using CfgS8 = SmallCfg<float, 8, 8, 8, 2, 4, 256, 64, 3, true>;
using CfgS16 = SmallCfg<float, 16, 16, 16, 4, 4, 256, 16, 3>;
using CfgS32 = SmallCfg<float, 32, 32, 32, 8, 4, 256, 8, 3>;
using CfgD4 = SmallCfg<double, 4, 4, 4, 1, 4, 256, 128, 4>;
using CfgD4t = SmallCfg<double, 4, 4, 4, 1, 4, 256, 64, 3, true>;    // bulk-store variants (benchmark knob gemm4d_variant)
using CfgD4u = SmallCfg<double, 4, 4, 4, 1, 4, 256, 128, 3, true>;
using CfgD4v = SmallCfg<double, 4, 4, 4, 1, 2, 256, 64, 3, true>;
using CfgD4w = SmallCfg<double, 4, 4, 4, 1, 4, 256, 64, 4, false>;
using CfgD8 = SmallCfg<double, 8, 8, 8, 2, 4, 256, 32, 3, true>;
using CfgD16 = SmallCfg<double, 16, 16, 16, 4, 4, 256, 16, 3>;
using CfgD32 = SmallCfg<double, 32, 32, 32, 8, 4, 128, 4, 3>;
// 64 x 64: 32 / 64 KB per pair -- compute-bound (10.7 flop/byte), two / one matrices per stage
using CfgS64 = SmallCfg<float, 64, 64, 64, 8, 8, 128, 2, 3>;
using CfgD64 = SmallCfg<double, 64, 64, 64, 8, 4, 128, 1, 3>;
using CfgS64b = SmallCfg<float, 64, 64, 64, 8, 4, 256, 2, 3>;
using CfgS64c = SmallCfg<float, 64, 64, 64, 4, 8, 256, 2, 3>;
// benchmark variants (slas_cuda_set_tunable)
using CfgD32b = SmallCfg<double, 32, 32, 32, 4, 4, 256, 4, 3>;
using CfgS32b = SmallCfg<float, 32, 32, 32, 8, 8, 128, 8, 3>;
using CfgS16b = SmallCfg<float, 16, 16, 16, 4, 4, 256, 32, 3>;
using CfgS16c = SmallCfg<float, 16, 16, 16, 8, 4, 256, 32, 3>;
using CfgS16t = SmallCfg<float, 16, 16, 16, 8, 4, 256, 32, 2, true>;
using CfgS16v4 = SmallCfg<float, 16, 16, 16, 8, 4, 128, 16, 3, true>;   // two CTAs per SM, bulk C store
using CfgS16v5 = SmallCfg<float, 16, 16, 16, 8, 4, 128, 16, 3, false>;
using CfgS16v6 = SmallCfg<float, 16, 16, 16, 8, 4, 64, 8, 3, true>;     // four CTAs per SM
using CfgS32v4 = SmallCfg<float, 32, 32, 32, 8, 4, 128, 4, 3, true>;
using CfgS32v5 = SmallCfg<float, 32, 32, 32, 8, 4, 128, 4, 3, false>;
using CfgS32t = SmallCfg<float, 32, 32, 32, 8, 4, 256, 8, 2, true>;
using CfgD16t = SmallCfg<double, 16, 16, 16, 4, 4, 256, 16, 2, true>;

// Shapes without a built-in cube configuration: thread per matrix (every dimension <= 4), the run-time specialised
// kernels, the runtime-shape kernel.  All of them stream whole 16-byte words, so a batch whose byte count is not a
// multiple of 16 (1001 3x3 f32 matrices ...) is split: the largest prefix that is goes through them, the last one to
// three objects through the plain kernel -- instead of the whole batch falling off the fast path.
template <typename T>
static int launch_other_shapes(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m, size_t n,
                               size_t k, int ta, int tb) {
    // f64 objects of 256 bytes whose N and K are whole 16-byte vectors (2x4 . 4x4, 4x4 . 4x2, 4x2 . 2x4): a thread per
    // matrix reads power-of-two strided objects through 4- to 8-way shared-memory bank conflicts (0.77-0.80 of HBM peak);
    // the run-time specialised register-tile kernel splits a matrix over threads that read consecutive vectors (0.85-0.87,
    // profiles/r3t_tune_tinyvsjit.log).  Without NVRTC the thread-per-matrix kernel below still takes them.
    if (sizeof(T) == 8 && m <= 4 && n <= 4 && k <= 4 && n % 2 == 0 && k % 2 == 0 && (m * k + k * n + m * n) * sizeof(T) >= 256 &&
        !(m == 4 && n == 4 && k == 4) && tunable("gemm_tiny_variant") != 2) {
        const int rj = launch_gemm_jit_batched<T>(ctx, st, batch, a, b, c, m, n, k, ta, tb);
        if (rj != SLAS_ERR_UNSUPPORTED) return rj;
    }
    const int r = launch_gemm_tiny_batched<T>(ctx, st, batch, a, b, c, m, n, k, ta, tb);
    if (r != SLAS_ERR_UNSUPPORTED) return r;
    const int rj = launch_gemm_jit_batched<T>(ctx, st, batch, a, b, c, m, n, k, ta, tb);
    if (rj != SLAS_ERR_UNSUPPORTED) return rj;
    return batch >= 64 ? launch_generic_small<T>(ctx, st, batch, a, b, c, m, n, k, ta, tb) : SLAS_ERR_UNSUPPORTED;
}
template <typename T>
static int launch_other_shapes_split(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m,
                                     size_t n, size_t k, int ta, int tb) {
    const bool ragged = (batch * m * k * sizeof(T)) % 16 || (batch * k * n * sizeof(T)) % 16;
    if (!ragged) return launch_other_shapes<T>(ctx, st, batch, a, b, c, m, n, k, ta, tb);
    const size_t unit = 16 / sizeof(T), rem = batch % unit, main = batch - rem;  // `unit` objects are always whole words
    if (main < 1024 || rem == 0) return SLAS_ERR_UNSUPPORTED;
    const int r = launch_other_shapes<T>(ctx, st, main, a, b, c, m, n, k, ta, tb);
    if (r != SLAS_OK) return r;
    return launch_gemm_generic<T>(ctx, st, rem, a + main * m * k, b + main * k * n, c + main * m * n, m, n, k, ta ? m : k,
                                  tb ? k : n, n, ta, tb, m * k, k * n, m * n);
}

template <>
int launch_gemm_small_batched<float>(Context *ctx, cudaStream_t st, size_t batch, const float *a, const float *b,
                                     float *c, size_t m, size_t n, size_t k, int ta, int tb) {
    if (batch == 0) return SLAS_OK;
    if (!(aligned16(a) && aligned16(b) && aligned16(c))) return SLAS_ERR_UNSUPPORTED;
    if (m != n || n != k || (m != 4 && m != 8 && m != 16 && m != 32 && m != 64)) {
        return launch_other_shapes_split<float>(ctx, st, batch, a, b, c, m, n, k, ta, tb);
    }
    switch (m) {
    case 4:
        // measured on B200 (2^24 pairs): 256 pairs per stage + bulk C store 0.490 ms (1.00 of the HBM copy
        // peak); 128 per stage with per-warp stores 0.502-0.506; the other knob values are the rest of the sweep
        if (tunable("gemm4s_variant") == 1) return launch_cfg<float, CfgS4t>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm4s_variant") == 7) return launch_cfg<float, CfgS4>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm4s_variant") == 3) return launch_cfg<float, CfgS4v3>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm4s_variant") == 4) return launch_cfg<float, CfgS4v4>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm4s_variant") == 5) return launch_cfg<float, CfgS4v5>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm4s_variant") == 6) return launch_cfg<float, CfgS4v6>(ctx, st, batch, a, b, c, ta, tb);
        return launch_cfg<float, CfgS4u>(ctx, st, batch, a, b, c, ta, tb);
    case 8: return launch_cfg<float, CfgS8>(ctx, st, batch, a, b, c, ta, tb);
    case 16:
        // measured on B200 (2^20 pairs, profiles/r3v_tune_gemm16s.log): 8x4 tiles, 128 consumers, 16 matrices per stage,
        // bulk C store = TWO CTAs per SM: 1.00 of the HBM copy peak; 256 consumers and 32 per stage (one CTA per SM) 0.96,
        // the same with per-warp stores 0.96, four CTAs of 64 consumers 0.86
        if (tunable("gemm16s_variant") == 1) return launch_cfg<float, CfgS16b>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm16s_variant") == 3) return launch_cfg<float, CfgS16t>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm16s_variant") == 2) return launch_cfg<float, CfgS16>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm16s_variant") == 4) return launch_cfg<float, CfgS16c>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm16s_variant") == 5) return launch_cfg<float, CfgS16v5>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm16s_variant") == 6) return launch_cfg<float, CfgS16v6>(ctx, st, batch, a, b, c, ta, tb);
        return launch_cfg<float, CfgS16v4>(ctx, st, batch, a, b, c, ta, tb);
    case 64:
        // measured (B200, batch 2^18): 8x4 tiles on 8 consumer warps 39.8 TFLOP/s (0.74 of the FMA-peak
        // microbench), 4x8 38.7, 8x8 tiles on 4 warps 20.7
        if (tunable("gemm32s_variant") == 1) return launch_cfg<float, CfgS64>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm32s_variant") == 2) return launch_cfg<float, CfgS64c>(ctx, st, batch, a, b, c, ta, tb);
        return launch_cfg<float, CfgS64b>(ctx, st, batch, a, b, c, ta, tb);
    case 32:
        if (tunable("gemm32s_variant") == 1) return launch_cfg<float, CfgS32b>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm32s_variant") == 3) return launch_cfg<float, CfgS32t>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm32s_variant") == 4) return launch_cfg<float, CfgS32v4>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm32s_variant") == 5) return launch_cfg<float, CfgS32v5>(ctx, st, batch, a, b, c, ta, tb);
        return launch_cfg<float, CfgS32>(ctx, st, batch, a, b, c, ta, tb);
    }
    return SLAS_ERR_UNSUPPORTED;
}

template <>
int launch_gemm_small_batched<double>(Context *ctx, cudaStream_t st, size_t batch, const double *a, const double *b,
                                      double *c, size_t m, size_t n, size_t k, int ta, int tb) {
    if (batch == 0) return SLAS_OK;
    if (!(aligned16(a) && aligned16(b) && aligned16(c))) return SLAS_ERR_UNSUPPORTED;
    if (m != n || n != k || (m != 4 && m != 8 && m != 16 && m != 32 && m != 64)) {
        return launch_other_shapes_split<double>(ctx, st, batch, a, b, c, m, n, k, ta, tb);
    }
    switch (m) {
    case 4:
        // measured on B200 (2^23 pairs, profiles/r3u_tune_gemm4d.log): 64 pairs per stage, three stages, bulk C store =
        // 1.08 of the HBM copy peak (three CTAs per SM); 128 per stage with per-warp stores (one CTA per SM) 0.89
        if (tunable("gemm4d_variant") == 1) return launch_cfg<double, CfgD4>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm4d_variant") == 2) return launch_cfg<double, CfgD4u>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm4d_variant") == 3) return launch_cfg<double, CfgD4v>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm4d_variant") == 4) return launch_cfg<double, CfgD4w>(ctx, st, batch, a, b, c, ta, tb);
        return launch_cfg<double, CfgD4t>(ctx, st, batch, a, b, c, ta, tb);
    case 8: return launch_cfg<double, CfgD8>(ctx, st, batch, a, b, c, ta, tb);
    case 16: return launch_cfg<double, CfgD16>(ctx, st, batch, a, b, c, ta, tb);
    case 64: return launch_cfg<double, CfgD64>(ctx, st, batch, a, b, c, ta, tb);
    case 32:
        if (tunable("gemm32d_variant") == 1) return launch_cfg<double, CfgD32b>(ctx, st, batch, a, b, c, ta, tb);
        return launch_cfg<double, CfgD32>(ctx, st, batch, a, b, c, ta, tb);
    }
    return SLAS_ERR_UNSUPPORTED;
}

}  // namespace slas
