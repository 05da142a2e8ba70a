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
#include "async.cuh"

namespace slas {

// ---- register-tile helpers -----------------------------------------------------------------------
template <typename T> struct KV { static constexpr int N = 16 / (int)sizeof(T); };  // elements per 16-byte vector

// load CNT contiguous elements (CNT * sizeof(T) a multiple of 16, 16-byte aligned) from shared memory
template <typename T, int CNT>
__device__ __forceinline__ void lds_vec(T (&dst)[CNT], const T *p) {
    static_assert((CNT * sizeof(T)) % 16 == 0, "vector loads move whole 16-byte words");
    constexpr int NV = CNT * (int)sizeof(T) / 16;
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < NV; ++i) d[i] = src[i];
}
template <typename T, int CNT>
__device__ __forceinline__ void stg_vec(T *p, const T (&src)[CNT]) {
    constexpr int NV = CNT * (int)sizeof(T) / 16;
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(p);
#pragma unroll
    for (int i = 0; i < NV; ++i) __stcs(d + i, s[i]);
}

template <typename T, int M_, int N_, int K_, int TM_, int TN_, int CONS_, int G_, int STAGES_, bool STORE_TMA_ = false>
struct SmallCfg {
    static constexpr int M = M_, N = N_, K = K_, TM = TM_, TN = TN_;
    // STORE_TMA: C goes registers -> a double-buffered shared staging tile -> ONE bulk copy per chunk
    // (cp.async.bulk shared -> global), so HBM sees long write bursts instead of per-warp 512-byte stores
    static constexpr bool STORE_TMA = STORE_TMA_;
    static constexpr size_t CSTAGE_BYTES = STORE_TMA_ ? (size_t)G_ * M_ * N_ * sizeof(T) : 0;
    static constexpr int CONSUMERS = CONS_;           // consumer threads
    static constexpr int G = G_;                      // matrices per stage
    static constexpr int STAGES = STAGES_;
    static constexpr int TPM = (M / TM) * (N / TN);   // threads per matrix
    static constexpr int MPP = CONSUMERS / TPM;       // matrices per pass
    static constexpr int PASSES = G / MPP;
    static constexpr int A_ELEMS = M * K, B_ELEMS = K * N, C_ELEMS = M * N;
    static constexpr size_t STAGE_BYTES = (size_t)G * (A_ELEMS + B_ELEMS) * sizeof(T);
    static constexpr size_t SMEM_BYTES = STAGES * STAGE_BYTES + 2 * CSTAGE_BYTES + 2 * STAGES * sizeof(uint64_t);
    static_assert(M % TM == 0 && N % TN == 0, "tile must divide the matrix");
    static_assert(CONSUMERS % TPM == 0 && G % MPP == 0, "chunk must be whole passes");
    static_assert((A_ELEMS * sizeof(T)) % 16 == 0 && (B_ELEMS * sizeof(T)) % 16 == 0, "bulk copies move 16-byte words");
    static_assert((TN * sizeof(T)) % 16 == 0, "C rows are stored with 128-bit stores");
};

template <typename T, typename Cfg, bool TA, bool TB>
__global__ void __launch_bounds__(Cfg::CONSUMERS + 32)
gemm_small_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t batch, size_t nchunks) {
    constexpr int M = Cfg::M, N = Cfg::N, K = Cfg::K, TM = Cfg::TM, TN = Cfg::TN;
    constexpr int G = Cfg::G, STAGES = Cfg::STAGES, TPM = Cfg::TPM, MPP = Cfg::MPP, PASSES = Cfg::PASSES;
    constexpr int KVN = KV<T>::N;
    constexpr int NCHUNK = K / KVN;
    static_assert(K % KVN == 0, "K must be a whole number of 16-byte vectors");

    extern __shared__ __align__(128) unsigned char smem[];
    T *stage_base = reinterpret_cast<T *>(smem);
    T *cstage_base = reinterpret_cast<T *>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * Cfg::STAGE_BYTES + 2 * Cfg::CSTAGE_BYTES);
    uint64_t *empty = full + STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, Cfg::CONSUMERS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == Cfg::CONSUMERS / 32) {
        // ===== producer warp: one elected lane issues the bulk copies =====
        if (lane == 0) {
            uint32_t it = 0;
            for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
                const uint32_t s = it % STAGES, use = it / STAGES;
                if (use > 0) mbar_wait(empty + s, (use - 1) & 1);
                const size_t first = ch * G;
                const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
                T *sa = stage_base + (size_t)s * (Cfg::STAGE_BYTES / sizeof(T));
                T *sb = sa + (size_t)G * Cfg::A_ELEMS;
                const uint32_t bytes_a = cnt * Cfg::A_ELEMS * (uint32_t)sizeof(T);
                const uint32_t bytes_b = cnt * Cfg::B_ELEMS * (uint32_t)sizeof(T);
                mbar_expect_tx(full + s, bytes_a + bytes_b);
                bulk_g2s(sa, a + first * Cfg::A_ELEMS, bytes_a, full + s);
                bulk_g2s(sb, b + first * Cfg::B_ELEMS, bytes_b, full + s);
            }
        }
        return;
    }

    // ===== consumers =====
    const int t = threadIdx.x % TPM;          // thread within its matrix
    const int slot = threadIdx.x / TPM;       // matrix within the pass
    constexpr int TJ = N / TN, TI = M / TM;
    const int tj = t % TJ, ti = t / TJ;
    // row s of this thread's tile: interleaved for row-major A (conflict-free 128-bit smem reads
    // across ti), contiguous for transposed A (its TM rows are one vector of A^T's storage)
    auto row_of = [&](int s) { return TA ? ti * TM + s : ti + s * TI; };
    // columns of this thread's tile: NG groups of one 16-byte vector each, group g at
    // g * (N / NG) + tj * KVN -- so the lanes of a matrix read (and store) CONTIGUOUS 16-byte pieces of
    // a row in every instruction (a 32-byte-per-lane footprint would 2-way bank-conflict for f64)
    constexpr int NG = TN / KVN;
    static_assert(TN % KVN == 0, "TN must be whole 16-byte vectors");
    auto col_of = [&](int j) { return (j / KVN) * (N / NG) + tj * KVN + (j % KVN); };
    const int rot = TB ? tj : 0;  // k-chunk rotation: spreads transposed-B reads over the banks

    uint32_t it = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        const size_t first = ch * G;
        const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
        const T *sa = stage_base + (size_t)s * (Cfg::STAGE_BYTES / sizeof(T));
        const T *sb = sa + (size_t)G * Cfg::A_ELEMS;
        mbar_wait(full + s, use & 1);
        T *cst = cstage_base + (size_t)(it & 1) * (Cfg::CSTAGE_BYTES / sizeof(T));
        // (bulk store) the copy out of this staging buffer issued two chunks ago has been read: thread 0 waited
        // for it (wait_group.read 1) before arriving here
        if (Cfg::STORE_TMA) asm volatile("bar.sync 1, %0;" ::"n"(Cfg::CONSUMERS) : "memory");
#pragma unroll 1
        for (int p = 0; p < PASSES; ++p) {
            const uint32_t mat = p * MPP + slot;
            if (mat >= cnt) break;
            const T *ma = sa + (size_t)mat * Cfg::A_ELEMS;
            const T *mb = sb + (size_t)mat * Cfg::B_ELEMS;
            T acc[TM][TN];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = T(0);
#pragma unroll
            for (int kc = 0; kc < NCHUNK; ++kc) {
                const int k0 = ((kc + rot) % NCHUNK) * KVN;
                T ablk[TM][KVN], bblk[KVN][TN];
                if (!TA) {
#pragma unroll
                    for (int i = 0; i < TM; ++i) lds_vec<T, KVN>(ablk[i], ma + row_of(i) * K + k0);
                } else {
#pragma unroll
                    for (int kk = 0; kk < KVN; ++kk) {
                        if constexpr ((TM * sizeof(T)) % 16 == 0) {
                            T col[TM];
                            lds_vec<T, TM>(col, ma + (k0 + kk) * M + ti * TM);
#pragma unroll
                            for (int i = 0; i < TM; ++i) ablk[i][kk] = col[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < TM; ++i) ablk[i][kk] = ma[(k0 + kk) * M + row_of(i)];
                        }
                    }
                }
                if (!TB) {
#pragma unroll
                    for (int kk = 0; kk < KVN; ++kk)
#pragma unroll
                        for (int g = 0; g < NG; ++g)
                            lds_vec<T, KVN>(*reinterpret_cast<T(*)[KVN]>(&bblk[kk][g * KVN]), mb + (k0 + kk) * N + col_of(g * KVN));
                } else {
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        T row[KVN];
                        lds_vec<T, KVN>(row, mb + col_of(j) * K + k0);
#pragma unroll
                        for (int kk = 0; kk < KVN; ++kk) bblk[kk][j] = row[kk];
                    }
                }
#pragma unroll
                for (int kk = 0; kk < KVN; ++kk)
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[i][j] = fma(ablk[i][kk], bblk[kk][j], acc[i][j]);
            }
            if (Cfg::STORE_TMA) {
                T *mc = cst + (size_t)mat * Cfg::C_ELEMS;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int g = 0; g < NG; ++g)
                        *reinterpret_cast<uint4 *>(mc + row_of(i) * N + col_of(g * KVN)) = *reinterpret_cast<const uint4 *>(&acc[i][g * KVN]);
            } else {
                T *mc = c + (first + mat) * Cfg::C_ELEMS;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int g = 0; g < NG; ++g)
                        stg_vec<T, KVN>(mc + row_of(i) * N + col_of(g * KVN), *reinterpret_cast<const T(*)[KVN]>(&acc[i][g * KVN]));
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (Cfg::STORE_TMA) {
            // make the generic-proxy writes of every consumer visible to the async proxy, then one bulk store
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 2, %0;" ::"n"(Cfg::CONSUMERS) : "memory");
            if (threadIdx.x == 0) {
                const uint32_t bytes_c = cnt * Cfg::C_ELEMS * (uint32_t)sizeof(T);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(c + first * Cfg::C_ELEMS),
                             "r"(smem_u32(cst)), "r"(bytes_c)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the OTHER buffer is free again
            }
        }
    }
    if (Cfg::STORE_TMA && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

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
                if (use > 0) mbar_wait(empty + s, (use - 1) & 1);
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
using CfgS32t = SmallCfg<float, 32, 32, 32, 8, 4, 256, 8, 2, true>;
using CfgD16t = SmallCfg<double, 16, 16, 16, 4, 4, 256, 16, 2, true>;

template <>
int launch_gemm_small_batched<float>(Context *ctx, cudaStream_t st, size_t batch, const float *a, const float *b,
                                     float *c, size_t m, size_t n, size_t k, int ta, int tb) {
    if (batch == 0) return SLAS_OK;
    if (!(aligned16(a) && aligned16(b) && aligned16(c))) return SLAS_ERR_UNSUPPORTED;
    if (m != n || n != k || (m != 4 && m != 8 && m != 16 && m != 32 && m != 64)) {
        const int r = launch_gemm_tiny_batched<float>(ctx, st, batch, a, b, c, m, n, k, ta, tb);
        if (r != SLAS_ERR_UNSUPPORTED) return r;
        return batch >= 64 ? launch_generic_small<float>(ctx, st, batch, a, b, c, m, n, k, ta, tb) : SLAS_ERR_UNSUPPORTED;
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
        // measured on B200 (profiles/r1b_tune.log): 8x4 tiles, 32 matrices per stage = 0.96 of HBM peak
        if (tunable("gemm16s_variant") == 1) return launch_cfg<float, CfgS16b>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm16s_variant") == 3) return launch_cfg<float, CfgS16t>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm16s_variant") == 2) return launch_cfg<float, CfgS16>(ctx, st, batch, a, b, c, ta, tb);
        return launch_cfg<float, CfgS16c>(ctx, st, batch, a, b, c, ta, tb);
    case 64:
        // measured (B200, batch 2^18): 8x4 tiles on 8 consumer warps 39.8 TFLOP/s (0.74 of the FMA-peak
        // microbench), 4x8 38.7, 8x8 tiles on 4 warps 20.7
        if (tunable("gemm32s_variant") == 1) return launch_cfg<float, CfgS64>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm32s_variant") == 2) return launch_cfg<float, CfgS64c>(ctx, st, batch, a, b, c, ta, tb);
        return launch_cfg<float, CfgS64b>(ctx, st, batch, a, b, c, ta, tb);
    case 32:
        if (tunable("gemm32s_variant") == 1) return launch_cfg<float, CfgS32b>(ctx, st, batch, a, b, c, ta, tb);
        if (tunable("gemm32s_variant") == 3) return launch_cfg<float, CfgS32t>(ctx, st, batch, a, b, c, ta, tb);
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
        const int r = launch_gemm_tiny_batched<double>(ctx, st, batch, a, b, c, m, n, k, ta, tb);
        if (r != SLAS_ERR_UNSUPPORTED) return r;
        return batch >= 64 ? launch_generic_small<double>(ctx, st, batch, a, b, c, m, n, k, ta, tb) : SLAS_ERR_UNSUPPORTED;
    }
    switch (m) {
    case 4: return launch_cfg<double, CfgD4>(ctx, st, batch, a, b, c, ta, tb);
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
