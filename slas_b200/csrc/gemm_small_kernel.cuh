// gemm_small_kernel.cuh -- the device side of the batched static-shape GEMM (see gemm_small.cu for the design):
// register-tile helpers, the SmallCfg tile configuration and gemm_small_kernel.  Kept free of host headers so that the
// same text compiles under nvcc (gemm_small.cu, the ahead-of-time instantiations) and under NVRTC (gemm_jit.cu, which
// specialises it at run time for static shapes without an ahead-of-time instantiation -- the analogue of rustc
// monomorphising Matrix<T, B, M, K> for every shape a slas program uses).
#pragma once
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
                if (use > 0) mbar_wait_backoff(empty + s, (use - 1) & 1);
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

}  // namespace slas
