// gemm_tiny_kernel.cuh -- the device side of the thread-per-matrix batched GEMM (see gemm_tiny.cu for the design).
// Free of host headers: compiled by nvcc for every (m, n, k) in 1..4 (gemm_tiny.cu) and by NVRTC at run time for
// other shapes small enough to live in one thread's registers (gemm_jit.cu).
#pragma once
#include "async.cuh"

namespace slas {

// CNT contiguous elements between shared memory and registers with the widest vectors that divide the object
template <typename T, int CNT>
__device__ __forceinline__ void lds_linear(T (&dst)[CNT], const T *p) {
    constexpr int BYTES = CNT * (int)sizeof(T);
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(p)[i];
    } else if constexpr (BYTES % 8 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 8; ++i) reinterpret_cast<uint2 *>(dst)[i] = reinterpret_cast<const uint2 *>(p)[i];
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) dst[i] = p[i];
    }
}
template <typename T, int CNT>
__device__ __forceinline__ void sts_linear(T *p, const T (&src)[CNT]) {
    constexpr int BYTES = CNT * (int)sizeof(T);
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) reinterpret_cast<uint4 *>(p)[i] = reinterpret_cast<const uint4 *>(src)[i];
    } else if constexpr (BYTES % 8 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 8; ++i) reinterpret_cast<uint2 *>(p)[i] = reinterpret_cast<const uint2 *>(src)[i];
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) p[i] = src[i];
    }
}

template <typename T, int M, int N, int K, bool TA, bool TB>
__device__ __forceinline__ void tiny_product(const T (&la)[M * K], const T (&lb)[K * N], T (&lc)[M * N]) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            T acc = T(0);
#pragma unroll
            for (int p = 0; p < K; ++p) acc = fma(TA ? la[p * M + i] : la[i * K + p], TB ? lb[j * K + p] : lb[p * N + j], acc);
            lc[i * N + j] = acc;
        }
}

// G = matrices per chunk = consumer threads (a multiple of 32; 256 for the built-in shapes, fewer for run-time
// specialised shapes whose stages would not fit otherwise); one more warp produces.
template <typename T, int M, int N, int K, int G = 256>
__global__ void __launch_bounds__(G + 32)
gemm_tiny_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t batch, size_t nchunks, int ta,
                 int tb, unsigned stage_elems, unsigned cstage_elems) {
    constexpr int CONSUMERS = G, STAGES = 3, MK = M * K, KN = K * N, MN = M * N;
    extern __shared__ __align__(128) unsigned char smem[];
    T *stage_base = reinterpret_cast<T *>(smem);
    T *cstage_base = stage_base + (size_t)STAGES * stage_elems;
    uint64_t *full = reinterpret_cast<uint64_t *>(cstage_base + 2 * (size_t)cstage_elems);
    uint64_t *empty = full + STAGES;
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
                T *sb = sa + (size_t)G * MK;
                // ragged final chunk: whole 16-byte words (the arrays end on one: checked by the launcher)
                const uint32_t bytes_a = (cnt * MK * (uint32_t)sizeof(T) + 15u) & ~15u;
                const uint32_t bytes_b = (cnt * KN * (uint32_t)sizeof(T) + 15u) & ~15u;
                mbar_expect_tx(full + s, bytes_a + bytes_b);
                bulk_g2s(sa, a + first * MK, bytes_a, full + s);
                bulk_g2s(sb, b + first * KN, bytes_b, full + s);
            }
        }
        return;
    }
    uint32_t it = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        const size_t first = ch * G;
        const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
        const T *ma = stage_base + (size_t)s * stage_elems + (size_t)threadIdx.x * MK;
        const T *mb = stage_base + (size_t)s * stage_elems + (size_t)G * MK + (size_t)threadIdx.x * KN;
        T *cst = cstage_base + (size_t)(it & 1) * cstage_elems;
        mbar_wait(full + s, use & 1);
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");  // the bulk store that read this C stage is done
        if (threadIdx.x < cnt) {
            alignas(16) T la[MK];
            alignas(16) T lb[KN];
            alignas(16) T lc[MN];
            lds_linear<T, MK>(la, ma);
            lds_linear<T, KN>(lb, mb);
            if (!ta && !tb) tiny_product<T, M, N, K, false, false>(la, lb, lc);
            else if (ta && !tb) tiny_product<T, M, N, K, true, false>(la, lb, lc);
            else if (!ta && tb) tiny_product<T, M, N, K, false, true>(la, lb, lc);
            else tiny_product<T, M, N, K, true, true>(la, lb, lc);
            sts_linear<T, MN>(cst + (size_t)threadIdx.x * MN, lc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 2, %0;" ::"n"(CONSUMERS) : "memory");
        if (threadIdx.x == 0) {
            const size_t cbytes = (size_t)cnt * MN * sizeof(T);
            const uint32_t whole = (uint32_t)(cbytes & ~(size_t)15);
            if (whole)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(c + first * MN), "r"(smem_u32(cst)),
                             "r"(whole)
                             : "memory");
            // (a ragged last chunk may end inside a 16-byte word: those < 16 bytes go out with plain stores)
            for (size_t e = whole / sizeof(T); e < (size_t)cnt * MN; ++e) c[first * MN + e] = cst[e];
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- any other small shape: TM x TN register tiles, scalar shared-memory reads ----------------------------------
// For shapes that fit neither the vector-width rules of gemm_small_kernel (N or K not whole 16-byte vectors: 9x9,
// 10x10 f32, 5x9 ...) nor a single thread's registers.  TM | M, TN | N; TPM = (M/TM)*(N/TN) threads share a matrix
// (any count: MPP = 256 / TPM matrices per pass, the 256 % TPM left-over threads idle), G matrices per chunk, same
// bulk-TMA ring and bulk C store as above.  Per k a thread reads TM + TN scalars for TM * TN FMAs; a_trans / b_trans
// are run-time strides.  Only ever compiled by NVRTC (gemm_jit.cu), one specialisation per shape.
template <typename T, int M, int N, int K, int TM, int TN, int G>
__global__ void __launch_bounds__(288)
gemm_rect_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t batch, size_t nchunks, int ta,
                 int tb, unsigned stage_elems, unsigned cstage_elems) {
    constexpr int CONSUMERS = 256, STAGES = 3, MK = M * K, KN = K * N, MN = M * N;
    constexpr int TJ = N / TN, TPM = (M / TM) * TJ, MPP = CONSUMERS / TPM;
    static_assert(M % TM == 0 && N % TN == 0 && TPM <= CONSUMERS, "tile must divide the matrix");
    extern __shared__ __align__(128) unsigned char smem[];
    T *stage_base = reinterpret_cast<T *>(smem);
    T *cstage_base = stage_base + (size_t)STAGES * stage_elems;
    uint64_t *full = reinterpret_cast<uint64_t *>(cstage_base + 2 * (size_t)cstage_elems);
    uint64_t *empty = full + STAGES;
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
                T *sb = sa + (size_t)G * MK;
                const uint32_t bytes_a = (cnt * MK * (uint32_t)sizeof(T) + 15u) & ~15u;
                const uint32_t bytes_b = (cnt * KN * (uint32_t)sizeof(T) + 15u) & ~15u;
                mbar_expect_tx(full + s, bytes_a + bytes_b);
                bulk_g2s(sa, a + first * MK, bytes_a, full + s);
                bulk_g2s(sb, b + first * KN, bytes_b, full + s);
            }
        }
        return;
    }
    const int slot = threadIdx.x / TPM, tile = threadIdx.x % TPM;  // matrix within the pass, tile within the matrix
    const int ti = tile / TJ, tj = tile % TJ;
    const int sai = ta ? 1 : K, sap = ta ? M : 1;  // A(i, p) = ma[i * sai + p * sap]
    const int sbp = tb ? 1 : N, sbj = tb ? K : 1;  // B(p, j) = mb[p * sbp + j * sbj]
    uint32_t it = 0;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, use = it / STAGES;
        const size_t first = ch * G;
        const uint32_t cnt = (uint32_t)(batch - first < (size_t)G ? batch - first : G);
        const T *sa = stage_base + (size_t)s * stage_elems;
        const T *sb = sa + (size_t)G * MK;
        T *cst = cstage_base + (size_t)(it & 1) * cstage_elems;
        mbar_wait(full + s, use & 1);
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");  // the bulk store that read this C stage is done
        if (slot < MPP) {
#pragma unroll 1
            for (uint32_t mat = slot; mat < cnt; mat += MPP) {
                const T *ma = sa + (size_t)mat * MK + ti * TM * sai;
                const T *mb = sb + (size_t)mat * KN + tj * TN * sbj;
                T acc[TM][TN];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);
#pragma unroll
                for (int p = 0; p < K; ++p) {
                    T av[TM], bv[TN];
#pragma unroll
                    for (int i = 0; i < TM; ++i) av[i] = ma[i * sai + p * sap];
#pragma unroll
                    for (int j = 0; j < TN; ++j) bv[j] = mb[p * sbp + j * sbj];
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
                }
                T *mc = cst + (size_t)mat * MN + (ti * TM) * N + tj * TN;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) mc[i * N + j] = acc[i][j];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 2, %0;" ::"n"(CONSUMERS) : "memory");
        if (threadIdx.x == 0) {
            const size_t cbytes = (size_t)cnt * MN * sizeof(T);
            const uint32_t whole = (uint32_t)(cbytes & ~(size_t)15);
            if (whole)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(c + first * MN), "r"(smem_u32(cst)),
                             "r"(whole)
                             : "memory");
            for (size_t e = whole / sizeof(T); e < (size_t)cnt * MN; ++e) c[first * MN + e] = cst[e];
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace slas
