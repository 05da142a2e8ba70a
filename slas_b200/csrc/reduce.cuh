// reduce.cuh -- what the grid-wide reductions (reduce.cu: dot / norm / dotu; fused_chain.cu: the fused chains)
// share: 16-byte vector types, streaming loads, the per-thread accumulators, the launch geometry and the
// deterministic cross-block combine.  Both kernel families use the SAME geometry, accumulator layout and combine
// order for a given length, so a fused chain carries the same bits as the unfused sequence on this backend.
//
// Two geometries:
//   tiled     (long vectors) -- grid-stride tiles of THREADS x 4 vectors, grid capped at 8 CTAs per SM; the last
//             block to finish (self-resetting ticket) adds the per-block partials in index order.
//   balanced  (LEN <= ~2^22: BASELINE configs[0], where one call is a few microseconds of HBM time) -- exactly one
//             CTA wave, a multiple of the SM count, block c owns the contiguous vectors [c*per, (c+1)*per) so every
//             SM gets the same number of bytes; every load of a thread is issued before the first use; and the
//             cross-block step needs neither a fence nor an atomic: a block publishes its f64 partial as two
//             8-byte {half, flag} words (each store is single-copy atomic, so data and flag arrive together) and
//             block 0 polls the slots, adds them in index order and resets them for the next launch.
//             Saves the st -> MEMBAR.GPU -> ATOMG -> re-read chain of the ticket (~2 us of a 7.7 us kernel).
#pragma once
#include "common.cuh"
#include "p2p.cuh"

namespace slas {

template <typename T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; };

template <typename V> __device__ __forceinline__ V ld_stream(const V *p) { return __ldcs(p); }
template <typename V> __device__ __forceinline__ void st_stream(V *p, V v) { __stcs(p, v); }
__device__ __forceinline__ float4 vzero(float4 *) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ double2 vzero(double2 *) { return make_double2(0.0, 0.0); }

constexpr int kRedUnroll = 4;
constexpr int kFlushEvery = 16;  // outer iterations between flushes of the fp accumulators into f64
// Loads in flight per thread and operand: 4 x 16 bytes for the 256- / 512-thread CTAs; the 1024-thread CTA keeps 2 (the
// same bytes in flight per SM) so that reduce_kernel fits 32 registers per thread and TWO of them -- a kernel and its
// dependent-launch successor -- are resident on an SM at once.
template <int THREADS> struct RedTile { static constexpr int U = THREADS == 1024 ? 2 : kRedUnroll, MINB = THREADS == 1024 ? 2 : 1; };

// KIND: 0 = dot (sum a*b), 1 = squared norm (sum a*a, one stream), 2 = complex dotu (two sums over interleaved pairs)
template <typename T, int KIND>
struct Reducer {
    T acc[kRedUnroll];
    T acc_im[kRedUnroll];
    double sum = 0.0, sum_im = 0.0;
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int u = 0; u < kRedUnroll; ++u) { acc[u] = T(0); acc_im[u] = T(0); }
    }
    __device__ __forceinline__ void flush() {
        T s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        sum += (double)s;
        if (KIND == 2) {
            T t = (acc_im[0] + acc_im[1]) + (acc_im[2] + acc_im[3]);
            sum_im += (double)t;
        }
        clear();
    }
    // one 16-byte vector from each operand
    __device__ __forceinline__ void add(int u, float4 x, float4 y) {
        if (KIND == 2) {
            acc[u] = fmaf(x.x, y.x, acc[u]); acc[u] = fmaf(-x.y, y.y, acc[u]);
            acc[u] = fmaf(x.z, y.z, acc[u]); acc[u] = fmaf(-x.w, y.w, acc[u]);
            acc_im[u] = fmaf(x.x, y.y, acc_im[u]); acc_im[u] = fmaf(x.y, y.x, acc_im[u]);
            acc_im[u] = fmaf(x.z, y.w, acc_im[u]); acc_im[u] = fmaf(x.w, y.z, acc_im[u]);
        } else {
            acc[u] = fmaf(x.x, y.x, acc[u]); acc[u] = fmaf(x.y, y.y, acc[u]);
            acc[u] = fmaf(x.z, y.z, acc[u]); acc[u] = fmaf(x.w, y.w, acc[u]);
        }
    }
    __device__ __forceinline__ void add(int u, double2 x, double2 y) {
        if (KIND == 2) {
            acc[u] = fma(x.x, y.x, acc[u]); acc[u] = fma(-x.y, y.y, acc[u]);
            acc_im[u] = fma(x.x, y.y, acc_im[u]); acc_im[u] = fma(x.y, y.x, acc_im[u]);
        } else {
            acc[u] = fma(x.x, y.x, acc[u]); acc[u] = fma(x.y, y.y, acc[u]);
        }
    }
    __device__ __forceinline__ void add_scalar(T x, T y) { sum += (double)x * (double)y; }
};

// ---- launch geometry ---------------------------------------------------------------------------------------------
struct RedGeom {
    unsigned blocks = 1;
    unsigned threads = 256;       // 256 (tiled) or the balanced geometry's CTA size
    unsigned balanced = 0;        // 1: block c owns vectors [c*per, (c+1)*per)
    unsigned poll = 0;            // 1: partials travel as flagged words polled by block 0; 0: ticket + last block
    unsigned long long per = 0;   // vectors per block (balanced), a multiple of 8 (= 128 bytes)
};

// One partial in flight between a block and block 0: {low half | flag << 32}, {high half | flag << 32}.
struct LLSlot { unsigned long long w[2]; };
constexpr unsigned kLLFlag = 0x5A5A5A5Au;

struct RedScratch {
    double *partials;      // [2 * kMaxPartials]
    unsigned int *ticket;  // self-resetting
    LLSlot *ll;            // [3 * kMaxPartials], all zero between launches
    unsigned long long *phases;  // optional: %globaltimer stamps of block 0 (slas_cuda_set_tunable("reduce_phases", 1))
};

// `work` = number of 16-byte vectors (or scalars on the unaligned path).  nres = results per launch (<= 3).
RedGeom reduce_geometry(const Context *ctx, size_t work, bool vectorised);
RedScratch reduce_scratch(Context *ctx);

// ---- programmatic dependent launch (sm_90+) ------------------------------------------------------------------------
// The launch-bound vector kernels (one call = a few microseconds) are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: each lets its successor on the stream be scheduled as soon as
// all of its own blocks have executed pdl_trigger, so the successor's launch latency, and -- when the host-side hazard
// check (common.cuh) found no overlap with the operands of any kernel that may still be running -- its loads and FMAs
// overlap this kernel's cross-block tail.  pdl_wait blocks until the predecessor grid has completed and its writes are
// visible; every block executes it before touching shared scratch / dependent data and before it exits (so completion
// order stays the stream order).  Two kinds of kernel: "early" ones trigger first and wait late; the others wait
// first and trigger afterwards, which makes them barriers (nothing launched before them is still running when their
// successors start) -- that bounds the set of kernels an early one can overlap with to the host's hazard window.
// Without the launch attribute both instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
// Launch `kern` with the programmatic-serialization attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned blocks, unsigned threads, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long red_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void ll_store(LLSlot *slot, double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long f = (unsigned long long)kLLFlag << 32;
    // two independent 8-byte stores; each carries its own flag, so no ordering between them (or fence) is needed
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&slot->w[0]), "l"((bits & 0xffffffffull) | f) : "memory");
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&slot->w[1]), "l"((bits >> 32) | f) : "memory");
}
__device__ __forceinline__ double ll_wait(LLSlot *slot) {
    unsigned long long w0, w1;
    const unsigned long long t0 = red_timer_ns();
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0) : "l"(&slot->w[0]) : "memory");
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w1) : "l"(&slot->w[1]) : "memory");
        if ((unsigned)(w0 >> 32) == kLLFlag && (unsigned)(w1 >> 32) == kLLFlag) break;
        // a block that never arrives (a fault elsewhere in the grid) must not hang the GPU
        if ((++spins & 0xfff) == 0 && red_timer_ns() - t0 > 20000000000ull) __trap();
    }
    // reset for the next launch on this stream (kernel boundaries order it)
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&slot->w[0]), "l"(0ull) : "memory");
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&slot->w[1]), "l"(0ull) : "memory");
    return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}

// The range of 16-byte vectors this thread walks: i, i + THREADS, ... in groups of RedTile<THREADS>::U, up to `end`.
template <int THREADS>
__device__ __forceinline__ void red_range(const RedGeom &g, size_t nvec, size_t &i, size_t &end, size_t &stride) {
    if (g.balanced) {
        size_t lo = (size_t)blockIdx.x * g.per;
        if (lo > nvec) lo = nvec;
        end = lo + g.per < nvec ? lo + (size_t)g.per : nvec;
        i = lo + threadIdx.x;
        stride = (size_t)THREADS * RedTile<THREADS>::U;
    } else {
        i = (size_t)blockIdx.x * (THREADS * RedTile<THREADS>::U) + threadIdx.x;
        end = nvec;
        stride = (size_t)gridDim.x * (THREADS * RedTile<THREADS>::U);
    }
}

// Cross-block combine of NRES block sums (s[], valid in thread 0 of every block).  Returns true in every thread of
// the ONE block that holds the grid totals afterwards (p[], valid in thread 0); all other blocks get false and
// must return.  The order of additions depends only on the geometry: deterministic.
template <int THREADS, int NRES>
__device__ __forceinline__ bool grid_combine(const RedGeom &g, const RedScratch &sc, const double (&s)[NRES], double (&p)[NRES],
                                             double *red /* [THREADS / 32] */, bool *flag_smem) {
    const unsigned G = gridDim.x;
    if (G == 1) {
#pragma unroll
        for (int q = 0; q < NRES; ++q) p[q] = s[q];
        return true;
    }
    if (g.poll) {
        if (blockIdx.x != 0) {
            if (threadIdx.x == 0) {
#pragma unroll
                for (int q = 0; q < NRES; ++q) ll_store(sc.ll + (size_t)q * G + blockIdx.x, s[q]);
            }
            return false;
        }
#pragma unroll
        for (int q = 0; q < NRES; ++q) {
            double v = 0.0;
            for (unsigned j = threadIdx.x; j < G; j += THREADS) v += j == 0 ? s[q] : ll_wait(sc.ll + (size_t)q * G + j);
            p[q] = block_sum<THREADS>(v, red);
        }
        return true;
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NRES; ++q) sc.partials[(size_t)q * G + blockIdx.x] = s[q];
        __threadfence();
        const unsigned int t = atomicInc(sc.ticket, G - 1);  // wraps to 0 after the last block
        *flag_smem = (t == G - 1);
    }
    __syncthreads();
    if (!*flag_smem) return false;
    __threadfence();
#pragma unroll
    for (int q = 0; q < NRES; ++q) {
        double v = 0.0;
        for (unsigned j = threadIdx.x; j < G; j += THREADS) v += __ldcg(sc.partials + (size_t)q * G + j);
        p[q] = block_sum<THREADS>(v, red);
    }
    return true;
}
#endif

}  // namespace slas
