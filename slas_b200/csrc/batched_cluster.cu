// batched_cluster.cu -- batched norm / normalize of MEDIUM vectors (8 KB .. 32 KB and 128 KB .. 512 KB each): one CTA or
// one THREAD-BLOCK CLUSTER per vector, the vector held in registers between the norm and the division.
// Replaces a user's loop over Normalize::{norm, normalize} (src/backends/rust.rs:116-127) for such lengths.
//
// Why: normalize needs the whole vector twice.  The CTA-per-vector kernel (vector_ops.cu) reads it from HBM, reduces, then
// reads it AGAIN (an L2 hit at best) before it writes: 0.68-0.80 of HBM peak at 16 K .. 64 K f32 elements, and a CTA's
// load -> reduce -> re-load -> store phases do not overlap.  A CTA's registers hold 32 KB comfortably (256 threads x 8
// 16-byte vectors), a cluster of 2 / 4 / 8 CTAs holds 64 / 128 / 256 KB (512 KB with 16 vectors per thread): every
// element crosses HBM once in and once out (8 B / element f32), the per-CTA partial sums meet through distributed
// shared memory (st.shared::cluster into every peer's slot, one barrier.cluster per vector), and four resident CTAs
// per SM overlap one cluster's barrier with the others' loads and stores.
//
// norm and normalize share this kernel (MODE 1 / 2): normalize must divide by exactly the bits norm returns.
// Summation order (fixed => deterministic): a thread adds its vectors p = 0 .. PER-1 into accumulator p & 3 by FMA in T,
// the four accumulators are added in f64, block_sum (fixed butterfly) gives the CTA partial, the CL partials are added
// in rank order in f64, sqrt.rn in f64, rounded to T.
#include <atomic>

#include "common.cuh"
#include "divby.cuh"
#include "kernels.cuh"
#include "reduce.cuh"

namespace slas {

__device__ __forceinline__ uint32_t bc_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t bc_cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t bc_cluster_id() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t bc_nclusters() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}

// resident CTAs per SM asked of the compiler: norm keeps nothing (8); normalize keeps PER 16-byte vectors per thread
template <int MODE, int PER> struct ClusterOcc { static constexpr int kMinBlocks = MODE == 1 ? 8 : (PER <= 4 ? 5 : (PER <= 8 ? 3 : 2)); };

template <typename T, int MODE /* 1 norm, 2 normalize in place */, int PER>
__global__ void __launch_bounds__(256, ClusterOcc<MODE, PER>::kMinBlocks)
batched_norm_cluster_kernel(const T *__restrict__ a, T *__restrict__ out, T *a_mut, size_t batch, size_t len, unsigned chunk) {
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    __shared__ double red[8];
    __shared__ double part[2][8];  // [parity of the iteration][cluster rank]: written by the peers through DSMEM
    const uint32_t rank = bc_cluster_ctarank(), CL = bc_cluster_nctarank();
    const size_t nvec = len / N;
    const size_t lo = (size_t)rank * chunk;
    const size_t hi = lo + chunk < nvec ? lo + chunk : nvec;
    uint32_t it = 0;
    for (size_t v = bc_cluster_id(); v < batch; v += bc_nclusters(), ++it) {
        const V *av = reinterpret_cast<const V *>((MODE == 2 ? a_mut : a) + v * len);
        V x[PER];
        bool ok[PER];
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const size_t i = lo + threadIdx.x + (size_t)p * 256;
            ok[p] = i < hi;
            x[p] = ok[p] ? ld_stream(av + i) : vzero((V *)nullptr);
        }
        T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const T *px = reinterpret_cast<const T *>(&x[p]);
#pragma unroll
            for (int j = 0; j < N; ++j) acc[p & 3] = fma(px[j], px[j], acc[p & 3]);
        }
        double s = ((double)acc[0] + (double)acc[1]) + ((double)acc[2] + (double)acc[3]);
        s = block_sum<256>(s, red);  // valid in warp 0
        double total;
        if (CL > 1) {
            if (threadIdx.x < CL) {
                const uint32_t local = (uint32_t)__cvta_generic_to_shared(&part[it & 1][rank]);
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(threadIdx.x));
                asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(s) : "memory");
            }
            __syncwarp();
            // every CTA's partial is in every CTA's part[it & 1] after this barrier; the other parity is free again because
            // a CTA arrives here only after it has read the previous iteration's slots
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            total = 0.0;
            for (uint32_t r = 0; r < CL; ++r) total += part[it & 1][r];
        } else {
            if (threadIdx.x == 0) part[0][0] = s;
            __syncthreads();
            total = part[0][0];
            __syncthreads();
        }
        const T nrm = (T)sqrt(total);
        if (MODE == 1) {
            if (rank == 0 && threadIdx.x == 0) out[v] = nrm;
        } else {
            DivBy<T> by;
            by.init(nrm);
            V *mv = reinterpret_cast<V *>(a_mut + v * len);
#pragma unroll
            for (int p = 0; p < PER; ++p) {
                if (ok[p]) {
                    T *px = reinterpret_cast<T *>(&x[p]);
#pragma unroll
                    for (int j = 0; j < N; ++j) px[j] = by.div(px[j]);
                    st_stream(mv + lo + threadIdx.x + (size_t)p * 256, x[p]);
                }
            }
        }
    }
    // no CTA may exit while a peer can still write into its shared memory
    if (CL > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

template <typename T, int MODE, int PER>
static int launch_cluster_per(Context *ctx, cudaStream_t st, size_t batch, size_t len, const T *a, T *out, T *a_mut, unsigned cl) {
    constexpr int N = Vec16<T>::N;
    const size_t nvec = len / N;
    const unsigned chunk = (unsigned)((nvec + cl - 1) / cl);
    auto kern = batched_norm_cluster_kernel<T, MODE, PER>;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // as many clusters as can be resident at once (persistent: a cluster strides over the vectors)
    cfg.gridDim = dim3(cl);
    static std::atomic<int> resident_cache[9];  // per instantiation and cluster size (every device of a box is the same part)
    int resident = resident_cache[cl].load(std::memory_order_relaxed);
    if (resident == 0) {
        SLAS_CUDA_TRY(cudaOccupancyMaxActiveClusters(&resident, kern, &cfg));
        if (resident < 1) return fail(SLAS_ERR_CUDA, "no cluster of %u CTAs fits the device", cl);
        resident_cache[cl].store(resident, std::memory_order_relaxed);
    }
    const long cap = tunable("bnorm_cluster_ctas") > 0 ? tunable("bnorm_cluster_ctas") : ClusterOcc<MODE, PER>::kMinBlocks;  // CTAs per SM
    size_t clusters = (size_t)resident;
    const size_t by_cap = (size_t)ctx->sm_count * (size_t)cap / cl;
    if (clusters > by_cap) clusters = by_cap;
    if (clusters > batch) clusters = batch;
    if (clusters < 1) clusters = 1;
    cfg.gridDim = dim3((unsigned)(clusters * cl));
    SLAS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a, out, a_mut, batch, len, chunk));
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

// *handled = false: not a length / alignment for this kernel (the caller's other kernels take it)
template <typename T, int MODE>
int launch_batched_norm_cluster(Context *ctx, cudaStream_t st, size_t batch, size_t len, const T *a, T *out, T *a_mut, bool *handled) {
    *handled = false;
    const size_t bytes = len * sizeof(T);
    const T *base = MODE == 2 ? a_mut : a;
    if (tunable("bnorm_cluster") == 1 || batch == 0 || bytes % 16 || !aligned16(base) || bytes <= 8 * 1024 || bytes > 512 * 1024)
        return SLAS_OK;
    // Up to 32 KB: ONE CTA holds the vector (4 / 8 vectors per thread) -- 0.97-0.98 of HBM peak where the re-reading kernel
    // measured 0.69-0.80.  128 KB .. 512 KB: clusters of 5 .. 8 CTAs (8 vectors per thread up to 256 KB, 16 beyond; any
    // cluster size up to 8) -- 0.77-0.82 against 0.64-0.68.  In between the re-reading CTA-per-vector kernel still wins
    // (64 KB f32: 0.91 against 0.87-0.89, 128 KB f64: 0.82 against 0.77-0.80) and keeps the range.  Small clusters of fat
    // CTAs beat big clusters of thin ones (32 KB: 1 x 8 vectors 0.97, 2 x 4 vectors 0.91): the cluster barrier and the
    // co-scheduling of a cluster's CTAs cost more than the lost occupancy (profiles/r2z_tune_bnorm*.log).
    const size_t per4 = 256 * 4 * 16;  // bytes a CTA holds with 4 vectors per thread
    if (bytes > 2 * per4 && bytes <= 8 * per4) return SLAS_OK;
    *handled = true;
    if (bytes <= per4) return launch_cluster_per<T, MODE, 4>(ctx, st, batch, len, a, out, a_mut, 1);
    if (bytes <= 16 * per4) return launch_cluster_per<T, MODE, 8>(ctx, st, batch, len, a, out, a_mut, (unsigned)((bytes + 2 * per4 - 1) / (2 * per4)));
    return launch_cluster_per<T, MODE, 16>(ctx, st, batch, len, a, out, a_mut, (unsigned)((bytes + 4 * per4 - 1) / (4 * per4)));
}
template int launch_batched_norm_cluster<float, 1>(Context *, cudaStream_t, size_t, size_t, const float *, float *, float *, bool *);
template int launch_batched_norm_cluster<float, 2>(Context *, cudaStream_t, size_t, size_t, const float *, float *, float *, bool *);
template int launch_batched_norm_cluster<double, 1>(Context *, cudaStream_t, size_t, size_t, const double *, double *, double *, bool *);
template int launch_batched_norm_cluster<double, 2>(Context *, cudaStream_t, size_t, size_t, const double *, double *, double *, bool *);

}  // namespace slas
