// gemm_generic.cu -- shape-agnostic MatrixMul kernels: the fallback for every product that has
// no shape-specialised kernel (odd sizes, leading dimensions wider than the matrix, tiny
// golden-test shapes), and matrix_vector_mul.
//
// Argument convention = the reference's call into cblas (src/backends/blas.rs:94-109, 136-149):
//   gemm: C[m x n] = op(A) . op(B), row-major, alpha 1, beta 0,
//         op(A)[i][p] = a_trans ? a[p*lda + i] : a[i*lda + p]
//         op(B)[p][j] = b_trans ? b[j*ldb + p] : b[p*ldb + j]
//   gemv: trait arguments m = underlying width, n = underlying height (src/tensor.rs:428-437);
//         y = op(A) . x with A stored n rows x m columns, lda = m.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"

namespace slas {

// One thread per output element; consecutive threads walk along a row of C so that B (and C)
// accesses coalesce.  Serves tiny shapes where launch latency dominates anyway.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_naive_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c, size_t total, unsigned m,
                  unsigned n, unsigned k, size_t lda, size_t ldb, size_t ldc, int ta, int tb, size_t sa, size_t sb,
                  size_t sc) {
    const size_t per = (size_t)m * n;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (size_t)gridDim.x * 256) {
        const size_t obj = t / per;
        const unsigned e = (unsigned)(t - obj * per);
        const unsigned i = e / n, j = e % n;
        const T *pa = a + obj * sa;
        const T *pb = b + obj * sb;
        T acc = T(0);
        for (unsigned p = 0; p < k; ++p) {
            const T av = ta ? pa[(size_t)p * lda + i] : pa[(size_t)i * lda + p];
            const T bv = tb ? pb[(size_t)j * ldb + p] : pb[(size_t)p * ldb + j];
            acc = fma(av, bv, acc);
        }
        c[obj * sc + (size_t)i * ldc + j] = acc;
    }
}

template <typename T>
int launch_gemm_generic(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *b, T *c, size_t m,
                        size_t n, size_t k, size_t lda, size_t ldb, size_t ldc, int ta, int tb, size_t stride_a,
                        size_t stride_b, size_t stride_c) {
    if (batch == 0 || m == 0 || n == 0) return SLAS_OK;
    SLAS_REQUIRE(m <= 0xffffffffu && n <= 0xffffffffu && k <= 0xffffffffu && m * n <= 0xffffffffu,
                 "gemm: dimension too large for the generic kernel");
    const size_t total = batch * m * n;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    gemm_naive_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a, b, c, total, (unsigned)m, (unsigned)n, (unsigned)k, lda,
                                                           ldb, ldc, ta, tb, stride_a, stride_b, stride_c);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}
template int launch_gemm_generic<float>(Context *, cudaStream_t, size_t, const float *, const float *, float *,
                                        size_t, size_t, size_t, size_t, size_t, size_t, int, int, size_t, size_t,
                                        size_t);
template int launch_gemm_generic<double>(Context *, cudaStream_t, size_t, const double *, const double *, double *,
                                         size_t, size_t, size_t, size_t, size_t, size_t, int, int, size_t, size_t,
                                         size_t);

// ---- gemv ------------------------------------------------------------------------------------
// not transposed: y[i] = sum_j a[i*lda + j] * x[j]   (i < n rows, j < m columns): a warp per row.
template <typename T>
__global__ void __launch_bounds__(256)
gemv_n_kernel(const T *__restrict__ a, const T *__restrict__ x, T *__restrict__ y, size_t total_rows, unsigned m,
              unsigned n, size_t lda, size_t sa, size_t sx, size_t sy) {
    const unsigned lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * 8;
    for (size_t row = warp; row < total_rows; row += nwarps) {
        const size_t obj = row / n;
        const unsigned i = (unsigned)(row - obj * n);
        const T *pa = a + obj * sa + (size_t)i * lda;
        const T *px = x + obj * sx;
        T acc = T(0);
        for (unsigned j = lane; j < m; j += 32) acc = fma(pa[j], px[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[obj * sy + i] = acc;
    }
}

// transposed: y[j] = sum_i a[i*lda + j] * x[i]   (j < m, i < n): 32 columns x 8 row slices per CTA.
template <typename T>
__global__ void __launch_bounds__(256)
gemv_t_kernel(const T *__restrict__ a, const T *__restrict__ x, T *__restrict__ y, unsigned m, unsigned n,
              size_t lda, size_t sa, size_t sx, size_t sy, unsigned col_tiles) {
    __shared__ T red[8][33];
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t obj = blockIdx.x / col_tiles;
    const unsigned j = (blockIdx.x % col_tiles) * 32 + tx;
    const T *pa = a + obj * sa;
    const T *px = x + obj * sx;
    T acc = T(0);
    if (j < m)
        for (unsigned i = ty; i < n; i += 8) acc = fma(pa[(size_t)i * lda + j], px[i], acc);
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && j < m) {
        T s = red[0][tx];
#pragma unroll
        for (int r = 1; r < 8; ++r) s += red[r][tx];
        y[obj * sy + j] = s;
    }
}

// Batched dense gemv for power-of-two shapes: the batch is one contiguous stream of rows, so a warp
// reads 512 consecutive bytes of A per load (LPR = m / VW lanes per row, R = 32 / LPR rows per load),
// multiplies in registers and reduces with shuffles -- over the lanes of a row for y = A x, over the
// lanes of a column for y = A^T x.  HBM traffic = A + x + y, every request fully coalesced.
template <typename T, bool TRANS>
__global__ void __launch_bounds__(256)
gemv_small_kernel(const T *__restrict__ a, const T *__restrict__ x, T *__restrict__ y, size_t batch, unsigned m, unsigned n) {
    constexpr int VW = 16 / (int)sizeof(T);
    struct alignas(16) V { T v[VW]; };
    const unsigned lane = threadIdx.x & 31;
    const unsigned lpr = m / VW, rpl = 32 / lpr;     // lanes per row, rows per warp-load
    const unsigned q = lane % lpr, rl = lane / lpr;  // column vector of this lane, row within the load
    const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 8;
    const size_t xlen = TRANS ? n : m, ylen = TRANS ? m : n;
    if (n >= rpl) {
        // one object per warp iteration, n / rpl loads
        for (size_t obj = warp; obj < batch; obj += nwarps) {
            const T *pa = a + obj * (size_t)m * n;
            const T *px = x + obj * xlen;
            T *py = y + obj * ylen;
            V xv, acc;
            if (!TRANS) xv = *reinterpret_cast<const V *>(px + q * VW);
#pragma unroll
            for (int j = 0; j < VW; ++j) acc.v[j] = T(0);
            for (unsigned r0 = 0; r0 < n; r0 += rpl) {
                const unsigned i = r0 + rl;
                const uint4 raw = __ldcs(reinterpret_cast<const uint4 *>(pa + (size_t)i * m + q * VW));
                const V av = *reinterpret_cast<const V *>(&raw);
                if (!TRANS) {
                    T p = T(0);
#pragma unroll
                    for (int j = 0; j < VW; ++j) p = fma(av.v[j], xv.v[j], p);
                    for (unsigned o = 1; o < lpr; o <<= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
                    if (q == 0) py[i] = p;
                } else {
                    const T xi = px[i];
#pragma unroll
                    for (int j = 0; j < VW; ++j) acc.v[j] = fma(av.v[j], xi, acc.v[j]);
                }
            }
            if (TRANS) {
                for (unsigned o = lpr; o < 32; o <<= 1)
#pragma unroll
                    for (int j = 0; j < VW; ++j) acc.v[j] += __shfl_xor_sync(0xffffffffu, acc.v[j], o);
                if (rl == 0) *reinterpret_cast<V *>(py + q * VW) = acc;
            }
        }
    } else {
        // several whole objects per warp-load
        const unsigned opw = rpl / n, ol = rl / n, i = rl % n;
        for (size_t obj0 = warp * opw; obj0 < batch; obj0 += nwarps * opw) {
            const size_t obj = obj0 + ol;
            const bool live = obj < batch;
            const size_t o = live ? obj : batch - 1;
            const uint4 raw = __ldcs(reinterpret_cast<const uint4 *>(a + o * (size_t)m * n + (size_t)i * m + q * VW));
            const V av = *reinterpret_cast<const V *>(&raw);
            if (!TRANS) {
                const V xv = *reinterpret_cast<const V *>(x + o * xlen + q * VW);
                T p = T(0);
#pragma unroll
                for (int j = 0; j < VW; ++j) p = fma(av.v[j], xv.v[j], p);
                for (unsigned s = 1; s < lpr; s <<= 1) p += __shfl_xor_sync(0xffffffffu, p, s);
                if (q == 0 && live) y[o * ylen + i] = p;
            } else {
                const T xi = x[o * xlen + i];
                V acc;
#pragma unroll
                for (int j = 0; j < VW; ++j) acc.v[j] = av.v[j] * xi;
                for (unsigned s = lpr; s < lpr * n; s <<= 1)
#pragma unroll
                    for (int j = 0; j < VW; ++j) acc.v[j] += __shfl_xor_sync(0xffffffffu, acc.v[j], s);
                if (i == 0 && live) *reinterpret_cast<V *>(y + o * ylen + q * VW) = acc;
            }
        }
    }
}

// Batched dense y = A x, rows of a multiple of 32 bytes: ONE LANE PER ROW.  The batch is one contiguous stream of
// rows and y one contiguous stream of results (y index == global row index), so a lane walks its row with 256-bit
// loads (LDG.E.256: every request is a whole 32-byte sector), multiplies by x -- shared by the rows of an object, read
// through L1 -- and the warp stores 32 consecutive results: no shuffles, no partial-sector stores (the lanes-per-row
// kernel above writes y in pieces of a few elements: 0.83-0.93 of HBM peak for 16x16 / 32x32).
template <typename T>
__global__ void __launch_bounds__(256)
gemv_rowlane_kernel(const T *__restrict__ a, const T *__restrict__ x, T *__restrict__ y, size_t rows_total, unsigned m, unsigned n) {
    constexpr int E32 = 32 / (int)sizeof(T);  // elements per 32-byte chunk
    const unsigned chunks = m / E32;
    for (size_t row = (size_t)blockIdx.x * 256 + threadIdx.x; row < rows_total; row += (size_t)gridDim.x * 256) {
        const T *pa = a + row * m;
        const T *px = x + (row / n) * m;
        T acc0 = T(0), acc1 = T(0);
#pragma unroll 4
        for (unsigned c = 0; c < chunks; ++c) {
            uint32_t w[8];
            asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                         : "l"(pa + c * E32));
            const T *av = reinterpret_cast<const T *>(w);
            const uint4 x0 = *reinterpret_cast<const uint4 *>(px + c * E32);
            const uint4 x1 = *reinterpret_cast<const uint4 *>(px + c * E32 + E32 / 2);
            const T *xa = reinterpret_cast<const T *>(&x0), *xb = reinterpret_cast<const T *>(&x1);
#pragma unroll
            for (int j = 0; j < E32 / 2; ++j) {
                acc0 = fma(av[j], xa[j], acc0);
                acc1 = fma(av[E32 / 2 + j], xb[j], acc1);
            }
        }
        y[row] = acc0 + acc1;
    }
}

// The same batched y = A x with FULLY COALESCED loads: VPR = row bytes / 16 consecutive lanes share a row (lane q reads
// 16-byte vector q), so a warp instruction reads 512 consecutive bytes of the row stream instead of one 32-byte sector
// from each of 32 rows; a butterfly over the VPR lanes adds the partial sums (fixed order) and one more shuffle per step
// hands row t of the warp's 32-row block to lane t, so y leaves as one coalesced store per block.  All VPR loads of a lane
// are issued before the first use.
// CH = 16-byte vectors per lane and row: a row of VPR * CH vectors is read as CH pieces of VPR consecutive vectors (256-byte
// rows: 8 lanes x 2 pieces of 128 bytes -- whole lines -- instead of 16 lanes per row, which leaves two rows per warp
// instruction and measured 0.74-0.80).
template <typename T, int VPR, int CH = 1>
__global__ void __launch_bounds__(256)
gemv_rowgroup_kernel(const T *__restrict__ a, const T *__restrict__ x, T *__restrict__ y, size_t rows_total, unsigned n) {
    constexpr int N = 16 / (int)sizeof(T), RPW = 32 / VPR, STEPS = 32 / RPW;  // rows per warp and step; steps per 32-row block
    constexpr int PASS = (STEPS * CH <= 8) ? STEPS : 8 / CH;                   // steps whose loads are in flight together
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    const int lane = threadIdx.x & 31, q = lane % VPR, g = lane / VPR;
    const size_t warps = (size_t)gridDim.x * 8, w0 = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const V *av = reinterpret_cast<const V *>(a);
    const V *xv = reinterpret_cast<const V *>(x);
    for (size_t base = w0 * 32; base < rows_total; base += warps * 32) {
        T mine = T(0);
#pragma unroll
        for (int s0 = 0; s0 < STEPS; s0 += PASS) {
            V ra[PASS][CH], rx[PASS][CH];
#pragma unroll
            for (int s = 0; s < PASS; ++s) {
                size_t row = base + (size_t)(s0 + s) * RPW + g;
                if (row >= rows_total) row = rows_total - 1;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    ra[s][c] = __ldcs(av + row * (VPR * CH) + c * VPR + q);
                    rx[s][c] = __ldg(xv + (row / n) * (VPR * CH) + c * VPR + q);
                }
            }
#pragma unroll
            for (int s = 0; s < PASS; ++s) {
                T acc = T(0);
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const T *pa = reinterpret_cast<const T *>(&ra[s][c]), *px = reinterpret_cast<const T *>(&rx[s][c]);
#pragma unroll
                    for (int j = 0; j < N; ++j) acc = fma(pa[j], px[j], acc);
                }
#pragma unroll
                for (int o = VPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, VPR);
                // row (s0 + s) * RPW + g' of the block goes to lane (s0 + s) * RPW + g': that lane reads group g' = lane % RPW
                const T v = __shfl_sync(0xffffffffu, acc, (lane % RPW) * VPR);
                if (lane / RPW == s0 + s) mine = v;
            }
        }
        if (base + lane < rows_total) y[base + lane] = mine;
    }
}

// One large matrix, y = A x: a warp per row, 128-bit loads, four independent accumulators, x through L1/L2.
template <typename T>
__global__ void __launch_bounds__(256)
gemv_n_vec_kernel(const T *__restrict__ a, const T *__restrict__ x, T *__restrict__ y, unsigned m, unsigned n, size_t lda) {
    constexpr int VW = 16 / (int)sizeof(T);
    struct alignas(16) V { T v[VW]; };
    const unsigned lane = threadIdx.x & 31;
    const unsigned nvec = m / VW;
    for (size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += (size_t)gridDim.x * 8) {
        const V *pa = reinterpret_cast<const V *>(a + row * lda);
        const V *px = reinterpret_cast<const V *>(x);
        T acc[4] = {T(0), T(0), T(0), T(0)};
        unsigned j = lane;
        for (; j + 96 < nvec; j += 128) {
            V av[4], xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint4 raw = __ldcs(reinterpret_cast<const uint4 *>(pa + j + 32 * u));
                av[u] = *reinterpret_cast<const V *>(&raw);
                xv[u] = px[j + 32 * u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int e = 0; e < VW; ++e) acc[u] = fma(av[u].v[e], xv[u].v[e], acc[u]);
        }
        for (; j < nvec; j += 32) {
            const V av = pa[j], xv = px[j];
#pragma unroll
            for (int e = 0; e < VW; ++e) acc[0] = fma(av.v[e], xv.v[e], acc[0]);
        }
        T s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        for (unsigned e = nvec * VW + lane; e < m; e += 32) s = fma(a[row * lda + e], x[e], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[row] = s;
    }
}

// One large matrix, y = A^T x: CTA = 32 column-vectors (128-bit each) x 8 row lanes; grid.y splits the rows,
// each split writes partial[split][m]; gemv_t_finish adds the splits in order (deterministic, no atomics).
template <typename T>
__global__ void __launch_bounds__(256)
gemv_t_split_kernel(const T *__restrict__ a, const T *__restrict__ x, T *__restrict__ partial, unsigned m, unsigned n,
                    size_t lda, unsigned rows_per_split) {
    constexpr int VW = 16 / (int)sizeof(T);
    struct alignas(16) V { T v[VW]; };
    __shared__ V red[8][32];
    const unsigned cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const unsigned col = (blockIdx.x * 32 + cl) * VW;
    const unsigned r0 = blockIdx.y * rows_per_split;
    const unsigned r1 = r0 + rows_per_split < n ? r0 + rows_per_split : n;
    V acc[2];
#pragma unroll
    for (int e = 0; e < VW; ++e) { acc[0].v[e] = T(0); acc[1].v[e] = T(0); }
    if (col < m) {
        unsigned i = r0 + rl;
        for (; i + 8 < r1; i += 16) {
            const uint4 w0 = __ldcs(reinterpret_cast<const uint4 *>(a + (size_t)i * lda + col));
            const uint4 w1 = __ldcs(reinterpret_cast<const uint4 *>(a + (size_t)(i + 8) * lda + col));
            const V a0 = *reinterpret_cast<const V *>(&w0), a1 = *reinterpret_cast<const V *>(&w1);
            const T x0 = x[i], x1 = x[i + 8];
#pragma unroll
            for (int e = 0; e < VW; ++e) { acc[0].v[e] = fma(a0.v[e], x0, acc[0].v[e]); acc[1].v[e] = fma(a1.v[e], x1, acc[1].v[e]); }
        }
        if (i < r1) {
            const V a0 = *reinterpret_cast<const V *>(a + (size_t)i * lda + col);
            const T x0 = x[i];
#pragma unroll
            for (int e = 0; e < VW; ++e) acc[0].v[e] = fma(a0.v[e], x0, acc[0].v[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < VW; ++e) acc[0].v[e] += acc[1].v[e];
    red[rl][cl] = acc[0];
    __syncthreads();
    if (rl == 0 && col < m) {
        V s = red[0][cl];
#pragma unroll
        for (int r = 1; r < 8; ++r)
#pragma unroll
            for (int e = 0; e < VW; ++e) s.v[e] += red[r][cl].v[e];
        *reinterpret_cast<V *>(partial + (size_t)blockIdx.y * m + col) = s;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) gemv_t_finish_kernel(const T *__restrict__ partial, T *__restrict__ y, unsigned m, unsigned splits) {
    const unsigned j = blockIdx.x * 256 + threadIdx.x;
    if (j >= m) return;
    T s = T(0);
    for (unsigned sp = 0; sp < splits; ++sp) s += partial[(size_t)sp * m + j];
    y[j] = s;
}

template <typename T>
int launch_gemv(Context *ctx, cudaStream_t st, size_t batch, const T *a, const T *x, T *y, size_t m, size_t n,
                size_t lda, int ta, size_t stride_a, size_t stride_x, size_t stride_y) {
    if (batch == 0) return SLAS_OK;
    SLAS_REQUIRE(m <= 0xffffffffu && n <= 0xffffffffu, "gemv: dimension too large");
    {
        // one large matrix with 16-byte aligned rows: the vectorised kernels
        constexpr size_t VW = 16 / sizeof(T);
        if (batch == 1 && m >= 32 * VW && n >= 64 && aligned16(a) && aligned16(x) && aligned16(y) && lda % VW == 0) {
            if (!ta) {
                size_t blocks = (n + 7) / 8;
                const size_t cap = (size_t)ctx->sm_count * 16;
                if (blocks > cap) blocks = cap;
                gemv_n_vec_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, (unsigned)m, (unsigned)n, lda);
                SLAS_LAUNCH_CHECK();
                return SLAS_OK;
            }
            if (m % VW == 0) {
                const size_t col_tiles = (m / VW + 31) / 32;
                size_t splits = ((size_t)ctx->sm_count * 4 + col_tiles - 1) / col_tiles;
                if (splits > (n + 63) / 64) splits = (n + 63) / 64;
                if (splits < 1) splits = 1;
                const size_t rows_per_split = (n + splits - 1) / splits;
                splits = (n + rows_per_split - 1) / rows_per_split;
                void *ws = nullptr;
                SLAS_TRY(ensure_workspace(ctx, splits * m * sizeof(T), &ws));
                gemv_t_split_kernel<T><<<dim3((unsigned)col_tiles, (unsigned)splits), 256, 0, st>>>(a, x, (T *)ws, (unsigned)m, (unsigned)n,
                                                                                                  lda, (unsigned)rows_per_split);
                SLAS_LAUNCH_CHECK();
                gemv_t_finish_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, st>>>((const T *)ws, y, (unsigned)m, (unsigned)splits);
                SLAS_LAUNCH_CHECK();
                return SLAS_OK;
            }
        }
    }
    {
        // batched dense power-of-two shapes: the coalesced streaming kernel
        constexpr size_t VW = 16 / sizeof(T);
        const bool pow2 = m && n && !(m & (m - 1)) && !(n & (n - 1));
        const size_t xlen = ta ? n : m, ylen = ta ? m : n;
        // rows of 32 / 64 / 128 / 256 bytes: lanes share a row, loads fully coalesced -- 16x16 / 32x32 / 64x64 f32 0.95 / 0.93 /
        // 0.89 -> 1.01 / 1.00 / 0.99 of HBM peak, 32x32 f64 0.93 -> 0.95 (profiles/r4a_tune_gemv.log, r4d_tune_gemv.log); 256-byte
        // rows as 8 lanes x 2 pieces of 128 bytes (16 lanes per row measured 0.74-0.80).  gemv_variant 3: the lane-per-row kernel
        if (batch > 1 && !ta && lda == m && n > 0 && stride_a == m * n && stride_x == m && stride_y == n && aligned16(a) && aligned16(x) &&
            tunable("gemv_variant") != 1 && tunable("gemv_variant") != 3) {
            const size_t rows_total = batch * n;
            size_t blocks = (rows_total + 255) / 256;
            const size_t cap = (size_t)ctx->sm_count * 16;
            if (blocks > cap) blocks = cap;
            bool done = true;
            switch (m * sizeof(T)) {
            case 32: gemv_rowgroup_kernel<T, 2><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, rows_total, (unsigned)n); break;
            case 64: gemv_rowgroup_kernel<T, 4><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, rows_total, (unsigned)n); break;
            case 128: gemv_rowgroup_kernel<T, 8><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, rows_total, (unsigned)n); break;
            case 256: gemv_rowgroup_kernel<T, 8, 2><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, rows_total, (unsigned)n); break;
            default: done = false;
            }
            if (done) {
                SLAS_LAUNCH_CHECK();
                return SLAS_OK;
            }
        }
        if (batch > 1 && !ta && lda == m && (m * sizeof(T)) % 32 == 0 && m * sizeof(T) <= 1024 && n > 0 && stride_a == m * n &&
            stride_x == m && stride_y == n && (uintptr_t)a % 32 == 0 && aligned16(x) && tunable("gemv_variant") != 1) {
            const size_t rows_total = batch * n;
            size_t blocks = (rows_total + 255) / 256;
            const size_t cap = (size_t)ctx->sm_count * 16;
            if (blocks > cap) blocks = cap;
            gemv_rowlane_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, rows_total, (unsigned)m, (unsigned)n);
            SLAS_LAUNCH_CHECK();
            return SLAS_OK;
        }
        if (batch > 1 && pow2 && lda == m && m >= VW && m <= 32 * VW && n <= 4096 && stride_a == m * n &&
            stride_x == xlen && stride_y == ylen && aligned16(a) && aligned16(x) && aligned16(y)) {
            size_t blocks = (batch + 7) / 8;
            const size_t cap = (size_t)ctx->sm_count * 8;
            if (blocks > cap) blocks = cap;
            if (!ta) gemv_small_kernel<T, false><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, batch, (unsigned)m, (unsigned)n);
            else gemv_small_kernel<T, true><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, batch, (unsigned)m, (unsigned)n);
            SLAS_LAUNCH_CHECK();
            return SLAS_OK;
        }
        // any other small dense shape (3x3 . 3-vector, 6x6, 12x12 ...): y = op(A) x is the batched product with a one-column
        // B, which the shape-specialised batched GEMM kernels run at HBM speed (a warp per row of a 3x3 matrix: 0.02)
        if (batch >= 64 && m > 0 && n > 0 && lda == m && stride_a == m * n && stride_x == xlen && stride_y == ylen &&
            m * n * sizeof(T) <= 44 * 1024 && aligned16(a) && aligned16(x) && aligned16(y) && tunable("gemv_variant") != 2) {
            // trait: A is n (height) x m (width); y = A x: M = n, K = m; y = A^T x: M = m, K = n with a_trans
            const int r = launch_gemm_small_batched<T>(ctx, st, batch, a, x, y, ta ? m : n, 1, ta ? n : m, ta, 0);
            if (r != SLAS_ERR_UNSUPPORTED) return r;
        }
    }
    if (!ta) {
        if (n == 0) return SLAS_OK;
        const size_t rows = batch * n;
        size_t blocks = (rows + 7) / 8;
        const size_t cap = (size_t)ctx->sm_count * 16;
        if (blocks > cap) blocks = cap;
        gemv_n_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, rows, (unsigned)m, (unsigned)n, lda, stride_a,
                                                           stride_x, stride_y);
    } else {
        if (m == 0) return SLAS_OK;
        const size_t col_tiles = (m + 31) / 32;
        const size_t blocks = batch * col_tiles;
        SLAS_REQUIRE(blocks <= 0x7fffffffu, "gemv: too many tiles");
        gemv_t_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a, x, y, (unsigned)m, (unsigned)n, lda, stride_a, stride_x,
                                                           stride_y, (unsigned)col_tiles);
    }
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}
template int launch_gemv<float>(Context *, cudaStream_t, size_t, const float *, const float *, float *, size_t,
                                size_t, size_t, int, size_t, size_t, size_t);
template int launch_gemv<double>(Context *, cudaStream_t, size_t, const double *, const double *, double *, size_t,
                                 size_t, size_t, int, size_t, size_t, size_t);

}  // namespace slas
