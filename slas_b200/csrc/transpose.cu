// transpose.cu -- Transpose::{transpose, transpose_inplace} (replaces src/backends/rust.rs:151-177)
// for single objects and for batches [[T; len]; batch].
//
// Reference semantics, per object:   rows = len / columns
//     buffer[columns*row + column] = a[rows*column + row]   column in [0,columns), row in [0,rows)
// i.e. the input is read as `columns` x `rows` row-major and written `rows` x `columns`
// row-major; `columns` is the column count of the OUTPUT.  Elements past rows*columns (when
// columns does not divide len) are never written, exactly like the reference's loop.
// Pure data movement on opaque elements => bit-exact for any Copy type: 1/2/4/8/16-byte elements on the vector / tiled /
// staged kernels, any other size (a 12-byte [f32; 3] ...) on the granule kernel at the end of the file.
//
// Kernels: (1) big matrices: 32x32 element tiles through padded shared memory, both sides
// coalesced; (2) many small objects: one warp per object, coalesced read into a padded
// per-warp shared tile, coalesced write of the transposed order; (3) square n x n with n a
// multiple of the 16-byte vector width: 128-bit loads, register micro-transpose, 128-bit
// stores, no shared memory.
#include <algorithm>

#include "async.cuh"
#include "common.cuh"
#include "jit.cuh"
#include "kernels.cuh"
#include "staged_objects_kernel.cuh"

namespace slas {

struct alignas(16) Elem16 { uint64_t lo, hi; };

// sm_100 has 256-bit global loads / stores (LDG.E.256 / STG.E.256)
__device__ __forceinline__ void ld256(const uint32_t *p, uint32_t *x) {
    asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                 : "l"(p));
}
__device__ __forceinline__ void st256(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                      uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
                 "r"(g), "r"(h)
                 : "memory");
}
// 16- or 32-byte vector of elements, streamed
template <typename V> __device__ __forceinline__ V ldvec(const V *p) {
    if constexpr (sizeof(V) == 16) {
        uint4 r = __ldcs(reinterpret_cast<const uint4 *>(p));
        return *reinterpret_cast<V *>(&r);
    } else {
        static_assert(sizeof(V) == 32, "vector of 16 or 32 bytes");
        union { uint32_t w[8]; V v; } u;
        ld256(reinterpret_cast<const uint32_t *>(p), u.w);
        return u.v;
    }
}
template <typename V> __device__ __forceinline__ void stvec(V *p, const V &y) {
    if constexpr (sizeof(V) == 16) {
        __stcs(reinterpret_cast<uint4 *>(p), *reinterpret_cast<const uint4 *>(&y));
    } else {
        static_assert(sizeof(V) == 32, "vector of 16 or 32 bytes");
        const uint32_t *w = reinterpret_cast<const uint32_t *>(&y);
        st256(reinterpret_cast<uint32_t *>(p), w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
    }
}

// ---- (1) tiled ---------------------------------------------------------------------------
template <typename E>
__global__ void __launch_bounds__(256)
transpose_tiled_kernel(const E *__restrict__ in, E *__restrict__ out, size_t len, unsigned rows, unsigned columns,
                       unsigned tiles_r, unsigned tiles_c) {
    __shared__ E tile[32][33];
    // blockIdx.x enumerates (object, tile_c, tile_r) with tile_r fastest
    size_t bid = blockIdx.x;
    const unsigned tr = (unsigned)(bid % tiles_r); bid /= tiles_r;
    const unsigned tc = (unsigned)(bid % tiles_c); bid /= tiles_c;
    const size_t obj = bid;
    const E *src = in + obj * len;
    E *dst = out + obj * len;
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    // input: `columns` rows of `rows` elements.  tile covers input rows [tc*32, +32), cols [tr*32, +32)
#pragma unroll
    for (unsigned j = 0; j < 32; j += 8) {
        const unsigned c = tc * 32 + ty + j, r = tr * 32 + tx;
        if (c < columns && r < rows) tile[ty + j][tx] = src[(size_t)c * rows + r];
    }
    __syncthreads();
#pragma unroll
    for (unsigned j = 0; j < 32; j += 8) {
        const unsigned r = tr * 32 + ty + j, c = tc * 32 + tx;
        if (r < rows && c < columns) dst[(size_t)r * columns + c] = tile[tx][ty + j];
    }
}

// ---- (1b) tiled, 4-byte elements, 128-bit global accesses: 64 x 64 tiles ----------------------------
// rows % 64 == 0, columns % 64 == 0, 16-byte aligned.  Each thread moves four 16-byte vectors in and four
// out; the element transpose happens through a padded (pitch 65) shared tile with scalar accesses.
__global__ void __launch_bounds__(256)
transpose_tiled64_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t len, unsigned rows,
                         unsigned columns, unsigned tiles_r, unsigned tiles_c) {
    __shared__ uint32_t tile[64][65];  // tile[out_row_in_tile][out_col_in_tile]
    size_t bid = blockIdx.x;
    const unsigned tr = (unsigned)(bid % tiles_r); bid /= tiles_r;
    const unsigned tc = (unsigned)(bid % tiles_c); bid /= tiles_c;
    const uint32_t *src = in + bid * len;
    uint32_t *dst = out + bid * len;
    // input: `columns` rows of `rows` elements; this tile = input rows [tc*64, +64), input cols [tr*64, +64)
    const unsigned vq = threadIdx.x & 15, rr = threadIdx.x >> 4;  // 16 vectors per row, 16 rows per pass
#pragma unroll
    for (unsigned p = 0; p < 4; ++p) {
        const unsigned ir = rr + 16 * p;  // input row within the tile = output column within the tile
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)(tc * 64 + ir) * rows + tr * 64 + vq * 4));
        tile[vq * 4 + 0][ir] = v.x;
        tile[vq * 4 + 1][ir] = v.y;
        tile[vq * 4 + 2][ir] = v.z;
        tile[vq * 4 + 3][ir] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (unsigned p = 0; p < 4; ++p) {
        const unsigned orow = rr + 16 * p;
        uint4 v;
        v.x = tile[orow][vq * 4 + 0];
        v.y = tile[orow][vq * 4 + 1];
        v.z = tile[orow][vq * 4 + 2];
        v.w = tile[orow][vq * 4 + 3];
        __stcs(reinterpret_cast<uint4 *>(dst + (size_t)(tr * 64 + orow) * columns + tc * 64 + vq * 4), v);
    }
}

// ---- (1c) tiled, 1- and 2-byte elements: 128-byte rows in, 32-byte sectors out ---------------------------------
// The element-granular kernels move one narrow element per thread and instruction (0.22 / 0.44 of HBM peak for
// 1- / 2-byte elements).  Here a tile is 128 x 128 elements: row segments of 128 * sizeof(E) bytes in and out.  Warps read whole input rows as 32-bit words into shared memory; then a lane owns one WORD COLUMN (EPW =
// 4 / sizeof(E) input columns = EPW output rows) of one sector's worth of input rows (32 / sizeof(E)): it reads that
// column word by word (consecutive lanes, consecutive words: conflict-free), transposes each EPW x EPW element
// block in registers with byte permutes (PRMT), and ends up holding one complete 32-byte sector for each of its EPW
// output rows, which leave with 256-bit stores.  Whole tiles only: rows % T == 0, columns % 128 == 0, 32-byte bases.
// With 2-byte elements a lane does this for two word columns (lane, lane + 32).  Measured: 0.66-0.73 (1-byte) and
// 0.55 (2-byte) of HBM peak at 8192^2 .. 16384^2, against 0.23 / 0.46 before; loads stall on DRAM latency (ncu: long
// scoreboard) -- narrow row segments at large strides, not instruction issue, are what is left.
template <typename E>
__global__ void __launch_bounds__(256)
transpose_tiled_narrow_kernel(const E *__restrict__ in, E *__restrict__ out, size_t len, unsigned rows, unsigned columns,
                              unsigned tiles_r, unsigned tiles_c) {
    constexpr unsigned WPL = sizeof(E);                // word columns per lane (measured best: 1-byte 1, 2-byte 2)
    constexpr unsigned T = WPL * 128 / sizeof(E);      // tile width in elements (input columns)
    constexpr unsigned TRI = 128;                      // tile height (input rows): output row segments of 128 * sizeof(E) bytes
    constexpr unsigned RPS = 32 / sizeof(E);           // input rows per 32-byte output sector
    constexpr unsigned NW = TRI / RPS;                 // warps = sectors per output row segment (4 / 8)
    constexpr unsigned EPW = 4 / sizeof(E);            // elements per 32-bit word
    __shared__ uint32_t sm[TRI][32 * WPL];
    size_t bid = blockIdx.x;
    const unsigned tr = (unsigned)(bid % tiles_r); bid /= tiles_r;
    const unsigned tc = (unsigned)(bid % tiles_c); bid /= tiles_c;
    const E *src = in + bid * len;
    E *dst = out + bid * len;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // NW warps = the sectors of an output row segment
    // input: `columns` rows of `rows` elements; this tile = input rows [tc*TRI, +TRI), input columns [tr*T, +T)
    for (unsigned r = warp; r < TRI; r += NW) {
        const uint32_t *row = reinterpret_cast<const uint32_t *>(src + (size_t)(tc * TRI + r) * rows + tr * T);
#pragma unroll
        for (unsigned h = 0; h < WPL; ++h) sm[r][lane + 32 * h] = __ldcs(row + lane + 32 * h);
    }
    __syncthreads();
#pragma unroll
    for (unsigned h = 0; h < WPL; ++h) {
        const unsigned wc = lane + 32 * h;  // word column = output rows wc * EPW + k
        uint32_t o[EPW][8];                 // o[k] = the sector of output row wc * EPW + k
#pragma unroll
        for (unsigned q = 0; q < 8; ++q) {
            const unsigned r0 = warp * RPS + q * EPW;  // first of the EPW input rows that make output word q
            if constexpr (sizeof(E) == 1) {
                const uint32_t w0 = sm[r0][wc], w1 = sm[r0 + 1][wc], w2 = sm[r0 + 2][wc], w3 = sm[r0 + 3][wc];
                const uint32_t a = __byte_perm(w0, w1, 0x5140), b = __byte_perm(w0, w1, 0x7362);
                const uint32_t c = __byte_perm(w2, w3, 0x5140), d = __byte_perm(w2, w3, 0x7362);
                o[0][q] = __byte_perm(a, c, 0x5410); o[1][q] = __byte_perm(a, c, 0x7632);
                o[2][q] = __byte_perm(b, d, 0x5410); o[3][q] = __byte_perm(b, d, 0x7632);
            } else {
                const uint32_t w0 = sm[r0][wc], w1 = sm[r0 + 1][wc];
                o[0][q] = __byte_perm(w0, w1, 0x5410); o[1][q] = __byte_perm(w0, w1, 0x7632);
            }
        }
#pragma unroll
        for (unsigned k = 0; k < EPW; ++k)
            st256(reinterpret_cast<uint32_t *>(dst + (size_t)(tr * T + wc * EPW + k) * columns + tc * TRI + warp * RPS), o[k][0],
                  o[k][1], o[k][2], o[k][3], o[k][4], o[k][5], o[k][6], o[k][7]);
    }
}

// ---- (2) warp per small object ---------------------------------------------------------------
// INPLACE: in == out is allowed (the warp holds the whole object in shared memory before it writes), and
// the elements past rows*columns are zeroed like the reference's zero-initialised temporary (rust.rs:157-159).
template <typename E, bool INPLACE>
__global__ void __launch_bounds__(256)
transpose_small_kernel(const E *in, E *out, size_t batch, unsigned len, unsigned rows, unsigned columns,
                       unsigned warp_elems /* smem elements per warp */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *smem = reinterpret_cast<E *>(smem_raw) + (size_t)(threadIdx.x >> 5) * warp_elems;
    const unsigned lane = threadIdx.x & 31;
    const unsigned pitch = rows + 1;
    const unsigned live = rows * columns;
    const size_t warp_global = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * 8;
    // per-step increments of (quotient, remainder) when the linear index advances by 32
    const unsigned qi = 32 / rows, ri = 32 % rows;
    const unsigned qo = 32 / columns, ro = 32 % columns;
    for (size_t obj = warp_global; obj < batch; obj += nwarps) {
        const E *src = in + obj * len;
        E *dst = out + obj * len;
        unsigned c = lane / rows, r = lane % rows;
        for (unsigned e = lane; e < live; e += 32) {
            smem[c * pitch + r] = src[e];
            c += qi; r += ri;
            if (r >= rows) { r -= rows; ++c; }
        }
        __syncwarp();
        unsigned orow = lane / columns, ocol = lane % columns;
        for (unsigned o = lane; o < live; o += 32) {
            dst[o] = smem[ocol * pitch + orow];
            orow += qo; ocol += ro;
            if (ocol >= columns) { ocol -= columns; ++orow; }
        }
        if (INPLACE)
            for (unsigned e = live + lane; e < len; e += 32) dst[e] = E{};
        __syncwarp();
    }
}

// ---- (2b) many small objects of any shape: bulk-TMA staging, transposed in shared memory --------------------------------
// The warp-per-object kernel above leaves most of a warp idle on objects of a few elements (3x3 f32: 0.07 of HBM
// peak).  transpose_staged_kernel (staged_objects_kernel.cuh): HBM only ever sees 1-D bulk copies, the transposition
// happens between two shared-memory tiles.  <E, 0, 0, 0, false> takes the shape as run-time arguments; for small objects
// of 8- and 16-byte elements the shape is a template constant and a thread moves a whole object (3x3 built in, other
// shapes specialised by NVRTC at run time, staged_jit.cu) -- see the launcher for the measurements behind the split.
template <typename E>
using StagedTransposeKernel = void (*)(const E *, E *, size_t, size_t, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned);
template <typename E>
static StagedTransposeKernel<E> staged_transpose_builtin(size_t rows, size_t columns) {
    if constexpr (sizeof(E) == 4 || sizeof(E) == 8) {
        if (rows == 3 && columns == 3) return transpose_staged_kernel<E, 9, 3, 3, true>;
    }
    return nullptr;
}
static unsigned gcd_u(unsigned a, unsigned b) { return b ? gcd_u(b, a % b) : a; }

// launcher shared by the out-of-place and in-place paths; *handled = false when the shape does not qualify
template <typename E>
static int launch_transpose_staged(Context *ctx, cudaStream_t st, size_t batch, size_t len, const E *a, E *buffer, size_t rows,
                                   size_t columns, bool *handled) {
    *handled = false;
    const size_t obj_bytes = len * sizeof(E);
    // objects up to 48 KB (100x100 f32: 0.59 -> 0.88 of HBM peak, in place 0.35 -> 0.89) -- unless the gather stride (one
    // input row, `rows` elements) is a multiple of half the shared-memory banks AND a warp spans many input rows
    // (columns >= 16): the contiguous stage cannot be padded, and 96x96 f32 measured 0.29 that way against 0.82 for the
    // padded tiles of the other kernels (with few columns the same stride is harmless: 16x3 0.93, 48x5 f64 1.00)
    const size_t limit = tunable("transpose_staged_limit") > 0 ? (size_t)tunable("transpose_staged_limit") : 48 * 1024;
    // ... objects of up to 16 KB take the staged kernel anyway, with a padded stage filled row by row (64x40 f32: 0.47 on the
    // warp-per-object kernel)
    const size_t stride_words = rows * sizeof(E) / 4;
    const bool one_bank = (rows * sizeof(E)) % 4 == 0 && stride_words % 16 == 0 && columns >= 16 && tunable("transpose_variant") != 12;
    const bool padded_stage = one_bank && obj_bytes <= 16 * 1024 && tunable("transpose_variant") != 14;
    if (one_bank && !padded_stage) return SLAS_OK;
    if (len != rows * columns || obj_bytes > limit || batch < 256 || !aligned16(a) || !aligned16(buffer) ||
        tunable("transpose_variant") == 11)
        return SLAS_OK;
    // G objects per chunk: ~16 KB, a multiple of 16 / gcd(object bytes, 16) so that every chunk starts on a 16-byte boundary
    size_t galign = 16;
    while (galign > 1 && obj_bytes % (16 / (galign / 2)) == 0) galign /= 2;  // 16 / gcd(obj_bytes, 16)
    // the kernel moves whole 16-byte words: a batch that does not end on one (1001 3x3 f32 objects ...) leaves its last
    // < galign objects to the warp-per-object kernel (which also works in place) instead of falling off the fast path
    const size_t rem = batch % galign;
    if (rem) {
        const size_t main = batch - rem, padded = columns * (rows + 1);
        if (main < 256 || padded * sizeof(E) > 12 * 1024) return SLAS_OK;
        SLAS_TRY(launch_transpose_staged<E>(ctx, st, main, len, a, buffer, rows, columns, handled));
        if (!*handled) return SLAS_OK;
        const size_t smem = 8 * padded * sizeof(E);
        const bool inplace = a == buffer;
        if (smem > 48 * 1024) {
            if (inplace) SLAS_CUDA_TRY(cudaFuncSetAttribute(transpose_small_kernel<E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            else SLAS_CUDA_TRY(cudaFuncSetAttribute(transpose_small_kernel<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        }
        const unsigned blocks = (unsigned)((rem + 7) / 8);
        if (inplace)
            transpose_small_kernel<E, true><<<blocks, 256, smem, st>>>(a + main * len, buffer + main * len, rem, (unsigned)len, (unsigned)rows,
                                                                       (unsigned)columns, (unsigned)padded);
        else
            transpose_small_kernel<E, false><<<blocks, 256, smem, st>>>(a + main * len, buffer + main * len, rem, (unsigned)len, (unsigned)rows,
                                                                        (unsigned)columns, (unsigned)padded);
        SLAS_LAUNCH_CHECK();
        return SLAS_OK;
    }
    // A thread per object (shape a compile-time constant: built in for 3x3, else specialised by NVRTC) when the object fits a
    // thread's registers and consecutive objects start in different shared-memory banks: 4-byte elements up to 120 bytes
    // per object (word stride sharing at most a factor 2 with the 32 banks), 8- / 16-byte elements up to 256 bytes (element
    // count sharing at most a factor 2 with the 16 / 8 accesses of a shared-memory phase).  Objects of up to 120 bytes then
    // take 4 KB stages, four CTAs per SM -- measured, fraction of HBM peak, run-time walk at 16 KB -> thread per object at
    // 4 KB: 2x3 / 3x3 / 3x5 / 3x7 f32 0.93 / 0.94 / 0.94 / 0.95 -> 0.98 / 0.97 / 0.99 / 0.97, 2x3 / 3x3 / 5x3 f64 0.88 / 0.85 /
    // 0.84 -> 0.99 / 0.98 / 1.00; the optimum is sharp (3 or 5 KB: 0.86-0.92) and the same wherever the output array lies
    // (profiles/r3c_tune_trsmall.log, r3c_tune_tr3place.log).  Bigger objects leave too few per 4 KB chunk (7x7 f32 0.60) and
    // keep 16 KB stages: per object for 8-byte elements (5x5 f64 0.93), the run-time walk for 4-byte ones (7x7 0.94, 9x9
    // 0.95, where neither a thread per object nor the walk with constant indices beat it, profiles/r2t_tune_odd.log).
    // tunable staged_jit: 1 = no NVRTC, 2 = run-time shape only.
    const long jit_mode = tunable("staged_jit");
    const bool banks_ok = sizeof(E) == 4 ? gcd_u((unsigned)len, 32) <= 2 : gcd_u((unsigned)len, 128 / (unsigned)sizeof(E)) <= 2;
    bool per_object = jit_mode != 2 && sizeof(E) >= 4 && banks_ok && obj_bytes <= (sizeof(E) == 4 ? 120 : 256) && !padded_stage;
    const size_t pitch = padded_stage ? rows + 16 / sizeof(E) : rows;
    // the specialised kernel first: without one (no NVRTC, compile failure) the run-time walk keeps its own stage size
    StagedTransposeKernel<E> kern = per_object ? staged_transpose_builtin<E>(rows, columns) : nullptr;
    CUfunction jit_fn = nullptr;
    if (!kern && per_object && jit_mode == 0) jit_fn = staged_transpose_jit(ctx, sizeof(E), (unsigned)len, (unsigned)rows, (unsigned)columns, true);
    if (!kern && !jit_fn) per_object = false;
    const size_t stage_kb = tunable("transpose_stage_kb") > 0 ? (size_t)tunable("transpose_stage_kb") : (per_object && obj_bytes <= 120 ? 4 : 16);
    size_t G = (stage_kb * 1024 / obj_bytes) / galign * galign;
    if (G < galign) G = galign;
    const size_t stage_bytes = (G * columns * pitch * sizeof(E) + 127) / 128 * 128;  // (== G * obj_bytes without padding)
    const size_t smem = 5 * stage_bytes + 6 * sizeof(uint64_t);
    if (smem > 200 * 1024) return SLAS_OK;
    const size_t nchunks = (batch + G - 1) / G;
    int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
    const int cta_cap = tunable("staged_cta_cap") > 0 ? (int)tunable("staged_cta_cap") : 4;
    if (per_sm > cta_cap) per_sm = cta_cap;
    if (per_sm < 1) per_sm = 1;
    size_t grid = (size_t)ctx->sm_count * per_sm;
    if (grid > nchunks) grid = nchunks;
    if (jit_fn) {
        size_t batch_arg = batch, nchunks_arg = nchunks;
        unsigned len_arg = (unsigned)len, rows_arg = (unsigned)rows, cols_arg = (unsigned)columns, g_arg = (unsigned)G,
                 stage_arg = (unsigned)stage_bytes, pitch_arg = (unsigned)pitch;
        void *params[] = {(void *)&a,        (void *)&buffer,   (void *)&batch_arg, (void *)&nchunks_arg, (void *)&len_arg,
                          (void *)&rows_arg, (void *)&cols_arg, (void *)&g_arg,     (void *)&stage_arg,   (void *)&pitch_arg};
        SLAS_TRY(jit_launch(jit_fn, (unsigned)grid, 288, smem, st, params));
        *handled = true;
        return SLAS_OK;
    }
    if (!kern) kern = padded_stage ? transpose_staged_kernel<E, 0, 0, 0, false, true> : transpose_staged_kernel<E, 0, 0, 0, false>;
    SLAS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, 288, smem, st>>>(a, buffer, batch, nchunks, (unsigned)len, (unsigned)rows, (unsigned)columns, (unsigned)G,
                                            (unsigned)stage_bytes, (unsigned)pitch);
    SLAS_LAUNCH_CHECK();
    *handled = true;
    return SLAS_OK;
}

// ---- (3) square n x n, 128-bit both sides, register micro-transpose ---------------------------
// VW = elements per 16-byte vector (4 for 4-byte, 2 for 8-byte elements).  One thread owns a
// VW x VW block: loads VW vectors from VW consecutive input rows, transposes in registers,
// stores VW vectors to VW consecutive output rows.  Threads are ordered with the input
// column block fastest, so loads are contiguous along input rows and the stores of the
// threads sharing an output row land in the same instruction (full 32-byte sectors).
// __ldcs overloads do not exist for arbitrary structs: route through uint4
template <typename V> __device__ __forceinline__ V ld16(const V *p) {
    uint4 r = __ldcs(reinterpret_cast<const uint4 *>(p));
    return *reinterpret_cast<V *>(&r);
}

// One thread owns a BR x VW block of the input (BR = MULT * VW rows, one 16-byte vector per row): BR
// 128-bit loads, a register transpose, then VW output rows of BR contiguous elements each (MULT
// 128-bit stores per row).  OUT_MAJOR orders the threads of an object with the input ROW block
// fastest, so that consecutive lanes store consecutive pieces of the same output row (full 128-byte
// lines per instruction); otherwise the input COLUMN block is fastest (contiguous loads).
template <typename E, int VW, int MULT, bool OUT_MAJOR>
__global__ void __launch_bounds__(256)
transpose_square_vec_kernel(const E *in, E *out, size_t nblocks_total, unsigned n) {
    struct alignas(VW * sizeof(E)) V { E v[VW]; };
    constexpr int BR = MULT * VW;
    const unsigned row_blocks = n / BR, col_blocks = n / VW;
    const unsigned blocks_per_obj = row_blocks * col_blocks;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < nblocks_total; t += (size_t)gridDim.x * 256) {
        const size_t obj = t / blocks_per_obj;
        const unsigned w = (unsigned)(t - obj * blocks_per_obj);
        const unsigned bi = OUT_MAJOR ? w % row_blocks : w / col_blocks;  // input row block
        const unsigned bj = OUT_MAJOR ? w / row_blocks : w % col_blocks;  // input column block
        const E *src = in + obj * (size_t)n * n + (size_t)(bi * BR) * n + bj * VW;
        E *dst = out + obj * (size_t)n * n + (size_t)(bj * VW) * n + bi * BR;
        V x[BR];
#pragma unroll
        for (int s = 0; s < BR; ++s) x[s] = ldvec(reinterpret_cast<const V *>(src + (size_t)s * n));
#pragma unroll
        for (int s = 0; s < VW; ++s) {
#pragma unroll
            for (int m = 0; m < MULT; ++m) {
                V y;
#pragma unroll
                for (int q = 0; q < VW; ++q) y.v[q] = x[m * VW + q].v[s];
                stvec(reinterpret_cast<V *>(dst + (size_t)s * n + m * VW), y);
            }
        }
    }
}

// ---- (3b) 64-byte objects (4x4 of 4-byte, 2x2 of 16-byte ... here: 4x4 b32), 256-bit accesses ---------
// sm_100 has 256-bit global loads / stores (LDG.E.256 / STG.E.256): one thread moves a whole 4x4 matrix
// with two of each, and every store instruction covers full 32-byte sectors (with 128-bit stores each
// instruction writes half of every sector it touches).
__global__ void __launch_bounds__(256)
transpose_4x4_b32_kernel(const uint32_t *in, uint32_t *out, size_t batch) {
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < batch; t += (size_t)gridDim.x * 256) {
        uint32_t x[16];
        ld256(in + t * 16, x);
        ld256(in + t * 16 + 8, x + 8);
        st256(out + t * 16, x[0], x[4], x[8], x[12], x[1], x[5], x[9], x[13]);
        st256(out + t * 16 + 8, x[2], x[6], x[10], x[14], x[3], x[7], x[11], x[15]);
    }
}

// n == VW (4x4 of 4-byte, 2x2 of 8-byte elements): one LANE per matrix ROW, so a warp reads and writes
// 512 consecutive bytes per instruction; the VW x VW transpose runs across the VW lanes of a matrix with
// VW shuffles.  Round r: lane q sends its element (q - r) mod VW to lane (q - r) mod VW ... i.e. it
// receives element q of lane (q + r) mod VW, which is exactly output element (q + r) mod VW of row q.
template <typename E, int VW>
__global__ void __launch_bounds__(256)
transpose_tiny_kernel(const E *__restrict__ in, E *__restrict__ out, size_t nrows_total) {
    struct alignas(16) V { E v[VW]; };
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < ((nrows_total + 31) & ~(size_t)31); t += (size_t)gridDim.x * 256) {
        const bool live = t < nrows_total;
        const unsigned q = (unsigned)(t % VW);
        V x;
        if (live) x = ld16(reinterpret_cast<const V *>(in) + t);
        V y;
#pragma unroll
        for (int r = 0; r < VW; ++r) {
            // element this lane contributes in round r: index (q - r) mod VW
            E send = x.v[0];
#pragma unroll
            for (int j = 1; j < VW; ++j)
                if (((q + VW - r) % VW) == (unsigned)j) send = x.v[j];
            const unsigned src_lane = (threadIdx.x & ~(VW - 1)) + ((q + r) % VW);
            const E got = __shfl_sync(0xffffffffu, send, src_lane & 31);
#pragma unroll
            for (int j = 0; j < VW; ++j)
                if (((q + r) % VW) == (unsigned)j) y.v[j] = got;
        }
        if (live) __stcs(reinterpret_cast<uint4 *>(out) + t, *reinterpret_cast<uint4 *>(&y));
    }
}


// ---- (4) in place, square n x n: tile pairs across the diagonal exchanged through shared memory --------
// Block (ti <= tj) loads tile (ti, tj) and its mirror (tj, ti), and writes each one transposed into the
// other's place: every element is read once and written once (2 * sizeof(E) bytes of traffic per element,
// against 5 * sizeof(E) for zeroed temporary + transpose + copy back).
template <typename E>
__global__ void __launch_bounds__(256)
transpose_inplace_square_kernel(E *data, size_t batch, unsigned n) {
    __shared__ E t0[32][33], t1[32][33];
    const unsigned ti = blockIdx.y, tj = blockIdx.x;
    if (ti > tj) return;
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (size_t o = blockIdx.z; o < batch; o += gridDim.z) {
        E *obj = data + o * (size_t)n * n;
#pragma unroll
        for (unsigned j = 0; j < 32; j += 8) {
            const unsigned r0 = ti * 32 + ty + j, c0 = tj * 32 + tx;
            if (r0 < n && c0 < n) t0[ty + j][tx] = obj[(size_t)r0 * n + c0];
            const unsigned r1 = tj * 32 + ty + j, c1 = ti * 32 + tx;
            if (ti != tj && r1 < n && c1 < n) t1[ty + j][tx] = obj[(size_t)r1 * n + c1];
        }
        __syncthreads();
#pragma unroll
        for (unsigned j = 0; j < 32; j += 8) {
            const unsigned r1 = tj * 32 + ty + j, c1 = ti * 32 + tx;  // mirror position receives t0 transposed
            if (r1 < n && c1 < n) obj[(size_t)r1 * n + c1] = t0[tx][ty + j];
            const unsigned r0 = ti * 32 + ty + j, c0 = tj * 32 + tx;
            if (ti != tj && r0 < n && c0 < n) obj[(size_t)r0 * n + c0] = t1[tx][ty + j];
        }
        __syncthreads();
    }
}

// 4-byte elements, n % 64 == 0, 16-byte aligned: the same exchange on 64 x 64 tiles with 128-bit accesses
__global__ void __launch_bounds__(256)
transpose_inplace_square64_kernel(uint32_t *data, size_t batch, unsigned n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t (*t0)[65] = reinterpret_cast<uint32_t (*)[65]>(smem_raw);
    uint32_t (*t1)[65] = t0 + 64;
    const unsigned ti = blockIdx.y, tj = blockIdx.x;
    if (ti > tj) return;
    const unsigned vq = threadIdx.x & 15, rr = threadIdx.x >> 4;
    for (size_t o = blockIdx.z; o < batch; o += gridDim.z) {
        uint32_t *obj = data + o * (size_t)n * n;
#pragma unroll
        for (unsigned p = 0; p < 4; ++p) {
            const unsigned ir = rr + 16 * p;
            const uint4 v = *reinterpret_cast<const uint4 *>(obj + (size_t)(ti * 64 + ir) * n + tj * 64 + vq * 4);
            t0[vq * 4 + 0][ir] = v.x; t0[vq * 4 + 1][ir] = v.y; t0[vq * 4 + 2][ir] = v.z; t0[vq * 4 + 3][ir] = v.w;
            if (ti != tj) {
                const uint4 w = *reinterpret_cast<const uint4 *>(obj + (size_t)(tj * 64 + ir) * n + ti * 64 + vq * 4);
                t1[vq * 4 + 0][ir] = w.x; t1[vq * 4 + 1][ir] = w.y; t1[vq * 4 + 2][ir] = w.z; t1[vq * 4 + 3][ir] = w.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (unsigned p = 0; p < 4; ++p) {
            const unsigned orow = rr + 16 * p;
            uint4 v;
            v.x = t0[orow][vq * 4 + 0]; v.y = t0[orow][vq * 4 + 1]; v.z = t0[orow][vq * 4 + 2]; v.w = t0[orow][vq * 4 + 3];
            *reinterpret_cast<uint4 *>(obj + (size_t)(tj * 64 + orow) * n + ti * 64 + vq * 4) = v;
            if (ti != tj) {
                uint4 w;
                w.x = t1[orow][vq * 4 + 0]; w.y = t1[orow][vq * 4 + 1]; w.z = t1[orow][vq * 4 + 2]; w.w = t1[orow][vq * 4 + 3];
                *reinterpret_cast<uint4 *>(obj + (size_t)(ti * 64 + orow) * n + tj * 64 + vq * 4) = w;
            }
        }
        __syncthreads();
    }
}

// 4- and 8-byte elements, rows of whole 16-byte vectors (n % VW == 0), 16-byte aligned: the same exchange on TS x TS tiles
// (64 x 64 for 4-byte, 32 x 32 for 8-byte elements: 16 vectors per tile row) with 128-bit accesses; edge tiles are
// predicated per vector, so any such n qualifies (96, 200 f32, 100 f64: 0.44-0.57 of HBM peak on the element-wise kernel
// above, 0.68-0.75 here; whole 64 x 64 tiles of 4-byte elements keep the unpredicated kernel: 0.97-1.02 against 0.89-0.91).
template <typename E>
__global__ void __launch_bounds__(256)
transpose_inplace_square_vec_tiles_kernel(E *data, size_t batch, unsigned n) {
    constexpr int VW = 16 / (int)sizeof(E), TS = 16 * VW, PASSES = TS / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E (*t0)[TS + 1] = reinterpret_cast<E (*)[TS + 1]>(smem_raw);
    E (*t1)[TS + 1] = t0 + TS;
    const unsigned ti = blockIdx.y, tj = blockIdx.x;
    if (ti > tj) return;
    const unsigned vq = threadIdx.x & 15, rr = threadIdx.x >> 4;
    struct alignas(16) Vec { E e[VW]; };
    for (size_t o = blockIdx.z; o < batch; o += gridDim.z) {
        E *obj = data + o * (size_t)n * n;
#pragma unroll
        for (unsigned p = 0; p < PASSES; ++p) {
            const unsigned ir = rr + 16 * p;
            const unsigned r0 = ti * TS + ir, c0 = tj * TS + vq * VW;
            if (r0 < n && c0 < n) {
                const Vec v = *reinterpret_cast<const Vec *>(obj + (size_t)r0 * n + c0);
#pragma unroll
                for (int q = 0; q < VW; ++q) t0[vq * VW + q][ir] = v.e[q];
            }
            const unsigned r1 = tj * TS + ir, c1 = ti * TS + vq * VW;
            if (ti != tj && r1 < n && c1 < n) {
                const Vec w = *reinterpret_cast<const Vec *>(obj + (size_t)r1 * n + c1);
#pragma unroll
                for (int q = 0; q < VW; ++q) t1[vq * VW + q][ir] = w.e[q];
            }
        }
        __syncthreads();
#pragma unroll
        for (unsigned p = 0; p < PASSES; ++p) {
            const unsigned orow = rr + 16 * p;
            // the mirror position (tile (tj, ti)) receives t0 transposed: its row orow is t0's column orow
            const unsigned r1 = tj * TS + orow, c1 = ti * TS + vq * VW;
            if (r1 < n && c1 < n) {
                Vec v;
#pragma unroll
                for (int q = 0; q < VW; ++q) v.e[q] = t0[orow][vq * VW + q];
                *reinterpret_cast<Vec *>(obj + (size_t)r1 * n + c1) = v;
            }
            const unsigned r0 = ti * TS + orow, c0 = tj * TS + vq * VW;
            if (ti != tj && r0 < n && c0 < n) {
                Vec w;
#pragma unroll
                for (int q = 0; q < VW; ++q) w.e[q] = t1[orow][vq * VW + q];
                *reinterpret_cast<Vec *>(obj + (size_t)r0 * n + c0) = w;
            }
        }
        __syncthreads();
    }
}

// ---- (4b) in place, square n x n (n a multiple of the vector width, n <= 64): register block pairs -----
// One thread owns the VW x VW blocks (bi, bj) and (bj, bi), bi <= bj: 2*VW 128-bit loads, two register
// transposes, 2*VW 128-bit stores into the exchanged places.  The upper triangle of the nb x nb block grid
// is enumerated by folding row q onto row nb-1-q (together nb+1 blocks), so every thread has work.
__device__ __forceinline__ void tri_decode(unsigned nb, unsigned w, unsigned &bi, unsigned &bj) {
    if (nb & 1) {  // rows q (>= 1) and nb-q hold nb blocks together; row 0 holds nb alone
        const unsigned q = w / nb, r = w - q * nb;
        if (q == 0 || r < nb - q) { bi = q; bj = q + r; }
        else { bi = nb - q; bj = nb - q + (r - (nb - q)); }
    } else {       // rows q and nb-1-q hold nb+1 blocks together
        const unsigned q = w / (nb + 1), r = w - q * (nb + 1);
        if (r < nb - q) { bi = q; bj = q + r; }
        else { bi = nb - 1 - q; bj = nb - 1 - q + (r - (nb - q)); }
    }
}
template <typename E, int VW>
__global__ void __launch_bounds__(256)
transpose_square_vec_inplace_kernel(E *data, size_t npairs_total, unsigned n) {
    struct alignas(VW * sizeof(E)) V { E v[VW]; };
    const unsigned nb = n / VW;
    const unsigned pairs_per_obj = nb * (nb + 1) / 2;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < npairs_total; t += (size_t)gridDim.x * 256) {
        const size_t obj = t / pairs_per_obj;
        unsigned bi, bj;
        tri_decode(nb, (unsigned)(t - obj * pairs_per_obj), bi, bj);
        E *p0 = data + obj * (size_t)n * n + (size_t)(bi * VW) * n + bj * VW;
        E *p1 = data + obj * (size_t)n * n + (size_t)(bj * VW) * n + bi * VW;
        V x[VW], z[VW];
#pragma unroll
        for (int s = 0; s < VW; ++s) x[s] = ldvec(reinterpret_cast<const V *>(p0 + (size_t)s * n));
#pragma unroll
        for (int s = 0; s < VW; ++s) z[s] = ldvec(reinterpret_cast<const V *>(p1 + (size_t)s * n));
#pragma unroll
        for (int s = 0; s < VW; ++s) {
            V y, u;
#pragma unroll
            for (int q = 0; q < VW; ++q) { y.v[q] = x[q].v[s]; u.v[q] = z[q].v[s]; }
            stvec(reinterpret_cast<V *>(p1 + (size_t)s * n), y);
            if (bi != bj) stvec(reinterpret_cast<V *>(p0 + (size_t)s * n), u);
        }
    }
}

template <typename E, int VW, int MULT, bool OUT_MAJOR>
static int launch_square_vec(Context *ctx, cudaStream_t st, size_t batch, unsigned n, const E *a, E *buffer) {
    const size_t total = batch * (size_t)(n / (MULT * VW)) * (n / VW);
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    transpose_square_vec_kernel<E, VW, MULT, OUT_MAJOR><<<(unsigned)blocks, 256, 0, st>>>(a, buffer, total, n);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

template <typename E>
static int launch_transpose_typed(Context *ctx, cudaStream_t st, size_t batch, size_t len, const E *a, E *buffer,
                                  size_t columns) {
    const size_t rows = len / columns;
    if (rows == 0 || batch == 0) return SLAS_OK;
    constexpr int VW = 16 / (int)sizeof(E);
    // (3) square, vector-aligned (4- and 8-byte elements; narrower ones take the tiled / per-object kernels)
    if constexpr (VW > 1 && sizeof(E) >= 4) {
      if (rows == columns && len == rows * columns && rows % VW == 0 && rows <= 64 && aligned16(a) &&
          aligned16(buffer)) {
        const unsigned n = (unsigned)rows;
        const long variant = tunable("transpose_variant");
        // measured (B200): the lane-per-row shuffle kernel wins for 2x2 of 8-byte elements (0.97 of HBM peak) but
        // its select chains lose to the thread-per-matrix register transpose for 4x4 of 4-byte ones (0.63 vs 0.82)
        if (n == (unsigned)VW && (VW == 2 ? variant != 1 : variant == 6)) {
            const size_t nrows_total = batch * VW;
            size_t blocks = (nrows_total + 255) / 256;
            const size_t cap = (size_t)ctx->sm_count * 32;
            if (blocks > cap) blocks = cap;
            transpose_tiny_kernel<E, VW><<<(unsigned)blocks, 256, 0, st>>>(a, buffer, nrows_total);
            SLAS_LAUNCH_CHECK();
            return SLAS_OK;
        }
        if constexpr (sizeof(E) == 4) {
            if (n == 4 && variant != 7 && variant != 1 && ((uintptr_t)a % 32 == 0) && ((uintptr_t)buffer % 32 == 0)) {
                size_t blocks = (batch + 255) / 256;
                const size_t cap = (size_t)ctx->sm_count * 32;
                if (blocks > cap) blocks = cap;
                transpose_4x4_b32_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(a),
                                                                          reinterpret_cast<uint32_t *>(buffer), batch);
                SLAS_LAUNCH_CHECK();
                return SLAS_OK;
            }
        }
        if constexpr (sizeof(E) <= 8) {
            constexpr int VW2 = 2 * VW;  // elements per 32-byte vector
            // 256-bit vectors: measured (profiles/r1e_tune.log) 0.97-1.04 of HBM peak against 0.90-1.00 with
            // 128-bit ones for 16..64 f32 / f64; a single VW2 x VW2 block per object (8x8 f32) is the exception
            if ((variant == 0 || variant == 8 || variant == 9) && n % VW2 == 0 && n > (unsigned)VW2 &&
                ((uintptr_t)a % 32 == 0) && ((uintptr_t)buffer % 32 == 0)) {
                if (variant == 9) return launch_square_vec<E, VW2, 1, false>(ctx, st, batch, n, a, buffer);
                return launch_square_vec<E, VW2, 1, true>(ctx, st, batch, n, a, buffer);
            }
        }
        // 128-bit vectors, out-major lanes (full-line stores; profiles/r1b_tune.log); variant 1 = in-major ordering
        if (variant == 1) return launch_square_vec<E, VW, 1, false>(ctx, st, batch, n, a, buffer);
        if (variant == 2 && n % (2 * VW) == 0) return launch_square_vec<E, VW, 2, true>(ctx, st, batch, n, a, buffer);
        if (variant == 3 && n % (4 * VW) == 0) return launch_square_vec<E, VW, 4, true>(ctx, st, batch, n, a, buffer);
        if (variant == 4 && n % (2 * VW) == 0) return launch_square_vec<E, VW, 2, false>(ctx, st, batch, n, a, buffer);
        if (variant == 5 && n % (4 * VW) == 0) return launch_square_vec<E, VW, 4, false>(ctx, st, batch, n, a, buffer);
        return launch_square_vec<E, VW, 1, true>(ctx, st, batch, n, a, buffer);
      }
    }
    // (1c) narrow elements in whole 128-byte tiles
    if constexpr (sizeof(E) <= 2) {
        SLAS_REQUIRE(rows <= 0xffffffffu && columns <= 0xffffffffu, "transpose: dimension too large");
        constexpr size_t T = 128, TRI = 128;  // tile: 128 input rows of 128 elements
        if (rows % T == 0 && columns % TRI == 0 && len == rows * columns && (uintptr_t)a % 32 == 0 && (uintptr_t)buffer % 32 == 0 &&
            tunable("transpose_variant") != 1) {
            const size_t t_r = rows / T, t_c = columns / TRI, nb = batch * t_r * t_c;
            SLAS_REQUIRE(nb <= 0x7fffffffu, "transpose: too many tiles (%zu)", nb);
            transpose_tiled_narrow_kernel<E><<<(unsigned)nb, 32 * (TRI * sizeof(E) / 32), 0, st>>>(a, buffer, len, (unsigned)rows, (unsigned)columns,
                                                                          (unsigned)t_r, (unsigned)t_c);
            SLAS_LAUNCH_CHECK();
            return SLAS_OK;
        }
    }
    // (2b) many small objects: bulk-TMA staged gather
    {
        bool staged = false;
        SLAS_TRY(launch_transpose_staged<E>(ctx, st, batch, len, a, buffer, rows, columns, &staged));
        if (staged) return SLAS_OK;
    }
    // (2) small objects: padded tile must fit the per-warp shared slice
    const size_t padded = columns * (rows + 1);
    constexpr size_t kWarpBytes = 12 * 1024;
    if (padded * sizeof(E) <= kWarpBytes && len <= 0xffffffffu) {
        const size_t smem = 8 * padded * sizeof(E);
        if (smem > 48 * 1024)
            SLAS_CUDA_TRY(cudaFuncSetAttribute(transpose_small_kernel<E, false>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * kWarpBytes)));
        size_t blocks = (batch + 7) / 8;
        const size_t cap = (size_t)ctx->sm_count * 8;
        if (blocks > cap) blocks = cap;
        transpose_small_kernel<E, false><<<(unsigned)blocks, 256, smem, st>>>(a, buffer, batch, (unsigned)len,
                                                                              (unsigned)rows, (unsigned)columns,
                                                                              (unsigned)padded);
        SLAS_LAUNCH_CHECK();
        return SLAS_OK;
    }
    // (1) tiled
    SLAS_REQUIRE(rows <= 0xffffffffu && columns <= 0xffffffffu, "transpose: dimension too large");
    if constexpr (sizeof(E) == 4) {
        if (rows % 64 == 0 && columns % 64 == 0 && len == rows * columns && aligned16(a) && aligned16(buffer) &&
            tunable("transpose_variant") != 1) {
            const size_t t_r = rows / 64, t_c = columns / 64, nb = batch * t_r * t_c;
            SLAS_REQUIRE(nb <= 0x7fffffffu, "transpose: too many tiles (%zu)", nb);
            transpose_tiled64_kernel<<<(unsigned)nb, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(a),
                                                                    reinterpret_cast<uint32_t *>(buffer), len, (unsigned)rows,
                                                                    (unsigned)columns, (unsigned)t_r, (unsigned)t_c);
            SLAS_LAUNCH_CHECK();
            return SLAS_OK;
        }
    }
    const size_t tiles_r = (rows + 31) / 32, tiles_c = (columns + 31) / 32;
    const size_t blocks = batch * tiles_r * tiles_c;
    SLAS_REQUIRE(blocks <= 0x7fffffffu, "transpose: too many tiles (%zu)", blocks);
    transpose_tiled_kernel<E><<<(unsigned)blocks, 256, 0, st>>>(a, buffer, len, (unsigned)rows, (unsigned)columns,
                                                                (unsigned)tiles_r, (unsigned)tiles_c);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}


// In place.  *handled = false when no in-place kernel applies (the caller falls back to the reference's
// own scheme: zeroed temporary, transpose, copy back).
template <typename E>
static int launch_transpose_inplace_typed(Context *ctx, cudaStream_t st, size_t batch, size_t len, E *a, size_t columns,
                                          bool *handled) {
    *handled = false;
    const size_t rows = len / columns;
    if (rows == 0 || batch == 0) return SLAS_OK;
    constexpr int VW = 16 / (int)sizeof(E);
    const bool square = rows == columns && len == rows * columns;
    // one thread owns the whole VW x VW object: all loads precede its stores
    if constexpr (VW > 1 && sizeof(E) >= 4) {
        if (square && rows == (size_t)VW && aligned16(a)) {
            *handled = true;
            return launch_square_vec<E, VW, 1, true>(ctx, st, batch, (unsigned)rows, a, a);
        }
    }
    if constexpr (VW > 1 && sizeof(E) >= 4) {
        if (square && rows % VW == 0 && rows < 64 && aligned16(a) && tunable("transpose_variant") != 1) {
            const unsigned n = (unsigned)rows;
            const bool wide = sizeof(E) <= 8 && n % (2 * VW) == 0 && n > 2u * VW && (uintptr_t)a % 32 == 0 &&
                              tunable("transpose_variant") != 10;
            const unsigned nb = n / (wide ? 2 * VW : VW);
            const size_t total = batch * (size_t)(nb * (nb + 1) / 2);
            size_t blocks = (total + 255) / 256;
            const size_t cap = (size_t)ctx->sm_count * 32;
            if (blocks > cap) blocks = cap;
            if constexpr (sizeof(E) <= 8) {
                if (wide) transpose_square_vec_inplace_kernel<E, 2 * VW><<<(unsigned)blocks, 256, 0, st>>>(a, total, n);
            }
            if (!wide) transpose_square_vec_inplace_kernel<E, VW><<<(unsigned)blocks, 256, 0, st>>>(a, total, n);
            SLAS_LAUNCH_CHECK();
            *handled = true;
            return SLAS_OK;
        }
    }
    // many small objects: the bulk-TMA staged gather reads a chunk completely before it stores it
    {
        bool staged = false;
        SLAS_TRY(launch_transpose_staged<E>(ctx, st, batch, len, a, a, rows, columns, &staged));
        if (staged) { *handled = true; return SLAS_OK; }
    }
    // one warp owns the whole object in shared memory
    const size_t padded = columns * (rows + 1);
    constexpr size_t kWarpBytes = 12 * 1024;
    if (padded * sizeof(E) <= kWarpBytes && len <= 0xffffffffu) {
        const size_t smem = 8 * padded * sizeof(E);
        if (smem > 48 * 1024)
            SLAS_CUDA_TRY(cudaFuncSetAttribute(transpose_small_kernel<E, true>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * kWarpBytes)));
        size_t blocks = (batch + 7) / 8;
        const size_t cap = (size_t)ctx->sm_count * 8;
        if (blocks > cap) blocks = cap;
        transpose_small_kernel<E, true><<<(unsigned)blocks, 256, smem, st>>>(a, a, batch, (unsigned)len, (unsigned)rows,
                                                                             (unsigned)columns, (unsigned)padded);
        SLAS_LAUNCH_CHECK();
        *handled = true;
        return SLAS_OK;
    }
    if (!square || rows > 0xffffffffu) return SLAS_OK;
    const unsigned n = (unsigned)rows;
    const unsigned gz = (unsigned)std::min<size_t>(batch, 65535);
    if constexpr (sizeof(E) == 4) {
        if (n % 64 == 0 && n / 64 <= 65535 && aligned16(a)) {
            constexpr int kSmem = 2 * 64 * 65 * 4;
            // (per call: the attribute belongs to the function ON THIS DEVICE, a process may drive several)
            SLAS_CUDA_TRY(cudaFuncSetAttribute(transpose_inplace_square64_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
            transpose_inplace_square64_kernel<<<dim3(n / 64, n / 64, gz), 256, kSmem, st>>>(
                reinterpret_cast<uint32_t *>(a), batch, n);
            SLAS_LAUNCH_CHECK();
            *handled = true;
            return SLAS_OK;
        }
    }
    if constexpr (sizeof(E) == 4 || sizeof(E) == 8) {
        constexpr unsigned VW = 16 / sizeof(E), TS = 16 * VW;
        const unsigned vt = (n + TS - 1) / TS;
        if (n % VW == 0 && vt <= 65535 && aligned16(a) && tunable("transpose_variant") != 13) {
            constexpr int kSmem = 2 * TS * (TS + 1) * (int)sizeof(E);
            // (per call: the attribute belongs to the function ON THIS DEVICE, a process may drive several)
            SLAS_CUDA_TRY(cudaFuncSetAttribute(transpose_inplace_square_vec_tiles_kernel<E>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
            transpose_inplace_square_vec_tiles_kernel<E><<<dim3(vt, vt, gz), 256, kSmem, st>>>(a, batch, n);
            SLAS_LAUNCH_CHECK();
            *handled = true;
            return SLAS_OK;
        }
    }
    const unsigned tiles = (n + 31) / 32;
    if (tiles > 65535) return SLAS_OK;
    transpose_inplace_square_kernel<E><<<dim3(tiles, tiles, gz), 256, 0, st>>>(a, batch, n);
    SLAS_LAUNCH_CHECK();
    *handled = true;
    return SLAS_OK;
}

// ---- any other element size (`Transpose<T: Copy>` is generic over T, rust.rs:151: a 12-byte [f32; 3], a 24-byte
// struct ...): the element is W granules of the widest power-of-two size G that divides it and that both pointers are
// aligned to; 32 x 32 element tiles go through shared memory as rows of 32 * W granules, so both sides move contiguous
// runs.  Same index map as above: out[(row * columns + column) * W + j] = in[(column * rows + row) * W + j].
template <typename G>
__global__ void __launch_bounds__(256)
transpose_granules_kernel(const G *__restrict__ a, G *__restrict__ out, size_t batch, size_t len, unsigned rows, unsigned columns,
                          unsigned w) {
    extern __shared__ __align__(16) unsigned char tile_raw[];
    G *tile = reinterpret_cast<G *>(tile_raw);            // [32][32 * w + 1]
    const unsigned seg = 32 * w, pitch = seg + 1;
    const unsigned c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (size_t obj = blockIdx.z; obj < batch; obj += gridDim.z) {
        const G *src = a + obj * len * w;
        G *dst = out + obj * len * w;
        for (unsigned idx = threadIdx.x; idx < 32 * seg; idx += 256) {
            const unsigned ic = idx / seg, q = idx - ic * seg;
            if (c0 + ic < columns && r0 + q / w < rows) tile[ic * pitch + q] = src[((size_t)(c0 + ic) * rows + r0) * w + q];
        }
        __syncthreads();
        for (unsigned idx = threadIdx.x; idx < 32 * seg; idx += 256) {
            const unsigned orow = idx / seg, q = idx - orow * seg;
            const unsigned oc = q / w, j = q - oc * w;
            if (r0 + orow < rows && c0 + oc < columns) dst[((size_t)(r0 + orow) * columns + c0) * w + q] = tile[oc * pitch + orow * w + j];
        }
        __syncthreads();
    }
}

template <typename G>
static int launch_transpose_granules(Context *ctx, cudaStream_t st, size_t batch, size_t len, size_t elem_size, const void *a, void *buffer,
                                     size_t columns) {
    const size_t rows = len / columns, w = elem_size / sizeof(G);
    if (rows == 0 || batch == 0) return SLAS_OK;
    SLAS_REQUIRE(rows <= 0x7fffffffu && columns <= 0x7fffffffu, "transpose: object too large");
    const size_t smem = (size_t)32 * (32 * w + 1) * sizeof(G);
    SLAS_REQUIRE(smem <= 200 * 1024, "transpose: element size %zu is too large (the 32 x 32 element tile must fit shared memory)", elem_size);
    const size_t gx = (columns + 31) / 32, gy = (rows + 31) / 32;
    SLAS_REQUIRE(gy <= 65535, "transpose: more than 2^21 rows of a generic-size element");
    auto kern = transpose_granules_kernel<G>;
    if (smem > 48 * 1024) SLAS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    size_t gz = batch;
    const size_t cap = std::max<size_t>(1, (size_t)ctx->sm_count * 16 / (gx * gy));
    if (gz > cap) gz = cap;
    if (gz > 65535) gz = 65535;
    kern<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), 256, smem, st>>>((const G *)a, (G *)buffer, batch, len, (unsigned)rows,
                                                                            (unsigned)columns, (unsigned)w);
    SLAS_LAUNCH_CHECK();
    return SLAS_OK;
}

static int launch_transpose_any_size(Context *ctx, cudaStream_t st, size_t batch, size_t len, size_t elem_size, const void *a, void *buffer,
                                     size_t columns) {
    SLAS_REQUIRE(elem_size > 0, "transpose: element size 0");
    const uintptr_t bits = (uintptr_t)a | (uintptr_t)buffer | (uintptr_t)elem_size;
    if (bits % 16 == 0) return launch_transpose_granules<Elem16>(ctx, st, batch, len, elem_size, a, buffer, columns);
    if (bits % 8 == 0) return launch_transpose_granules<uint64_t>(ctx, st, batch, len, elem_size, a, buffer, columns);
    if (bits % 4 == 0) return launch_transpose_granules<uint32_t>(ctx, st, batch, len, elem_size, a, buffer, columns);
    if (bits % 2 == 0) return launch_transpose_granules<uint16_t>(ctx, st, batch, len, elem_size, a, buffer, columns);
    return launch_transpose_granules<uint8_t>(ctx, st, batch, len, elem_size, a, buffer, columns);
}

int launch_transpose_inplace(Context *ctx, cudaStream_t st, size_t batch, size_t len, size_t elem_size, void *a,
                             size_t columns, bool *handled) {
    SLAS_REQUIRE(columns != 0, "transpose: columns == 0 (the reference divides by it, rust.rs:169)");
    switch (elem_size) {
    case 1: return launch_transpose_inplace_typed<uint8_t>(ctx, st, batch, len, (uint8_t *)a, columns, handled);
    case 2: return launch_transpose_inplace_typed<uint16_t>(ctx, st, batch, len, (uint16_t *)a, columns, handled);
    case 4: return launch_transpose_inplace_typed<uint32_t>(ctx, st, batch, len, (uint32_t *)a, columns, handled);
    case 8: return launch_transpose_inplace_typed<uint64_t>(ctx, st, batch, len, (uint64_t *)a, columns, handled);
    case 16: return launch_transpose_inplace_typed<Elem16>(ctx, st, batch, len, (Elem16 *)a, columns, handled);
    }
    *handled = false;  // any other element size: through the temporary, like the reference (rust.rs:152-160)
    return SLAS_OK;
}

int launch_transpose(Context *ctx, cudaStream_t st, size_t batch, size_t len, size_t elem_size, const void *a,
                     void *buffer, size_t columns) {
    SLAS_REQUIRE(columns != 0, "transpose: columns == 0 (the reference divides by it, rust.rs:169)");
    SLAS_REQUIRE(a != buffer, "transpose: input and buffer alias (use transpose_inplace)");
    switch (elem_size) {
    case 1: return launch_transpose_typed<uint8_t>(ctx, st, batch, len, (const uint8_t *)a, (uint8_t *)buffer, columns);
    case 2: return launch_transpose_typed<uint16_t>(ctx, st, batch, len, (const uint16_t *)a, (uint16_t *)buffer, columns);
    case 4: return launch_transpose_typed<uint32_t>(ctx, st, batch, len, (const uint32_t *)a, (uint32_t *)buffer, columns);
    case 8: return launch_transpose_typed<uint64_t>(ctx, st, batch, len, (const uint64_t *)a, (uint64_t *)buffer, columns);
    case 16: return launch_transpose_typed<Elem16>(ctx, st, batch, len, (const Elem16 *)a, (Elem16 *)buffer, columns);
    }
    return launch_transpose_any_size(ctx, st, batch, len, elem_size, a, buffer, columns);
}

}  // namespace slas
