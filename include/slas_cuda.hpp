// slas_cuda.hpp -- typed C++17 host mirror of the reference's interface for the hot path, on top of the
// C ABI (slas_cuda.h).  The reference is a Rust crate whose shapes are const generics; this header keeps
// the same names, argument order, argument meaning and failure behaviour with C++ templates, so that
// tests/cpp/test_reference_suite.cpp reads like the reference's tests/src/main.rs:
//
//   slas::backends::Cuda                 the ZST backend (src/backends/rust.rs:2-3): dot / norm / normalize /
//                                        add..div / transpose[_inplace] / matrix_mul / matrix_vector_mul,
//                                        exactly the operation traits of src/backends.rs:121-183
//   slas::CudaVec<T, LEN>                device-resident StaticVec (host accessors throw, cf. src/nullvec.rs)
//   std::array<T, LEN>                   plays [T; LEN] / StaticVecUnion (host pointer, staged by the library)
//   slas::WithStaticBackend<U, B>        src/backends.rs:185-249, 306-331
//   slas::MatrixShape<M, K>, Matrix<...> src/tensor.rs:86-115, 476-523, 364-465 (lazy transpose = a bool
//                                        template parameter, consumed as a_trans / b_trans)
//   slas::Panic                          thrown where the reference panics / asserts
//
// No arithmetic happens here: every op is one call into libslas_cuda.so.
#pragma once
#include <array>
#include <complex>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <type_traits>

#include "slas_cuda.h"

namespace slas {

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void check(int status) {
    if (status != SLAS_OK) throw Panic("slas_cuda status " + std::to_string(status) + ": " + slas_cuda_last_error());
}
#define SLAS_ASSERT(cond, msg)                 \
    do {                                       \
        if (!(cond)) throw ::slas::Panic(msg); \
    } while (0)

// ---- StaticVec contract: LEN contiguous elements behind as_ptr() (src/static_vec.rs:121-134) ----------------
template <class V> struct static_vec_traits;  // ::value_type, ::LEN, ::ptr(v), ::mut_ptr(v)

template <class T, std::size_t N> struct static_vec_traits<std::array<T, N>> {
    using value_type = T;
    static constexpr std::size_t LEN = N;
    static const T *ptr(const std::array<T, N> &v) { return v.data(); }
    static T *mut_ptr(std::array<T, N> &v) { return v.data(); }
};

// Device-resident StaticVec: as_ptr() is a device pointer; data moves only through from_host / to_host.
template <class T, std::size_t N> class CudaVec {
  public:
    static constexpr std::size_t LEN = N;
    using value_type = T;
    CudaVec() { check(slas_cuda_malloc(reinterpret_cast<void **>(&p_), N * sizeof(T))); }
    explicit CudaVec(const std::array<T, N> &host) : CudaVec() { check(slas_cuda_memcpy_h2d(p_, host.data(), N * sizeof(T))); }
    CudaVec(const CudaVec &) = delete;
    CudaVec &operator=(const CudaVec &) = delete;
    CudaVec(CudaVec &&o) noexcept : p_(o.p_) { o.p_ = nullptr; }
    ~CudaVec() { slas_cuda_free(p_); }
    const T *as_ptr() const { return p_; }
    T *as_mut_ptr() { return p_; }
    std::array<T, N> to_host() const {
        std::array<T, N> out{};
        check(slas_cuda_memcpy_d2h(out.data(), p_, N * sizeof(T)));
        return out;
    }
    // host accessors are not available on device storage (precedent: NullVec, src/nullvec.rs:24-57)
    const T &operator[](std::size_t) const { throw Panic("Tried to call get_unchecked on CudaVec: device memory"); }

  private:
    T *p_ = nullptr;
};
template <class T, std::size_t N> struct static_vec_traits<CudaVec<T, N>> {
    using value_type = T;
    static constexpr std::size_t LEN = N;
    static const T *ptr(const CudaVec<T, N> &v) { return v.as_ptr(); }
    static T *mut_ptr(CudaVec<T, N> &v) { return v.as_mut_ptr(); }
};

// `StaticCowVec<'a, T, LEN>` (src/lib.rs:395-456; StaticVec impl src/static_vec.rs:294-319): borrows a host array
// until the first mutable access, which copies it ("moo": copy on write).  as_ptr() never copies.
template <class T, std::size_t N> class StaticCowVec {
  public:
    static constexpr std::size_t LEN = N;
    using value_type = T;
    explicit StaticCowVec(const std::array<T, N> &borrowed) : borrowed_(&borrowed) {}
    static StaticCowVec from_owned(const std::array<T, N> &a) {
        StaticCowVec v(a);
        v.owned_ = a;
        v.borrowed_ = nullptr;
        return v;
    }
    bool is_owned() const { return borrowed_ == nullptr; }
    bool is_borrowed() const { return !is_owned(); }
    const T *as_ptr() const { return is_owned() ? owned_.data() : borrowed_->data(); }
    T *as_mut_ptr() {  // deref_mut: `self.data.owned = *self.data.borrowed` (src/lib.rs:446-455)
        if (!is_owned()) {
            owned_ = *borrowed_;
            borrowed_ = nullptr;
        }
        return owned_.data();
    }
    const T &operator[](std::size_t i) const { return as_ptr()[i]; }

  private:
    const std::array<T, N> *borrowed_;
    std::array<T, N> owned_{};
};
template <class T, std::size_t N> struct static_vec_traits<StaticCowVec<T, N>> {
    using value_type = T;
    static constexpr std::size_t LEN = N;
    static const T *ptr(const StaticCowVec<T, N> &v) { return v.as_ptr(); }
    static T *mut_ptr(StaticCowVec<T, N> &v) { return v.as_mut_ptr(); }
};

template <class V> using elem_t = typename static_vec_traits<std::remove_cv_t<std::remove_reference_t<V>>>::value_type;
template <class V> constexpr std::size_t len_v = static_vec_traits<std::remove_cv_t<std::remove_reference_t<V>>>::LEN;
template <class V> const elem_t<V> *ptr_of(const V &v) { return static_vec_traits<V>::ptr(v); }
template <class V> elem_t<V> *mut_ptr_of(V &v) { return static_vec_traits<V>::mut_ptr(v); }

// `moo![T: a..b]` (src/lib.rs:293-325) for ranges: a host array of consecutive values
template <class T, std::size_t N> std::array<T, N> moo_range(T first) {
    std::array<T, N> a{};
    for (std::size_t i = 0; i < N; ++i) a[i] = first + static_cast<T>(i);
    return a;
}

namespace backends {

// The operation traits of src/backends.rs:121-183, same method names and argument order.
struct Cuda {
    // DotProduct::dot -- both operands carry the same LEN at compile time (src/backends.rs:122-126)
    template <class A, class B> elem_t<A> dot(const A &a, const B &b) const {
        static_assert(len_v<A> == len_v<B>, "dot: vectors of different static lengths");
        static_assert(std::is_same_v<elem_t<A>, elem_t<B>>, "dot: element types differ");
        using T = elem_t<A>;
        T out{};
        if constexpr (std::is_same_v<T, float>) check(slas_cuda_sdot(len_v<A>, ptr_of(a), ptr_of(b), &out));
        else if constexpr (std::is_same_v<T, double>) check(slas_cuda_ddot(len_v<A>, ptr_of(a), ptr_of(b), &out));
        else if constexpr (std::is_same_v<T, std::complex<float>>)
            check(slas_cuda_cdotu(len_v<A>, reinterpret_cast<const float *>(ptr_of(a)), reinterpret_cast<const float *>(ptr_of(b)),
                                  reinterpret_cast<float *>(&out)));
        else
            check(slas_cuda_zdotu(len_v<A>, reinterpret_cast<const double *>(ptr_of(a)), reinterpret_cast<const double *>(ptr_of(b)),
                                  reinterpret_cast<double *>(&out)));
        return out;
    }
    // Normalize::norm -- NormOutput is the real type (src/backends/rust.rs:117, 130)
    template <class A> auto norm(const A &a) const {
        using T = elem_t<A>;
        if constexpr (std::is_same_v<T, float>) { float o; check(slas_cuda_snrm2(len_v<A>, ptr_of(a), &o)); return o; }
        else if constexpr (std::is_same_v<T, double>) { double o; check(slas_cuda_dnrm2(len_v<A>, ptr_of(a), &o)); return o; }
        else if constexpr (std::is_same_v<T, std::complex<float>>) {
            float o; check(slas_cuda_scnrm2(len_v<A>, reinterpret_cast<const float *>(ptr_of(a)), &o)); return o;
        } else { double o; check(slas_cuda_dznrm2(len_v<A>, reinterpret_cast<const double *>(ptr_of(a)), &o)); return o; }
    }
    template <class A> void normalize(A &a) const {
        using T = elem_t<A>;
        if constexpr (std::is_same_v<T, float>) check(slas_cuda_snormalize(len_v<A>, mut_ptr_of(a)));
        else if constexpr (std::is_same_v<T, double>) check(slas_cuda_dnormalize(len_v<A>, mut_ptr_of(a)));
        else if constexpr (std::is_same_v<T, std::complex<float>>) check(slas_cuda_cnormalize(len_v<A>, reinterpret_cast<float *>(mut_ptr_of(a))));
        else check(slas_cuda_znormalize(len_v<A>, reinterpret_cast<double *>(mut_ptr_of(a))));
    }
    // Addition / Subtraction / Multiplication / Divition (src/backends.rs:156-182)
#define SLAS_EWISE(NAME)                                                                                        \
    template <class A, class B, class C> void NAME(const A &a, const B &b, C &c) const {                        \
        static_assert(len_v<A> == len_v<B> && len_v<A> == len_v<C>, #NAME ": vectors of different static lengths"); \
        if constexpr (std::is_same_v<elem_t<A>, float>) check(slas_cuda_s##NAME(len_v<A>, ptr_of(a), ptr_of(b), mut_ptr_of(c))); \
        else check(slas_cuda_d##NAME(len_v<A>, ptr_of(a), ptr_of(b), mut_ptr_of(c)));                           \
    }
    SLAS_EWISE(add) SLAS_EWISE(sub) SLAS_EWISE(mul) SLAS_EWISE(div)
#undef SLAS_EWISE
    // Fused chains (slas_cuda.h "Fused chains"): X_dot(a, b, d) = dot(a X b, d) without materialising a X b;
    // X_dot(a, b, d, c) also writes the intermediate.  Same bits as the unfused chain on this backend.
#define SLAS_EWISE_DOT(NAME)                                                                                    \
    template <class A, class B, class D> elem_t<A> NAME##_dot(const A &a, const B &b, const D &d) const {       \
        static_assert(len_v<A> == len_v<B> && len_v<A> == len_v<D>, #NAME "_dot: vectors of different static lengths"); \
        elem_t<A> out{};                                                                                        \
        if constexpr (std::is_same_v<elem_t<A>, float>) check(slas_cuda_s##NAME##_dot(len_v<A>, ptr_of(a), ptr_of(b), nullptr, ptr_of(d), &out)); \
        else check(slas_cuda_d##NAME##_dot(len_v<A>, ptr_of(a), ptr_of(b), nullptr, ptr_of(d), &out));          \
        return out;                                                                                             \
    }                                                                                                           \
    template <class A, class B, class D, class C> elem_t<A> NAME##_dot(const A &a, const B &b, const D &d, C &c) const { \
        static_assert(len_v<A> == len_v<B> && len_v<A> == len_v<D> && len_v<A> == len_v<C>, #NAME "_dot: vectors of different static lengths"); \
        elem_t<A> out{};                                                                                        \
        if constexpr (std::is_same_v<elem_t<A>, float>) check(slas_cuda_s##NAME##_dot(len_v<A>, ptr_of(a), ptr_of(b), mut_ptr_of(c), ptr_of(d), &out)); \
        else check(slas_cuda_d##NAME##_dot(len_v<A>, ptr_of(a), ptr_of(b), mut_ptr_of(c), ptr_of(d), &out));    \
        return out;                                                                                             \
    }
    SLAS_EWISE_DOT(add) SLAS_EWISE_DOT(sub) SLAS_EWISE_DOT(mul) SLAS_EWISE_DOT(div)
#undef SLAS_EWISE_DOT
    // { dot(a, b), norm(a), norm(b) } from one read of both vectors
    template <class A, class B> std::array<elem_t<A>, 3> dot_norms(const A &a, const B &b) const {
        static_assert(len_v<A> == len_v<B>, "dot_norms: vectors of different static lengths");
        std::array<elem_t<A>, 3> out{};
        if constexpr (std::is_same_v<elem_t<A>, float>) check(slas_cuda_sdot_norms(len_v<A>, ptr_of(a), ptr_of(b), out.data()));
        else check(slas_cuda_ddot_norms(len_v<A>, ptr_of(a), ptr_of(b), out.data()));
        return out;
    }
    // out = a / norm(a), a untouched (normalising the borrowed side of a StaticCowVec without the copy)
    template <class A, class O> void normalize_into(const A &a, O &out) const {
        static_assert(len_v<A> == len_v<O>, "normalize_into: vectors of different static lengths");
        if constexpr (std::is_same_v<elem_t<A>, float>) check(slas_cuda_snormalize_into(len_v<A>, ptr_of(a), mut_ptr_of(out)));
        else check(slas_cuda_dnormalize_into(len_v<A>, ptr_of(a), mut_ptr_of(out)));
    }
    // Transpose (src/backends.rs:152-154): any Copy element
    template <class A> void transpose_inplace(A &a, std::size_t columns) const {
        check(slas_cuda_transpose_inplace(len_v<A>, sizeof(elem_t<A>), mut_ptr_of(a), columns));
    }
    template <class A, class B> void transpose(const A &a, B &buffer, std::size_t columns) const {
        static_assert(len_v<A> == len_v<B>, "transpose: buffer of a different static length");
        check(slas_cuda_transpose(len_v<A>, sizeof(elem_t<A>), ptr_of(a), mut_ptr_of(buffer), columns));
    }
    // MatrixMul (src/backends.rs:132-150)
    template <class A, class B, class C>
    void matrix_mul(const A &a, const B &b, C &buffer, std::size_t m, std::size_t n, std::size_t k, std::size_t lda,
                    std::size_t ldb, std::size_t ldc, bool a_trans, bool b_trans) const {
        if constexpr (std::is_same_v<elem_t<A>, float>)
            check(slas_cuda_sgemm(ptr_of(a), ptr_of(b), mut_ptr_of(buffer), m, n, k, lda, ldb, ldc, a_trans, b_trans));
        else check(slas_cuda_dgemm(ptr_of(a), ptr_of(b), mut_ptr_of(buffer), m, n, k, lda, ldb, ldc, a_trans, b_trans));
    }
    template <class A, class B, class C>
    void matrix_vector_mul(const A &a, const B &b, C &buffer, std::size_t m, std::size_t n, std::size_t lda, bool a_trans) const {
        if constexpr (std::is_same_v<elem_t<A>, float>) check(slas_cuda_sgemv(ptr_of(a), ptr_of(b), mut_ptr_of(buffer), m, n, lda, a_trans));
        else check(slas_cuda_dgemv(ptr_of(a), ptr_of(b), mut_ptr_of(buffer), m, n, lda, a_trans));
    }
    // -- process-level knobs (not trait methods: a backend is a ZST, the state lives in the library) --------------------
    // Drive `n` GPUs from this one process (0 = all visible; same as SLAS_CUDA_DEVICES=n): host operands are split over
    // them, device operands run where they live.  Returns the number of active devices.
    static int init_devices(int n = 0) {
        check(slas_cuda_init_devices(n));
        int nd = 1, px = 0;
        check(slas_cuda_devices(&nd, &px));
        return nd;
    }
    // Page-lock a long-lived host allocation in place: its operands are then read by the DMA engines directly instead
    // of through the library's bounce ring (unregister before freeing it).
    static void host_register(void *p, std::size_t bytes) { check(slas_cuda_host_register(p, bytes)); }
    static void host_unregister(void *p) { check(slas_cuda_host_unregister(p)); }
    // 0 = exact fp32 FMA chains for every f32 matrix_mul, 1 = tcgen05 3xTF32 wherever it applies, 2 = auto (default:
    // 3xTF32 from 2^27 multiply-adds -- an fp32-class result, not the ascending-k fp32 sum of Rust / Blas)
    static void set_sgemm_path(int path) { check(slas_cuda_set_sgemm_path(path)); }
    // dot / norm of ONE vector whose slices live on different GPUs of this process (slice i on its own device): the
    // exchange runs inside the reduction kernels over NVLink peer memory
    template <class T> static T dot_multi(const std::vector<std::pair<const T *, std::size_t>> &a, const std::vector<std::pair<const T *, std::size_t>> &b) {
        static_assert(std::is_same_v<T, float> || std::is_same_v<T, double>, "dot_multi: f32 / f64");
        if (a.size() != b.size()) throw Panic("dot_multi: different numbers of slices");
        std::vector<std::size_t> lens;
        std::vector<const T *> pa, pb;
        for (std::size_t i = 0; i < a.size(); ++i) {
            if (a[i].second != b[i].second) throw Panic("dot_multi: slices of different lengths");
            lens.push_back(a[i].second); pa.push_back(a[i].first); pb.push_back(b[i].first);
        }
        T out{};
        if constexpr (std::is_same_v<T, float>) check(slas_cuda_sdot_multi((int)a.size(), lens.data(), pa.data(), pb.data(), &out));
        else check(slas_cuda_ddot_multi((int)a.size(), lens.data(), pa.data(), pb.data(), &out));
        return out;
    }
};

}  // namespace backends

// ---- WithStaticBackend (src/backends.rs:185-249, 306-331) ---------------------------------------------------------
template <class U, class B = backends::Cuda> struct WithStaticBackend {
    U data;
    B backend{};
    template <class U2> auto dot(const WithStaticBackend<U2, B> &other) const { return backend.dot(data, other.data); }
    auto norm() const { return backend.norm(data); }
    void normalize() { backend.normalize(data); }
};
template <class U, class B> struct static_vec_traits<WithStaticBackend<U, B>> {
    using value_type = elem_t<U>;
    static constexpr std::size_t LEN = len_v<U>;
    static const value_type *ptr(const WithStaticBackend<U, B> &v) { return ptr_of(v.data); }
    static value_type *mut_ptr(WithStaticBackend<U, B> &v) { return mut_ptr_of(v.data); }
};
template <class B = backends::Cuda, class U> WithStaticBackend<U, B> static_backend(U data) { return {std::move(data), B{}}; }

// ---- MatrixShape / Matrix (src/tensor.rs) ------------------------------------------------------------------------------
template <std::size_t M, std::size_t K> struct MatrixShape {
    static constexpr std::size_t axis_len(std::size_t n) {
        if (n == 0) return K;  // axis 0 is the fastest axis: the width
        if (n == 1) return M;
        throw Panic("Cannot get len of axis higher than 1, as a matrix only has 2 axies (rows and columns)");
    }
    static constexpr std::size_t volume() { return M * K; }
    static constexpr std::array<std::size_t, 2> slice() { return {K, M}; }
};

// src/tensor.rs:202-219
template <std::size_t NDIM>
std::size_t tensor_index(const std::array<std::size_t, NDIM> &shape, const std::array<std::size_t, NDIM> &index) {
    std::size_t sum = 0, product = 1;
    for (std::size_t n = 0; n < NDIM; ++n) {
        SLAS_ASSERT(index[n] < shape[n], "Index out of bounds");
        sum += index[n] * product;
        product *= shape[n];
    }
    return sum;
}

template <class U, class B, std::size_t M, std::size_t K, bool IS_TRANS = false> class Matrix {
  public:
    using T = elem_t<U>;
    static constexpr std::size_t LEN = len_v<U>;
    using Shape = MatrixShape<M, K>;
    explicit Matrix(U &data) : data_(&data) {}
    static constexpr std::size_t rows() { return IS_TRANS ? Shape::axis_len(0) : Shape::axis_len(1); }     // tensor.rs:495-501
    static constexpr std::size_t columns() { return IS_TRANS ? Shape::axis_len(1) : Shape::axis_len(0); }  // tensor.rs:504-510
    Matrix<U, B, M, K, !IS_TRANS> transpose() const { return Matrix<U, B, M, K, !IS_TRANS>(*data_); }      // flag flip, no data movement
    Matrix<U, B, M, K, !IS_TRANS> as_transposed() const { return transpose(); }
    // Matrix[(r, c)] (tensor.rs:296-301)
    T &operator()(std::size_t r, std::size_t c) {
        const std::array<std::size_t, 2> i = IS_TRANS ? std::array<std::size_t, 2>{r, c} : std::array<std::size_t, 2>{c, r};
        return mut_ptr_of(*data_)[tensor_index<2>(Shape::slice(), i)];
    }
    // matrix_mul_buffer (tensor.rs:364-407): derive (m, n, k, lda, ldb, ldc), one backend call
    template <class U2, std::size_t M2, std::size_t K2, bool T2, class U3>
    void matrix_mul_buffer(const Matrix<U2, B, M2, K2, T2> &other, U3 &buffer) const {
        constexpr std::size_t m = rows(), k = Matrix<U2, B, M2, K2, T2>::rows(), n = Matrix<U2, B, M2, K2, T2>::columns();
        constexpr std::size_t lda = Shape::axis_len(0), ldb = MatrixShape<M2, K2>::axis_len(0), ldc = n;
        SLAS_ASSERT(m * n == len_v<U3>, "Matrix::matrix_mul_buffer expected buffer of " + std::to_string(m * n) +
                                            " elements, found one of " + std::to_string(len_v<U3>));
        B{}.matrix_mul(*data_, other.data(), buffer, m, n, k, lda, ldb, ldc, IS_TRANS, T2);
    }
    template <class U2, std::size_t M2, std::size_t K2, bool T2> auto matrix_mul(const Matrix<U2, B, M2, K2, T2> &other) const {
        std::array<T, rows() * Matrix<U2, B, M2, K2, T2>::columns()> buffer{};  // [T::_0; OLEN] (tensor.rs:451)
        matrix_mul_buffer(other, buffer);
        return buffer;
    }
    // vector_mul_buffer (tensor.rs:410-438): the trait's m = width, n = height
    template <class U2, class U3> void vector_mul_buffer(const U2 &other, U3 &buffer) const {
        SLAS_ASSERT(rows() == len_v<U3>, "Matrix::vector_mul_buffer expected buffer of " + std::to_string(rows()) + " elements");
        SLAS_ASSERT(len_v<U2> == columns(), "Matrix::vector_mul: vector length differs from the matrix width");
        B{}.matrix_vector_mul(*data_, other, buffer, Shape::axis_len(0), Shape::axis_len(1), Shape::axis_len(0), IS_TRANS);
    }
    template <class U2> auto vector_mul(const U2 &other) const {
        std::array<T, rows()> buffer{};
        vector_mul_buffer(other, buffer);
        return buffer;
    }
    const U &data() const { return *data_; }
    U &vec_ref() { return *data_; }

  private:
    U *data_;
};

// `.matrix::<B, M, K>()` (src/static_vec.rs:62-80): M * K must equal LEN
template <class B, std::size_t M, std::size_t K, class U> Matrix<U, B, M, K> matrix(U &data) {
    SLAS_ASSERT(M * K == len_v<U>, "Cannot reshape vector of " + std::to_string(len_v<U>) + " elements as matrix of " + std::to_string(M * K));
    return Matrix<U, B, M, K>(data);
}

// ---- batched views: `[[T; LEN]; batch]` contiguous, the static object shape as template parameters -------------------
// north_star's "batched view over many same-shape objects": what a slas user holds as a slice of static vectors /
// matrices and loops over today (a loop of Matrix::matrix_mul calls, src/tensor.rs:441-454).  Storage is either
// device-resident (CudaBatch) or a host buffer (HostBatch: staged by the library's H2D / kernel / D2H pipeline).
template <class T, std::size_t OBJ_LEN> class CudaBatch {
  public:
    using value_type = T;
    static constexpr std::size_t LEN = OBJ_LEN;
    explicit CudaBatch(std::size_t batch) : batch_(batch) { check(slas_cuda_malloc(reinterpret_cast<void **>(&p_), bytes())); }
    CudaBatch(std::size_t batch, const T *host) : CudaBatch(batch) { check(slas_cuda_memcpy_h2d(p_, host, bytes())); }
    CudaBatch(const CudaBatch &) = delete;
    CudaBatch &operator=(const CudaBatch &) = delete;
    CudaBatch(CudaBatch &&o) noexcept : p_(o.p_), batch_(o.batch_) { o.p_ = nullptr; }
    ~CudaBatch() { slas_cuda_free(p_); }
    std::size_t batch() const { return batch_; }
    std::size_t bytes() const { return batch_ * OBJ_LEN * sizeof(T); }
    const T *as_ptr() const { return p_; }
    T *as_mut_ptr() { return p_; }
    void to_host(T *out) const { check(slas_cuda_memcpy_d2h(out, p_, bytes())); }

  private:
    T *p_ = nullptr;
    std::size_t batch_;
};
template <class T, std::size_t OBJ_LEN> class HostBatch {  // non-owning view of caller memory
  public:
    using value_type = T;
    static constexpr std::size_t LEN = OBJ_LEN;
    HostBatch(T *data, std::size_t batch) : p_(data), batch_(batch) {}
    std::size_t batch() const { return batch_; }
    const T *as_ptr() const { return p_; }
    T *as_mut_ptr() { return p_; }

  private:
    T *p_;
    std::size_t batch_;
};

// Batch of dense M x K matrices with the lazy-transpose flag of Matrix (tensor.rs:512-522): the flag is passed to the
// batched kernel as a_trans / b_trans, never materialised.
template <class S, class B, std::size_t M, std::size_t K, bool IS_TRANS = false> class BatchedMatrix {
  public:
    using T = typename S::value_type;
    static_assert(S::LEN == M * K, "Cannot reshape the objects of this batch as M x K matrices");
    explicit BatchedMatrix(S &storage) : s_(&storage) {}
    static constexpr std::size_t rows() { return IS_TRANS ? K : M; }
    static constexpr std::size_t columns() { return IS_TRANS ? M : K; }
    std::size_t batch() const { return s_->batch(); }
    BatchedMatrix<S, B, M, K, !IS_TRANS> transpose() const { return BatchedMatrix<S, B, M, K, !IS_TRANS>(*s_); }
    const S &storage() const { return *s_; }
    // C_i = op(A_i) . op(B_i) for every i: one batched kernel launch
    template <class S2, std::size_t M2, std::size_t K2, bool T2, class S3>
    void matrix_mul_buffer(const BatchedMatrix<S2, B, M2, K2, T2> &other, S3 &buffer) const {
        constexpr std::size_t m = rows(), k = BatchedMatrix<S2, B, M2, K2, T2>::rows(), n = BatchedMatrix<S2, B, M2, K2, T2>::columns();
        static_assert(columns() == k, "BatchedMatrix::matrix_mul: inner dimensions differ");
        static_assert(S3::LEN == m * n, "BatchedMatrix::matrix_mul_buffer: buffer objects of the wrong length");
        SLAS_ASSERT(other.batch() == batch() && buffer.batch() == batch(), "BatchedMatrix::matrix_mul: batches differ");
        if constexpr (std::is_same_v<T, float>)
            check(slas_cuda_sgemm_batched(batch(), s_->as_ptr(), other.storage().as_ptr(), buffer.as_mut_ptr(), m, n, k, IS_TRANS, T2));
        else
            check(slas_cuda_dgemm_batched(batch(), s_->as_ptr(), other.storage().as_ptr(), buffer.as_mut_ptr(), m, n, k, IS_TRANS, T2));
    }
    // y_i = op(A_i) . x_i; the C ABI takes the trait's m = width, n = height (tensor.rs:428-437)
    template <class S2, class S3> void vector_mul_buffer(const S2 &x, S3 &y) const {
        static_assert(S2::LEN == columns() && S3::LEN == rows(), "BatchedMatrix::vector_mul: vector lengths differ from the matrix");
        SLAS_ASSERT(x.batch() == batch() && y.batch() == batch(), "BatchedMatrix::vector_mul: batches differ");
        if constexpr (std::is_same_v<T, float>) check(slas_cuda_sgemv_batched(batch(), s_->as_ptr(), x.as_ptr(), y.as_mut_ptr(), K, M, IS_TRANS));
        else check(slas_cuda_dgemv_batched(batch(), s_->as_ptr(), x.as_ptr(), y.as_mut_ptr(), K, M, IS_TRANS));
    }
    // every object transposed out of place / in place (Matrix::deref_mut per object, tensor.rs:546-564)
    template <class S3> void materialize_transpose(S3 &buffer) const {
        static_assert(S3::LEN == M * K, "BatchedMatrix::materialize_transpose: buffer objects of the wrong length");
        check(slas_cuda_transpose_batched(batch(), M * K, sizeof(T), s_->as_ptr(), buffer.as_mut_ptr(), M));
    }
    void materialize_transpose_inplace() { check(slas_cuda_transpose_inplace_batched(batch(), M * K, sizeof(T), s_->as_mut_ptr(), M)); }

  private:
    S *s_;
};
template <class B, std::size_t M, std::size_t K, class S> BatchedMatrix<S, B, M, K> batched_matrix(S &storage) {
    return BatchedMatrix<S, B, M, K>(storage);
}

// batched DotProduct / Normalize over the objects of a batch
template <class S, class O> void dot_batched(const S &a, const S &b, O &out) {
    using T = typename S::value_type;
    static_assert(O::LEN == 1, "dot_batched: one result per object");
    SLAS_ASSERT(a.batch() == b.batch() && out.batch() == a.batch(), "dot_batched: batches differ");
    if constexpr (std::is_same_v<T, float>) check(slas_cuda_sdot_batched(a.batch(), S::LEN, a.as_ptr(), b.as_ptr(), out.as_mut_ptr()));
    else check(slas_cuda_ddot_batched(a.batch(), S::LEN, a.as_ptr(), b.as_ptr(), out.as_mut_ptr()));
}
template <class S, class O> void norm_batched(const S &a, O &out) {
    using T = typename S::value_type;
    static_assert(O::LEN == 1, "norm_batched: one result per object");
    SLAS_ASSERT(out.batch() == a.batch(), "norm_batched: batches differ");
    if constexpr (std::is_same_v<T, float>) check(slas_cuda_snrm2_batched(a.batch(), S::LEN, a.as_ptr(), out.as_mut_ptr()));
    else check(slas_cuda_dnrm2_batched(a.batch(), S::LEN, a.as_ptr(), out.as_mut_ptr()));
}
template <class S> void normalize_batched(S &a) {
    using T = typename S::value_type;
    if constexpr (std::is_same_v<T, float>) check(slas_cuda_snormalize_batched(a.batch(), S::LEN, a.as_mut_ptr()));
    else check(slas_cuda_dnormalize_batched(a.batch(), S::LEN, a.as_mut_ptr()));
}

namespace slas_backend = backends;

}  // namespace slas
