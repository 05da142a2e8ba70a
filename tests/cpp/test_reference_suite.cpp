// test_reference_suite.cpp -- the reference's own hot-path tests (tests/src/main.rs, doc-tests) replayed
// against `slas_backend::Cuda` through the typed C++ host mirror (include/slas_cuda.hpp -> C ABI ->
// CUDA kernels).  Each case cites the reference test it mirrors; expectations are the reference's own
// literals (exact equality where the reference uses assert_eq!).  Needs a GPU: built by
// __graft_entry__.build(), run by tests/test_cpp_reference_suite.py under `pytest -m gpu`.
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "slas_cuda.hpp"

using namespace slas;
using slas_backend::Cuda;

static int g_failed = 0, g_run = 0;
#define CHECK(cond)                                                                      \
    do {                                                                                 \
        if (!(cond)) { std::printf("    FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); throw 1; } \
    } while (0)

static void run(const char *name, const std::function<void()> &body, bool should_panic = false) {
    ++g_run;
    bool panicked = false, failed = false;
    try { body(); } catch (const Panic &p) { panicked = true; if (!should_panic) std::printf("    panic: %s\n", p.what()); } catch (int) { failed = true; }
    const bool ok = !failed && (panicked == should_panic);
    std::printf("test %s ... %s\n", name, ok ? "ok" : "FAILED");
    if (!ok) ++g_failed;
}

template <class T, std::size_t N> static bool eq(const std::array<T, N> &a, std::initializer_list<T> b) {
    std::size_t i = 0;
    for (T v : b) if (a[i++] != v) return false;
    return i == N;
}

// a crash inside the library must name its frame in the pytest log (the box has no debugger)
static void on_fatal_signal(int sig) {
    void *frames[64];
    const int n = backtrace(frames, 64);
    dprintf(2, "fatal signal %d, backtrace:\n", sig);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

int main() {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    for (int sig : {SIGFPE, SIGSEGV, SIGBUS, SIGILL}) signal(sig, on_fatal_signal);
    check(slas_cuda_init(-1));

    // ---- mod thin_blas (tests/src/main.rs:37-96), on Cuda instead of Blas ----
    run("thin_blas::casting_and_dot (main.rs:41-46)", [] {
        std::array<float, 4> a{0, 1, 2, 3}, b{0, -1, -2, 3};
        CHECK(Cuda{}.dot(a, b) == 4.f);
    });
    run("thin_blas::static_backend (main.rs:70-78)", [] {
        auto a = static_backend<Cuda>(moo_range<float, 4>(0));
        auto b = static_backend<Cuda>(std::array<float, 4>{1, 2, 3, 4});
        CHECK(a.dot(b) == 20.f);
    });
    run("thin_blas::static_backend_macro (main.rs:81-95)", [] {
        auto a = static_backend<Cuda>(std::array<float, 4>{1, 2, 3, 4});
        auto b = static_backend<Cuda>(moo_range<float, 4>(0));
        CHECK(a.dot(b) == 20.f && b.dot(a) == 20.f);
    });

    // ---- mod moo (tests/src/main.rs:130-250): the dot / norm cases ----
    run("moo::norm_complex_2d (main.rs:133-143)", [] {
        std::array<std::complex<float>, 2> a{std::complex<float>{1.2f, 2.3f}, std::complex<float>{1.2f, 2.3f}};
        CHECK(static_backend<Cuda>(a).norm() == 3.6687872f);
    });
    run("moo::dot (main.rs:146-148)", [] {
        std::array<float, 3> a{1, 2, 3};
        CHECK(Cuda{}.dot(a, a) == 14.f);
    });
    run("moo::dot_rust_backend f64 (main.rs:151-160)", [] {
        std::array<double, 4> a{1, 2, 3, 4};
        std::array<double, 5> b{1, 2, 3, 4, 5};
        CHECK(Cuda{}.dot(a, a) == 30.0 && Cuda{}.dot(b, b) == 55.0);
    });
    run("moo::complex dotu (main.rs:196-201)", [] {
        std::array<std::complex<float>, 5> a;
        a.fill({1.f, 2.f});
        CHECK(Cuda{}.dot(a, a) == std::complex<float>(-15.f, 20.f));  // unconjugated
    });
    run("doc: rust.rs:16-19 / blas.rs:11-14", [] {
        std::array<float, 3> a{1, 2, 3}, b{-1, 2, -1};
        CHECK(Cuda{}.dot(a, b) == 0.f);
    });
    run("doc: backends.rs:267-272", [] {
        auto a = moo_range<float, 4>(0);
        std::array<float, 4> b{1.2f, 1.2f, 1.2f, 1.2f};
        CHECK(std::fabs(Cuda{}.dot(a, b) - 7.2f) < 0.000003f);
    });
    run("rust.rs:97-110 basic_{add,sub,mul,div}_{f32,f64}", [] {
        auto a = moo_range<float, 12>(1);
        auto d = moo_range<double, 12>(1);
        std::array<float, 12> c{};
        std::array<double, 12> e{};
        Cuda{}.add(a, a, c); for (int n = 0; n < 12; ++n) CHECK(c[n] == a[n] + a[n]);
        Cuda{}.sub(a, a, c); for (int n = 0; n < 12; ++n) CHECK(c[n] == a[n] - a[n]);
        Cuda{}.mul(a, a, c); for (int n = 0; n < 12; ++n) CHECK(c[n] == a[n] * a[n]);
        Cuda{}.div(a, a, c); for (int n = 0; n < 12; ++n) CHECK(c[n] == a[n] / a[n]);
        Cuda{}.add(d, d, e); for (int n = 0; n < 12; ++n) CHECK(e[n] == d[n] + d[n]);
        Cuda{}.mul(d, d, e); for (int n = 0; n < 12; ++n) CHECK(e[n] == d[n] * d[n]);
        Cuda{}.div(d, d, e); for (int n = 0; n < 12; ++n) CHECK(e[n] == d[n] / d[n]);
        Cuda{}.sub(d, d, e); for (int n = 0; n < 12; ++n) CHECK(e[n] == d[n] - d[n]);
    });

    // ---- mod tensors (tests/src/main.rs:253-582): matrix_mul, vector_mul, transposes, panics ----
    run("tensors::matrix_mul (main.rs:289-297, lib.rs:135-140)", [] {
        auto d = moo_range<float, 6>(1);
        auto a = matrix<Cuda, 2, 3>(d);
        auto b = matrix<Cuda, 3, 2>(d);
        CHECK(eq(a.matrix_mul(b), {22.f, 28.f, 49.f, 64.f}));
    });
    run("tensors::matrix_mul_trans (main.rs:300-308)", [] {
        auto d = moo_range<float, 6>(1);
        auto a = matrix<Cuda, 2, 3>(d);
        auto b = matrix<Cuda, 3, 2>(d);
        CHECK(eq(a.transpose().matrix_mul(b.transpose()), {9.f, 19.f, 29.f, 12.f, 26.f, 40.f, 15.f, 33.f, 51.f}));
    });
    run("tensors::matrix_mul_trans_b (main.rs:311-319)", [] {
        auto d = moo_range<float, 6>(1);
        auto a = matrix<Cuda, 3, 2>(d);
        auto b = matrix<Cuda, 3, 2>(d);
        CHECK(eq(a.matrix_mul(b.transpose()), {5.f, 11.f, 17.f, 11.f, 25.f, 39.f, 17.f, 39.f, 61.f}));
    });
    run("tensors::matrix_mul_trans_a (main.rs:322-330)", [] {
        auto d = moo_range<float, 6>(1);
        auto a = matrix<Cuda, 2, 3>(d);
        auto b = matrix<Cuda, 2, 3>(d);
        CHECK(eq(a.transpose().matrix_mul(b), {17.f, 22.f, 27.f, 22.f, 29.f, 36.f, 27.f, 36.f, 45.f}));
    });
    run("tensors::matrix_mul_trans_a2 (main.rs:333-341)", [] {
        auto d = moo_range<float, 6>(1);
        auto a = matrix<Cuda, 2, 3>(d);
        auto b = matrix<Cuda, 2, 3>(d);
        CHECK(eq(a.matrix_mul(b.transpose()), {14.f, 32.f, 32.f, 77.f}));
    });
    run("tensors::matrix_mul_trans_b2 (main.rs:344-352)", [] {
        auto d = moo_range<float, 6>(1);
        auto a = matrix<Cuda, 3, 2>(d);
        auto b = matrix<Cuda, 3, 2>(d);
        CHECK(eq(a.transpose().matrix_mul(b), {35.f, 44.f, 44.f, 56.f}));
    });
    run("tensors::matrix_mul f64 + device-resident operands", [] {
        auto d = moo_range<double, 6>(1);
        CudaVec<double, 6> dev(d);
        auto a = matrix<Cuda, 2, 3>(dev);
        auto b = matrix<Cuda, 3, 2>(dev);
        CudaVec<double, 4> out;
        a.matrix_mul_buffer(b, out);
        CHECK(eq(out.to_host(), {22.0, 28.0, 49.0, 64.0}));
    });
    run("doc: lib.rs:138-141 vector_mul", [] {
        auto d = moo_range<float, 6>(1);
        std::array<float, 2> x{1, 2};
        CHECK(eq(matrix<Cuda, 3, 2>(d).vector_mul(x), {5.f, 11.f, 17.f}));
    });
    run("tensors::matrix_vector_mul (main.rs:355-366)", [] {
        auto d = moo_range<float, 6>(1);
        auto v = moo_range<float, 3>(1);
        auto a = matrix<Cuda, 2, 3>(d);
        auto b = matrix<Cuda, 3, 1>(v);
        CHECK(a.matrix_mul(b) == a.vector_mul(v));
    });
    run("tensors::matrix_vector_mul_trans (main.rs:369-380)", [] {
        auto d = moo_range<float, 6>(1);
        auto v = moo_range<float, 3>(1);
        auto a = matrix<Cuda, 3, 2>(d);
        auto b = matrix<Cuda, 3, 1>(v);
        CHECK(a.transpose().matrix_mul(b) == a.transpose().vector_mul(v));
    });
    run("tensors::matrix_vector_mul_trans_2 (main.rs:383-394)", [] {
        auto d = moo_range<float, 12>(2);
        auto v = moo_range<float, 4>(2);
        auto a = matrix<Cuda, 4, 3>(d);
        auto b = matrix<Cuda, 4, 1>(v);
        CHECK(a.transpose().matrix_mul(b) == a.transpose().vector_mul(v));
        CHECK(eq(a.transpose().vector_mul(v), {106.f, 120.f, 134.f}));
    });
    run("tensors::trans_matrix (main.rs:397-415)", [] {
        auto d = moo_range<float, 6>(1);
        auto m = matrix<Cuda, 2, 3>(d);
        CHECK(m.rows() == 2 && m.columns() == 3);
        CHECK(m(1, 0) == m.as_transposed()(0, 1));
        CHECK(m(0, 2) == m.as_transposed()(2, 0));
        m.as_transposed()(0, 1) = 0.f;
        CHECK(m(1, 0) == 0.f);
        auto n = m.transpose();
        CHECK(n(0, 1) == 0.f && n.rows() == 3 && n.columns() == 2);
    });
    run("tensors::rust_transpose (main.rs:443-453)", [] {
        std::array<double, 6> a{1, 4, 2, 5, 3, 6}, b{};
        Cuda{}.transpose(a, b, 3);
        CHECK(eq(b, {1.0, 2.0, 3.0, 4.0, 5.0, 6.0}));
    });
    run("tensors::trans_fixed_deref (main.rs:427-440): materialised transpose", [] {
        auto a = moo_range<float, 6>(1);  // 1..=6 shaped [3, 2], lazily transposed -> transpose_inplace(columns = axis_len(1) = 2)
        Cuda{}.transpose_inplace(a, 2);
        CHECK(eq(a, {1.f, 4.f, 2.f, 5.f, 3.f, 6.f}));
    });
    run("tensors::wrong_size (main.rs:455-469) [should panic]", [] {
        auto d = moo_range<float, 5>(1);
        auto a = matrix<Cuda, 2, 3>(d);
        (void)a;
    }, /*should_panic=*/true);
    run("tensors::wrong buffer size (tensor.rs:384-391) [should panic]", [] {
        auto d = moo_range<float, 6>(1);
        std::array<float, 5> out{};
        matrix<Cuda, 2, 3>(d).matrix_mul_buffer(matrix<Cuda, 3, 2>(d), out);
    }, true);
    run("tensors::index out of bounds (main.rs:479-485) [should panic]", [] {
        auto d = moo_range<float, 6>(1);
        (void)matrix<Cuda, 2, 3>(d)(2, 0);
    }, true);
    run("doc: static_vec.rs:83-90 zeros", [] {
        auto a = moo_range<float, 6>(0);
        std::array<float, 6> z{};
        CHECK(eq(matrix<Cuda, 2, 3>(a).matrix_mul(matrix<Cuda, 3, 2>(z)), {0.f, 0.f, 0.f, 0.f}));
    });
    run("normalize is a true division (rust.rs:123-126)", [] {
        std::array<float, 4> a{3, 4, 0, 12};
        const float n = Cuda{}.norm(a);
        CHECK(n == 13.f);
        auto b = a;
        Cuda{}.normalize(b);
        for (int i = 0; i < 4; ++i) CHECK(b[i] == a[i] / n);
    });

    // ---- batched views: the matrix cases above over [[T; LEN]; batch], device-resident and host storage ----
    run("batched: matrix_multiplication (main.rs:289-301) x 1000 objects, lazy transposes, gemv, transposes", [] {
        constexpr std::size_t batch = 1000;
        std::vector<float> a(batch * 6), b(batch * 6), c(batch * 4), c2(batch * 4);
        for (std::size_t i = 0; i < batch; ++i)
            for (int j = 0; j < 6; ++j) { a[i * 6 + j] = float(j + 1) + float(i % 7); b[i * 6 + j] = float(j) - float(i % 5); }
        CudaBatch<float, 6> da(batch, a.data()), db(batch, b.data());
        CudaBatch<float, 4> dc(batch);
        auto A = batched_matrix<Cuda, 2, 3>(da);
        auto Bm = batched_matrix<Cuda, 3, 2>(db);
        A.matrix_mul_buffer(Bm, dc);
        dc.to_host(c.data());
        HostBatch<float, 6> ha(a.data(), batch), hb(b.data(), batch);  // the same product through host pointers
        HostBatch<float, 4> hc(c2.data(), batch);
        batched_matrix<Cuda, 2, 3>(ha).matrix_mul_buffer(batched_matrix<Cuda, 3, 2>(hb), hc);
        for (std::size_t i = 0; i < batch; ++i) {
            std::array<float, 6> ai{}, bi{};
            for (int j = 0; j < 6; ++j) { ai[j] = a[i * 6 + j]; bi[j] = b[i * 6 + j]; }
            const auto want = matrix<Cuda, 2, 3>(ai).matrix_mul(matrix<Cuda, 3, 2>(bi));  // small integers: exact
            for (int j = 0; j < 4; ++j) CHECK(c[i * 4 + j] == want[j] && c2[i * 4 + j] == want[j]);
        }
        // A^T (3 x 2) . A (2 x 3) -> 3 x 3 with the lazy flag on the left operand
        CudaBatch<float, 9> dg(batch);
        A.transpose().matrix_mul_buffer(A, dg);
        std::vector<float> g(batch * 9);
        dg.to_host(g.data());
        for (std::size_t i = 0; i < batch; i += 97)
            for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 3; ++q)
                    CHECK(g[i * 9 + r * 3 + q] == a[i * 6 + r] * a[i * 6 + q] + a[i * 6 + 3 + r] * a[i * 6 + 3 + q]);
        // y = A x per object
        std::vector<float> x(batch * 3, 1.f), y(batch * 2);
        CudaBatch<float, 3> dx(batch, x.data());
        CudaBatch<float, 2> dy(batch);
        A.vector_mul_buffer(dx, dy);
        dy.to_host(y.data());
        for (std::size_t i = 0; i < batch; i += 91) CHECK(y[i * 2] == a[i * 6] + a[i * 6 + 1] + a[i * 6 + 2]);
        // materialised transposes, out of place and in place
        CudaBatch<float, 6> dt(batch);
        A.materialize_transpose(dt);
        A.materialize_transpose_inplace();
        std::vector<float> t1(batch * 6), t2(batch * 6);
        dt.to_host(t1.data());
        da.to_host(t2.data());
        for (std::size_t i = 0; i < batch; i += 89)
            for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 2; ++q) CHECK(t1[i * 6 + r * 2 + q] == a[i * 6 + q * 3 + r] && t2[i * 6 + r * 2 + q] == t1[i * 6 + r * 2 + q]);
    });
    run("batched: dot / norm / normalize per object", [] {
        constexpr std::size_t batch = 513;
        std::vector<double> a(batch * 4), b(batch * 4);
        for (std::size_t i = 0; i < batch; ++i) {
            const double s = double(i % 9 + 1);
            a[i * 4] = 3 * s; a[i * 4 + 1] = 4 * s; a[i * 4 + 2] = 0; a[i * 4 + 3] = 12 * s;
            b[i * 4] = 1; b[i * 4 + 1] = 0; b[i * 4 + 2] = 5; b[i * 4 + 3] = 1;
        }
        CudaBatch<double, 4> da(batch, a.data()), db(batch, b.data());
        CudaBatch<double, 1> dd(batch), dn(batch);
        dot_batched(da, db, dd);
        norm_batched(da, dn);
        normalize_batched(da);
        std::vector<double> d(batch), n(batch), u(batch * 4);
        dd.to_host(d.data()); dn.to_host(n.data()); da.to_host(u.data());
        for (std::size_t i = 0; i < batch; ++i) {
            const double s = double(i % 9 + 1);
            CHECK(d[i] == 15 * s && n[i] == 13 * s);
            for (int j = 0; j < 4; ++j) CHECK(u[i * 4 + j] == a[i * 4 + j] / n[i]);
        }
    });

    // ---- StaticCowVec: copy on first mutable access (src/lib.rs:433-456, static_vec.rs:294-319) ----
    run("moo::copy on write (lib.rs:255-268 doc example, on Cuda)", [] {
        std::array<float, 3> source{1, 2, 3};
        StaticCowVec<float, 3> v(source);
        CHECK(v.is_borrowed() && v.as_ptr() == source.data());
        CHECK(Cuda{}.dot(v, source) == 14.f && v.is_borrowed());  // reads never copy
        Cuda{}.normalize(v);                                       // mutable access: copies, the source is untouched
        CHECK(v.is_owned() && source[2] == 3.f);
        const float n = Cuda{}.norm(source);
        for (int i = 0; i < 3; ++i) CHECK(v[i] == source[i] / n);
    });
    run("moo::mutations (main.rs:168-177)", [] {
        const std::array<float, 3> src{3, 2, 3};
        StaticCowVec<float, 3> t(src);
        CHECK(t.is_borrowed());
        t.as_mut_ptr()[0] = 1.f;  // IndexMut -> deref_mut -> copy
        CHECK(t.is_owned() && t[0] == 1.f && t[1] == 2.f && t[2] == 3.f && src[0] == 3.f);
        t.as_mut_ptr()[0] = 0.f;
        CHECK(t[0] == 0.f && t[1] == 2.f && t[2] == 3.f);
    });
    run("moo::dot with a StaticCowVec operand (main.rs:189-193)", [] {
        const auto a = moo_range<float, 4>(0);
        const std::array<float, 4> bs{0, -1, -2, 3};
        StaticCowVec<float, 4> b(bs);
        CHECK(Cuda{}.dot(a, b) == 4.f && b.is_borrowed());
    });
    run("moo::from_readme (main.rs:214-226)", [] {
        std::array<float, 3> source{1, 2, 3};
        StaticCowVec<float, 3> v(source);
        source[1] = 3.f;          // v still borrows: it sees this
        v.as_mut_ptr()[0] = 0.f;  // first mutable access: v copies [1, 3, 3] and becomes owned
        source[2] = 4.f;          // ... so it does not see this
        CHECK(v[0] == 0.f && v[1] == 3.f && v[2] == 3.f);
        CHECK(source[0] == 1.f && source[1] == 3.f && source[2] == 4.f);
        CHECK(Cuda{}.dot(v, source) == 0.f * 1 + 3.f * 3 + 3.f * 4);
    });

    // ---- fused chains and the 1-/2-byte transposes: same answers as the op sequences of the cases above ----
    run("fused: add -> dot equals the chain (main.rs:41-46 operands)", [] {
        std::array<float, 4> a{0, 1, 2, 3}, b{0, -1, -2, 3}, d{1, 2, 3, 4}, c{}, t{};
        Cuda{}.add(a, b, t);
        CHECK(Cuda{}.add_dot(a, b, d) == Cuda{}.dot(t, d));
        CHECK(Cuda{}.add_dot(a, b, d, c) == 24.f && c == t);
        CHECK(Cuda{}.mul_dot(a, b, d) == 0.f * 1 - 1.f * 2 - 4.f * 3 + 9.f * 4);
    });
    run("fused: dot_norms / normalize_into", [] {
        std::array<double, 4> a{3, 4, 0, 12}, b{1, 0, 0, 1}, o{};
        const auto r = Cuda{}.dot_norms(a, b);
        CHECK(r[0] == 15.0 && r[1] == 13.0 && r[2] == std::sqrt(2.0));
        Cuda{}.normalize_into(a, o);
        for (int i = 0; i < 4; ++i) CHECK(o[i] == a[i] / 13.0);
        CHECK(a[3] == 12.0);
    });
    run("transpose of 1- and 2-byte Copy elements (rust.rs:151: any T: Copy)", [] {
        std::array<unsigned char, 6> a{1, 2, 3, 4, 5, 6}, o{};
        Cuda{}.transpose(a, o, 2);  // input 2 x 3 row-major -> 3 x 2
        CHECK(eq(o, {(unsigned char)1, (unsigned char)4, (unsigned char)2, (unsigned char)5, (unsigned char)3, (unsigned char)6}));
        std::array<short, 6> s{1, 2, 3, 4, 5, 6};
        Cuda{}.transpose_inplace(s, 2);
        CHECK(eq(s, {(short)1, (short)4, (short)2, (short)5, (short)3, (short)6}));
    });

    std::printf("\ntest result: %s. %d passed; %d failed\n", g_failed ? "FAILED" : "ok", g_run - g_failed, g_failed);
    return g_failed ? 1 : 0;
}
