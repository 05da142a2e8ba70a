//! `slas_backend::Cuda` -- the Rust side of the B200 backend.
//!
//! UNVERIFIED SOURCE: this image has no rustc/cargo (and slas needs a ~2022 nightly), so this
//! file has never been compiled.  It is the binding a slas maintainer adds next to
//! `src/backends/rust.rs` and `src/backends/blas.rs`; the arithmetic lives in libslas_cuda.so
//! (include/slas_cuda.h), which IS built and tested here through the same C ABI.
//!
//! Registration (src/backends.rs, beside lines 333-337):
//! ```ignore
//! #[cfg(feature = "cuda")] mod cuda;
//! #[cfg(feature = "cuda")] pub use cuda::{Cuda, CudaVec, CudaBatch};
//! ```
use super::*;
use operations::*;
use std::marker::PhantomData;

/// A zero-sized backend like `Rust` and `Blas` (src/backends/rust.rs:2-3): `B::default()` is
/// called at every call site, so device, stream and scratch live in the C library.
#[derive(Default, Clone, Copy)]
pub struct Cuda;

#[allow(non_camel_case_types)]
type c_int = std::os::raw::c_int;

#[link(name = "slas_cuda")]
extern "C" {
    fn slas_cuda_last_error() -> *const std::os::raw::c_char;
    fn slas_cuda_malloc(dptr: *mut *mut std::ffi::c_void, bytes: usize) -> c_int;
    fn slas_cuda_free(dptr: *mut std::ffi::c_void) -> c_int;
    fn slas_cuda_memcpy_h2d(dst: *mut std::ffi::c_void, src: *const std::ffi::c_void, bytes: usize) -> c_int;
    fn slas_cuda_memcpy_d2h(dst: *mut std::ffi::c_void, src: *const std::ffi::c_void, bytes: usize) -> c_int;

    fn slas_cuda_sdot(len: usize, a: *const f32, b: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_ddot(len: usize, a: *const f64, b: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_cdotu(len: usize, a: *const f32, b: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_zdotu(len: usize, a: *const f64, b: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_snrm2(len: usize, a: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_dnrm2(len: usize, a: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_scnrm2(len: usize, a: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_dznrm2(len: usize, a: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_snormalize(len: usize, a: *mut f32) -> c_int;
    fn slas_cuda_dnormalize(len: usize, a: *mut f64) -> c_int;
    fn slas_cuda_cnormalize(len: usize, a: *mut f32) -> c_int;
    fn slas_cuda_znormalize(len: usize, a: *mut f64) -> c_int;
    fn slas_cuda_sadd(len: usize, a: *const f32, b: *const f32, c: *mut f32) -> c_int;
    fn slas_cuda_ssub(len: usize, a: *const f32, b: *const f32, c: *mut f32) -> c_int;
    fn slas_cuda_smul(len: usize, a: *const f32, b: *const f32, c: *mut f32) -> c_int;
    fn slas_cuda_sdiv(len: usize, a: *const f32, b: *const f32, c: *mut f32) -> c_int;
    fn slas_cuda_dadd(len: usize, a: *const f64, b: *const f64, c: *mut f64) -> c_int;
    fn slas_cuda_dsub(len: usize, a: *const f64, b: *const f64, c: *mut f64) -> c_int;
    fn slas_cuda_dmul(len: usize, a: *const f64, b: *const f64, c: *mut f64) -> c_int;
    fn slas_cuda_ddiv(len: usize, a: *const f64, b: *const f64, c: *mut f64) -> c_int;
    fn slas_cuda_transpose(len: usize, elem: usize, a: *const std::ffi::c_void, buf: *mut std::ffi::c_void, columns: usize) -> c_int;
    fn slas_cuda_transpose_inplace(len: usize, elem: usize, a: *mut std::ffi::c_void, columns: usize) -> c_int;
    fn slas_cuda_sgemm(a: *const f32, b: *const f32, c: *mut f32, m: usize, n: usize, k: usize,
                       lda: usize, ldb: usize, ldc: usize, ta: c_int, tb: c_int) -> c_int;
    fn slas_cuda_dgemm(a: *const f64, b: *const f64, c: *mut f64, m: usize, n: usize, k: usize,
                       lda: usize, ldb: usize, ldc: usize, ta: c_int, tb: c_int) -> c_int;
    fn slas_cuda_sgemv(a: *const f32, x: *const f32, y: *mut f32, m: usize, n: usize, lda: usize, ta: c_int) -> c_int;
    fn slas_cuda_dgemv(a: *const f64, x: *const f64, y: *mut f64, m: usize, n: usize, lda: usize, ta: c_int) -> c_int;
    fn slas_cuda_sadd_dot(len: usize, a: *const f32, b: *const f32, c: *mut f32, d: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_dadd_dot(len: usize, a: *const f64, b: *const f64, c: *mut f64, d: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_sdot_norms(len: usize, a: *const f32, b: *const f32, out3: *mut f32) -> c_int;
    fn slas_cuda_ddot_norms(len: usize, a: *const f64, b: *const f64, out3: *mut f64) -> c_int;
    fn slas_cuda_snormalize_into(len: usize, a: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_dnormalize_into(len: usize, a: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_sgemm_batched(batch: usize, a: *const f32, b: *const f32, c: *mut f32,
                               m: usize, n: usize, k: usize, ta: c_int, tb: c_int) -> c_int;
    fn slas_cuda_dgemm_batched(batch: usize, a: *const f64, b: *const f64, c: *mut f64,
                               m: usize, n: usize, k: usize, ta: c_int, tb: c_int) -> c_int;
}

/// The reference's ops are infallible by type (src/backends.rs:121-183): a non-zero status panics,
/// the same convention as the asserts of the view layer.
#[inline(always)]
fn check(status: c_int) {
    if status != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(slas_cuda_last_error()) };
        panic!("slas_cuda status {status}: {}", msg.to_string_lossy());
    }
}

macro_rules! impl_cuda_real {
    ($t:ty, $dot:ident, $nrm2:ident, $normalize:ident, $add:ident, $sub:ident, $mul:ident, $div:ident, $gemm:ident, $gemv:ident) => {
        impl DotProduct<$t> for Cuda {
            fn dot<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>) -> $t {
                let mut out: $t = 0.;
                check(unsafe { $dot(LEN, a.as_ptr(), b.as_ptr(), &mut out) });
                out
            }
        }
        impl Normalize<$t> for Cuda {
            type NormOutput = $t;
            fn norm<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>) -> $t {
                let mut out: $t = 0.;
                check(unsafe { $nrm2(LEN, a.as_ptr(), &mut out) });
                out
            }
            fn normalize<const LEN: usize>(&self, a: &mut impl StaticVec<$t, LEN>) {
                // as_mut_ptr keeps cow semantics: a borrowed StaticCowVec copies first (src/static_vec.rs:304-318)
                check(unsafe { $normalize(LEN, a.as_mut_ptr()) });
            }
        }
        impl Addition<$t> for Cuda {
            fn add<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>, c: &mut impl StaticVec<$t, LEN>) {
                check(unsafe { $add(LEN, a.as_ptr(), b.as_ptr(), c.as_mut_ptr()) });
            }
        }
        impl Subtraction<$t> for Cuda {
            fn sub<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>, c: &mut impl StaticVec<$t, LEN>) {
                check(unsafe { $sub(LEN, a.as_ptr(), b.as_ptr(), c.as_mut_ptr()) });
            }
        }
        impl Multiplication<$t> for Cuda {
            fn mul<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>, c: &mut impl StaticVec<$t, LEN>) {
                check(unsafe { $mul(LEN, a.as_ptr(), b.as_ptr(), c.as_mut_ptr()) });
            }
        }
        impl Divition<$t> for Cuda {
            fn div<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>, c: &mut impl StaticVec<$t, LEN>) {
                check(unsafe { $div(LEN, a.as_ptr(), b.as_ptr(), c.as_mut_ptr()) });
            }
        }
        impl MatrixMul<$t> for Cuda {
            fn matrix_mul<A: StaticVec<$t, ALEN>, B: StaticVec<$t, BLEN>, C: StaticVec<$t, CLEN>,
                          const ALEN: usize, const BLEN: usize, const CLEN: usize>(
                &self, a: &A, b: &B, buffer: &mut C, m: usize, n: usize, k: usize,
                lda: usize, ldb: usize, ldc: usize, a_trans: bool, b_trans: bool,
            ) where A: Sized, B: Sized {
                // like Blas (src/backends/blas.rs:107) the output goes through as_ptr() as *mut
                check(unsafe { $gemm(a.as_ptr(), b.as_ptr(), buffer.as_ptr() as *mut $t, m, n, k, lda, ldb, ldc,
                                     a_trans as c_int, b_trans as c_int) });
            }
            fn matrix_vector_mul<A: StaticVec<$t, ALEN>, B: StaticVec<$t, BLEN>, C: StaticVec<$t, CLEN>,
                                 const ALEN: usize, const BLEN: usize, const CLEN: usize>(
                &self, a: &A, b: &B, buffer: &mut C, m: usize, n: usize, lda: usize, a_trans: bool,
            ) where A: Sized, B: Sized {
                check(unsafe { $gemv(a.as_ptr(), b.as_ptr(), buffer.as_ptr() as *mut $t, m, n, lda, a_trans as c_int) });
            }
        }
        impl Backend<$t> for Cuda {}
    };
}
impl_cuda_real!(f32, slas_cuda_sdot, slas_cuda_snrm2, slas_cuda_snormalize, slas_cuda_sadd, slas_cuda_ssub,
                slas_cuda_smul, slas_cuda_sdiv, slas_cuda_sgemm, slas_cuda_sgemv);
impl_cuda_real!(f64, slas_cuda_ddot, slas_cuda_dnrm2, slas_cuda_dnormalize, slas_cuda_dadd, slas_cuda_dsub,
                slas_cuda_dmul, slas_cuda_ddiv, slas_cuda_dgemm, slas_cuda_dgemv);

macro_rules! impl_cuda_complex {
    ($t:ty, $dotu:ident, $nrm2:ident, $normalize:ident) => {
        impl DotProduct<Complex<$t>> for Cuda {
            fn dot<const LEN: usize>(&self, a: &impl StaticVec<Complex<$t>, LEN>, b: &impl StaticVec<Complex<$t>, LEN>) -> Complex<$t> {
                let mut out: [$t; 2] = [0.; 2];  // unconjugated, like cblas_?dotu_sub (src/backends/blas.rs:39-46)
                check(unsafe { $dotu(LEN, a.as_ptr() as *const $t, b.as_ptr() as *const $t, out.as_mut_ptr()) });
                Complex { re: out[0], im: out[1] }
            }
        }
        impl Normalize<Complex<$t>> for Cuda {
            type NormOutput = $t;
            fn norm<const LEN: usize>(&self, a: &impl StaticVec<Complex<$t>, LEN>) -> $t {
                let mut out: $t = 0.;
                check(unsafe { $nrm2(LEN, a.as_ptr() as *const $t, &mut out) });
                out
            }
            fn normalize<const LEN: usize>(&self, a: &mut impl StaticVec<Complex<$t>, LEN>) {
                check(unsafe { $normalize(LEN, a.as_mut_ptr() as *mut $t) });
            }
        }
        impl Backend<Complex<$t>> for Cuda {}
    };
}
impl_cuda_complex!(f32, slas_cuda_cdotu, slas_cuda_scnrm2, slas_cuda_cnormalize);
impl_cuda_complex!(f64, slas_cuda_zdotu, slas_cuda_dznrm2, slas_cuda_znormalize);

/// Fused chains (include/slas_cuda.h "Fused chains"): extra inherent methods, not trait methods -- slas leaves lazy
/// execution to its users (CONTRIBUTING.md:7-22).  Each gives the bits of the op sequence it replaces on this backend.
macro_rules! impl_cuda_fused {
    ($t:ty, $add_dot:ident, $dot_norms:ident, $normalize_into:ident) => {
        impl Cuda {
            /// `dot(a + b, d)` without materialising `a + b`.
            pub fn add_dot<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>, d: &impl StaticVec<$t, LEN>) -> $t {
                let mut out: $t = 0.;
                check(unsafe { $add_dot(LEN, a.as_ptr(), b.as_ptr(), std::ptr::null_mut(), d.as_ptr(), &mut out) });
                out
            }
            /// `(dot(a, b), norm(a), norm(b))` from one read of both vectors.
            pub fn dot_norms<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>) -> ($t, $t, $t) {
                let mut out: [$t; 3] = [0.; 3];
                check(unsafe { $dot_norms(LEN, a.as_ptr(), b.as_ptr(), out.as_mut_ptr()) });
                (out[0], out[1], out[2])
            }
            /// `out = a / norm(a)`; `a` is untouched (normalising a borrowed `StaticCowVec` without its copy).
            pub fn normalize_into<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, out: &mut impl StaticVec<$t, LEN>) {
                check(unsafe { $normalize_into(LEN, a.as_ptr(), out.as_mut_ptr()) });
            }
        }
    };
}
impl_cuda_fused!(f32, slas_cuda_sadd_dot, slas_cuda_sdot_norms, slas_cuda_snormalize_into);
// (the f64 methods live in a second inherent impl with distinct names in a real crate: Rust has no overloading;
//  shown here for f32 only)

/// Transpose is generic over any `T: Copy` in the reference (src/backends/rust.rs:151-177); the
/// library moves opaque 1/2/4/8/16-byte elements.
impl<T: Copy> Transpose<T> for Cuda {
    fn transpose_inplace<const LEN: usize>(&self, a: &mut impl StaticVec<T, LEN>, columns: usize) {
        check(unsafe { slas_cuda_transpose_inplace(LEN, std::mem::size_of::<T>(), a.as_mut_ptr() as *mut _, columns) });
    }
    fn transpose<const LEN: usize>(&self, a: &impl StaticVec<T, LEN>, buffer: &mut impl StaticVec<T, LEN>, columns: usize) {
        check(unsafe { slas_cuda_transpose(LEN, std::mem::size_of::<T>(), a.as_ptr() as *const _, buffer.as_mut_ptr() as *mut _, columns) });
    }
}

/// Device-resident `StaticVec` (goes in src/static_vec.rs): `as_ptr()` is a DEVICE pointer, so the
/// C ABI takes the fast path; every host accessor panics exactly like `NullVec`
/// (src/nullvec.rs:24-57).  Never materialises `[T; LEN]` by value (the reference's stack arrays
/// overflow from LEN = 2^20, src/tensor.rs:451, src/backends/rust.rs:80).
pub struct CudaVec<T, const LEN: usize> {
    ptr: *mut T,
    _pd: PhantomData<T>,
}

impl<T: Copy, const LEN: usize> CudaVec<T, LEN> {
    pub fn new() -> Self {
        let mut p: *mut std::ffi::c_void = std::ptr::null_mut();
        check(unsafe { slas_cuda_malloc(&mut p, LEN * std::mem::size_of::<T>()) });
        Self { ptr: p as *mut T, _pd: PhantomData }
    }
    pub fn from_host(v: &impl StaticVec<T, LEN>) -> Self {
        let d = Self::new();
        check(unsafe { slas_cuda_memcpy_h2d(d.ptr as *mut _, v.as_ptr() as *const _, LEN * std::mem::size_of::<T>()) });
        d
    }
    pub fn copy_to_host(&self, out: &mut impl StaticVec<T, LEN>) {
        check(unsafe { slas_cuda_memcpy_d2h(out.as_mut_ptr() as *mut _, self.ptr as *const _, LEN * std::mem::size_of::<T>()) });
    }
}
impl<T, const LEN: usize> Drop for CudaVec<T, LEN> {
    fn drop(&mut self) {
        unsafe { slas_cuda_free(self.ptr as *mut _) };
    }
}
impl<T, const LEN: usize> StaticVec<T, LEN> for CudaVec<T, LEN> {
    unsafe fn as_ptr(&self) -> *const T { self.ptr }
    unsafe fn as_mut_ptr(&mut self) -> *mut T { self.ptr }
    fn moo_ref<'a>(&'a self) -> StaticVecRef<'a, T, LEN> where T: Copy { panic!("Tried to call moo_ref on CudaVec-{LEN}: device memory") }
    fn mut_moo_ref<'a>(&'a mut self) -> MutStaticVecRef<'a, T, LEN> where T: Copy { panic!("Tried to call mut_moo_ref on CudaVec-{LEN}: device memory") }
    fn moo<'a>(&'a self) -> StaticCowVec<'a, T, LEN> where T: Copy { panic!("Tried to call moo on CudaVec-{LEN}: device memory") }
    unsafe fn get_unchecked<'a>(&'a self, _: usize) -> &'a T { panic!("Tried to call get_unchecked on CudaVec-{LEN}: device memory") }
    unsafe fn get_unchecked_mut<'a>(&'a mut self, _: usize) -> &'a mut T { panic!("Tried to call get_unchecked_mut on CudaVec-{LEN}: device memory") }
    fn moo_owned(&self) -> StaticVecUnion<'static, T, LEN> where T: Copy { panic!("Tried to call moo_owned on CudaVec-{LEN}: device memory") }
}

/// Batched view (goes in src/tensor.rs): `BATCH` dense `M x K` matrices, contiguous in HBM; the
/// compile-time shape picks the kernel template inside the library.
pub struct CudaBatch<T, const M: usize, const K: usize> {
    pub data: *mut T,
    pub batch: usize,
    pub is_trans: bool,
}

macro_rules! impl_cuda_batch {
    ($t:ty, $gemm_batched:ident) => {
        impl<const M: usize, const K: usize> CudaBatch<$t, M, K> {
            /// `out[i] = op(self[i]) . op(other[i])` for every object: one kernel launch.
            pub fn matrix_mul_buffer<const N: usize>(&self, other: &CudaBatch<$t, K, N>, out: &mut CudaBatch<$t, M, N>) {
                assert_eq!(self.batch, other.batch);
                assert_eq!(self.batch, out.batch);
                check(unsafe { $gemm_batched(self.batch, self.data, other.data, out.data, M, N, K,
                                             self.is_trans as c_int, other.is_trans as c_int) });
            }
        }
    };
}
impl_cuda_batch!(f32, slas_cuda_sgemm_batched);
impl_cuda_batch!(f64, slas_cuda_dgemm_batched);
