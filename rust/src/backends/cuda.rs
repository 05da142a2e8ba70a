//! `slas_backend::Cuda` -- the Rust side of the B200 backend.
//!
//! NON-BUILDING SOURCE: this image has no rustc/cargo (and slas needs a ~2022 nightly), so this file has never been
//! compiled; it is gated behind the `cuda` cargo feature and is the binding a slas maintainer adds next to
//! `src/backends/rust.rs` and `src/backends/blas.rs`.  The arithmetic lives in libslas_cuda.so
//! (include/slas_cuda.h), which IS built and tested here through the same C ABI; `tests/test_boundary_drift.py`
//! keeps the `extern "C"` block below identical, symbol by symbol and type by type, to that header
//! (`python tests/test_boundary_drift.py --emit` regenerates it).
//!
//! Registration (src/backends.rs, beside lines 333-337):
//! ```ignore
//! #[cfg(feature = "cuda")] mod cuda;
//! #[cfg(feature = "cuda")] pub use cuda::{Cuda, CudaVec, CudaBatch, CudaFused, CudaSharded};
//! ```
//! Run-time knobs (environment, read by the library): `SLAS_CUDA_DEVICES=n|all` -- drive n GPUs from this one
//! process (host operands are split over them, `*_multi` reductions exchange over NVLink peer memory);
//! `SLAS_CUDA_HOST_NUMA=0` -- do not place pinned memory on the GPU's NUMA node.
use super::*;
use operations::*;
use std::ffi::c_void;
use std::marker::PhantomData;
use std::os::raw::{c_char, c_int, c_long};

/// A zero-sized backend like `Rust` and `Blas` (src/backends/rust.rs:2-3): `B::default()` is
/// called at every call site, so device, stream and scratch live in the C library.
#[derive(Default, Clone, Copy)]
pub struct Cuda;

#[link(name = "slas_cuda")]
extern "C" {
    fn slas_cuda_init(device: c_int) -> c_int;
    fn slas_cuda_shutdown() -> c_int;
    fn slas_cuda_last_error() -> *const c_char;
    fn slas_cuda_device_count(count: *mut c_int) -> c_int;
    fn slas_cuda_device_info(sm_count: *mut c_int, l2_bytes: *mut usize, hbm_bytes: *mut usize, cc_major: *mut c_int, cc_minor: *mut c_int) -> c_int;
    fn slas_cuda_set_stream(cuda_stream: *mut c_void) -> c_int;
    fn slas_cuda_synchronize() -> c_int;
    fn slas_cuda_launch_count() -> u64;
    fn slas_cuda_graph_begin() -> c_int;
    fn slas_cuda_graph_end(graph_exec: *mut *mut c_void) -> c_int;
    fn slas_cuda_graph_launch(graph_exec: *mut c_void) -> c_int;
    fn slas_cuda_graph_destroy(graph_exec: *mut c_void) -> c_int;
    fn slas_cuda_set_tunable(name: *const c_char, value: c_long) -> c_int;
    fn slas_cuda_get_tunable(name: *const c_char, value: *mut c_long) -> c_int;
    fn slas_cuda_bench_fma_peak(is_double: c_int, tflops: *mut f64) -> c_int;
    fn slas_cuda_bench_tf32_peak(tflops: *mut f64) -> c_int;
    fn slas_cuda_bench_call_loop(op: c_int, len: usize, a: *const f32, b: *const f32, out: *mut f32, calls: c_int, stride: usize) -> c_int;
    fn slas_cuda_bench_divby_mismatches(divisors: *const f32, count: usize, mismatches: *mut u64) -> c_int;
    fn slas_cuda_get_reduce_phases(ns5: *mut u64) -> c_int;
    fn slas_cuda_malloc(dptr: *mut *mut c_void, bytes: usize) -> c_int;
    fn slas_cuda_free(dptr: *mut c_void) -> c_int;
    fn slas_cuda_malloc_host(hptr: *mut *mut c_void, bytes: usize) -> c_int;
    fn slas_cuda_free_host(hptr: *mut c_void) -> c_int;
    fn slas_cuda_host_register(hptr: *mut c_void, bytes: usize) -> c_int;
    fn slas_cuda_host_unregister(hptr: *mut c_void) -> c_int;
    fn slas_cuda_memcpy_h2d(dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    fn slas_cuda_memcpy_d2h(dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    fn slas_cuda_memcpy_d2d(dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    fn slas_cuda_memset(dst: *mut c_void, byte: c_int, bytes: usize) -> c_int;
    fn slas_cuda_is_device_ptr(p: *const c_void, is_device: *mut c_int) -> c_int;
    fn slas_cuda_sfill_uniform(len: usize, x: *mut f32, seed: u64, offset: u64, lo: f32, hi: f32) -> c_int;
    fn slas_cuda_dfill_uniform(len: usize, x: *mut f64, seed: u64, offset: u64, lo: f64, hi: f64) -> c_int;
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
    fn slas_cuda_sadd_dot(len: usize, a: *const f32, b: *const f32, c: *mut f32, d: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_ssub_dot(len: usize, a: *const f32, b: *const f32, c: *mut f32, d: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_smul_dot(len: usize, a: *const f32, b: *const f32, c: *mut f32, d: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_sdiv_dot(len: usize, a: *const f32, b: *const f32, c: *mut f32, d: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_dadd_dot(len: usize, a: *const f64, b: *const f64, c: *mut f64, d: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_dsub_dot(len: usize, a: *const f64, b: *const f64, c: *mut f64, d: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_dmul_dot(len: usize, a: *const f64, b: *const f64, c: *mut f64, d: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_ddiv_dot(len: usize, a: *const f64, b: *const f64, c: *mut f64, d: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_sdot_norms(len: usize, a: *const f32, b: *const f32, out3: *mut f32) -> c_int;
    fn slas_cuda_ddot_norms(len: usize, a: *const f64, b: *const f64, out3: *mut f64) -> c_int;
    fn slas_cuda_snormalize_into(len: usize, a: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_dnormalize_into(len: usize, a: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_transpose(len: usize, elem_size: usize, a: *const c_void, buffer: *mut c_void, columns: usize) -> c_int;
    fn slas_cuda_transpose_inplace(len: usize, elem_size: usize, a: *mut c_void, columns: usize) -> c_int;
    fn slas_cuda_sgemm(a: *const f32, b: *const f32, c: *mut f32, m: usize, n: usize, k: usize, lda: usize, ldb: usize, ldc: usize, a_trans: c_int, b_trans: c_int) -> c_int;
    fn slas_cuda_dgemm(a: *const f64, b: *const f64, c: *mut f64, m: usize, n: usize, k: usize, lda: usize, ldb: usize, ldc: usize, a_trans: c_int, b_trans: c_int) -> c_int;
    fn slas_cuda_set_sgemm_path(path: c_int) -> c_int;
    fn slas_cuda_get_sgemm_path(path: *mut c_int) -> c_int;
    fn slas_cuda_sgemv(a: *const f32, x: *const f32, y: *mut f32, m: usize, n: usize, lda: usize, a_trans: c_int) -> c_int;
    fn slas_cuda_dgemv(a: *const f64, x: *const f64, y: *mut f64, m: usize, n: usize, lda: usize, a_trans: c_int) -> c_int;
    fn slas_cuda_sgemm_batched(batch: usize, a: *const f32, b: *const f32, c: *mut f32, m: usize, n: usize, k: usize, a_trans: c_int, b_trans: c_int) -> c_int;
    fn slas_cuda_dgemm_batched(batch: usize, a: *const f64, b: *const f64, c: *mut f64, m: usize, n: usize, k: usize, a_trans: c_int, b_trans: c_int) -> c_int;
    fn slas_cuda_sgemv_batched(batch: usize, a: *const f32, x: *const f32, y: *mut f32, m: usize, n: usize, a_trans: c_int) -> c_int;
    fn slas_cuda_dgemv_batched(batch: usize, a: *const f64, x: *const f64, y: *mut f64, m: usize, n: usize, a_trans: c_int) -> c_int;
    fn slas_cuda_transpose_batched(batch: usize, len: usize, elem_size: usize, a: *const c_void, buffer: *mut c_void, columns: usize) -> c_int;
    fn slas_cuda_transpose_inplace_batched(batch: usize, len: usize, elem_size: usize, a: *mut c_void, columns: usize) -> c_int;
    fn slas_cuda_sdot_batched(batch: usize, len: usize, a: *const f32, b: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_ddot_batched(batch: usize, len: usize, a: *const f64, b: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_snrm2_batched(batch: usize, len: usize, a: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_dnrm2_batched(batch: usize, len: usize, a: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_snormalize_batched(batch: usize, len: usize, a: *mut f32) -> c_int;
    fn slas_cuda_dnormalize_batched(batch: usize, len: usize, a: *mut f64) -> c_int;
    fn slas_cuda_init_devices(ndev: c_int) -> c_int;
    fn slas_cuda_devices(ndev: *mut c_int, peer_exchange: *mut c_int) -> c_int;
    fn slas_cuda_malloc_on(device: c_int, dptr: *mut *mut c_void, bytes: usize) -> c_int;
    fn slas_cuda_sdot_multi(nslices: c_int, lens: *const usize, a: *const *const f32, b: *const *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_ddot_multi(nslices: c_int, lens: *const usize, a: *const *const f64, b: *const *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_snrm2_multi(nslices: c_int, lens: *const usize, a: *const *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_dnrm2_multi(nslices: c_int, lens: *const usize, a: *const *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_snormalize_multi(nslices: c_int, lens: *const usize, a: *const *mut f32) -> c_int;
    fn slas_cuda_dnormalize_multi(nslices: c_int, lens: *const usize, a: *const *mut f64) -> c_int;
    fn slas_cuda_nccl_unique_id(id128: *mut c_void) -> c_int;
    fn slas_cuda_comm_init(nranks: c_int, rank: c_int, id128: *const c_void) -> c_int;
    fn slas_cuda_comm_destroy() -> c_int;
    fn slas_cuda_comm_info(nranks: *mut c_int, rank: *mut c_int) -> c_int;
    fn slas_cuda_p2p_handle(handle64: *mut c_void) -> c_int;
    fn slas_cuda_p2p_init(nranks: c_int, rank: c_int, handles: *const c_void) -> c_int;
    fn slas_cuda_p2p_destroy() -> c_int;
    fn slas_cuda_shard_range(len: usize, elem_size: usize, nranks: c_int, rank: c_int, begin: *mut usize, end: *mut usize) -> c_int;
    fn slas_cuda_sdot_sharded(local_len: usize, a: *const f32, b: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_ddot_sharded(local_len: usize, a: *const f64, b: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_snrm2_sharded(local_len: usize, a: *const f32, out: *mut f32) -> c_int;
    fn slas_cuda_dnrm2_sharded(local_len: usize, a: *const f64, out: *mut f64) -> c_int;
    fn slas_cuda_snormalize_sharded(local_len: usize, a: *mut f32) -> c_int;
    fn slas_cuda_dnormalize_sharded(local_len: usize, a: *mut f64) -> c_int;
    fn slas_cuda_sdot_partial(len: usize, a: *const f32, b: *const f32, partial: *mut f64) -> c_int;
    fn slas_cuda_ddot_partial(len: usize, a: *const f64, b: *const f64, partial: *mut f64) -> c_int;
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

/// Fused chains (include/slas_cuda.h "Fused chains"): an extra trait, not new methods on the operation traits -- slas
/// leaves lazy execution to its users (CONTRIBUTING.md:7-22).  Each gives the bits of the op sequence it replaces on
/// this backend.  `c: None` never materialises the element-wise intermediate.
pub trait CudaFused<T> {
    fn add_dot<const LEN: usize>(&self, a: &impl StaticVec<T, LEN>, b: &impl StaticVec<T, LEN>, c: Option<&mut dyn StaticVecMut<T, LEN>>, d: &impl StaticVec<T, LEN>) -> T;
    fn sub_dot<const LEN: usize>(&self, a: &impl StaticVec<T, LEN>, b: &impl StaticVec<T, LEN>, c: Option<&mut dyn StaticVecMut<T, LEN>>, d: &impl StaticVec<T, LEN>) -> T;
    fn mul_dot<const LEN: usize>(&self, a: &impl StaticVec<T, LEN>, b: &impl StaticVec<T, LEN>, c: Option<&mut dyn StaticVecMut<T, LEN>>, d: &impl StaticVec<T, LEN>) -> T;
    fn div_dot<const LEN: usize>(&self, a: &impl StaticVec<T, LEN>, b: &impl StaticVec<T, LEN>, c: Option<&mut dyn StaticVecMut<T, LEN>>, d: &impl StaticVec<T, LEN>) -> T;
    /// `(dot(a, b), norm(a), norm(b))` from one read of both vectors.
    fn dot_norms<const LEN: usize>(&self, a: &impl StaticVec<T, LEN>, b: &impl StaticVec<T, LEN>) -> (T, T, T);
    /// `out = a / norm(a)`; `a` is untouched (normalising a borrowed `StaticCowVec` without its copy).
    fn normalize_into<const LEN: usize>(&self, a: &impl StaticVec<T, LEN>, out: &mut impl StaticVec<T, LEN>);
}
/// Object-safe face of "something with a mutable pointer to LEN elements" for the optional intermediate above.
pub trait StaticVecMut<T, const LEN: usize> { unsafe fn mut_ptr(&mut self) -> *mut T; }
impl<T, U: StaticVec<T, LEN>, const LEN: usize> StaticVecMut<T, LEN> for U { unsafe fn mut_ptr(&mut self) -> *mut T { self.as_mut_ptr() } }

macro_rules! impl_cuda_fused {
    ($t:ty, $add_dot:ident, $sub_dot:ident, $mul_dot:ident, $div_dot:ident, $dot_norms:ident, $normalize_into:ident) => {
        impl CudaFused<$t> for Cuda {
            fn add_dot<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>, c: Option<&mut dyn StaticVecMut<$t, LEN>>, d: &impl StaticVec<$t, LEN>) -> $t {
                let mut out: $t = 0.;
                let cp = c.map_or(std::ptr::null_mut(), |c| unsafe { c.mut_ptr() });
                check(unsafe { $add_dot(LEN, a.as_ptr(), b.as_ptr(), cp, d.as_ptr(), &mut out) });
                out
            }
            fn sub_dot<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>, c: Option<&mut dyn StaticVecMut<$t, LEN>>, d: &impl StaticVec<$t, LEN>) -> $t {
                let mut out: $t = 0.;
                let cp = c.map_or(std::ptr::null_mut(), |c| unsafe { c.mut_ptr() });
                check(unsafe { $sub_dot(LEN, a.as_ptr(), b.as_ptr(), cp, d.as_ptr(), &mut out) });
                out
            }
            fn mul_dot<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>, c: Option<&mut dyn StaticVecMut<$t, LEN>>, d: &impl StaticVec<$t, LEN>) -> $t {
                let mut out: $t = 0.;
                let cp = c.map_or(std::ptr::null_mut(), |c| unsafe { c.mut_ptr() });
                check(unsafe { $mul_dot(LEN, a.as_ptr(), b.as_ptr(), cp, d.as_ptr(), &mut out) });
                out
            }
            fn div_dot<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>, c: Option<&mut dyn StaticVecMut<$t, LEN>>, d: &impl StaticVec<$t, LEN>) -> $t {
                let mut out: $t = 0.;
                let cp = c.map_or(std::ptr::null_mut(), |c| unsafe { c.mut_ptr() });
                check(unsafe { $div_dot(LEN, a.as_ptr(), b.as_ptr(), cp, d.as_ptr(), &mut out) });
                out
            }
            fn dot_norms<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>) -> ($t, $t, $t) {
                let mut out: [$t; 3] = [0.; 3];
                check(unsafe { $dot_norms(LEN, a.as_ptr(), b.as_ptr(), out.as_mut_ptr()) });
                (out[0], out[1], out[2])
            }
            fn normalize_into<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, out: &mut impl StaticVec<$t, LEN>) {
                check(unsafe { $normalize_into(LEN, a.as_ptr(), out.as_mut_ptr()) });
            }
        }
    };
}
impl_cuda_fused!(f32, slas_cuda_sadd_dot, slas_cuda_ssub_dot, slas_cuda_smul_dot, slas_cuda_sdiv_dot, slas_cuda_sdot_norms, slas_cuda_snormalize_into);
impl_cuda_fused!(f64, slas_cuda_dadd_dot, slas_cuda_dsub_dot, slas_cuda_dmul_dot, slas_cuda_ddiv_dot, slas_cuda_ddot_norms, slas_cuda_dnormalize_into);

/// Transpose is generic over any `T: Copy` in the reference (src/backends/rust.rs:151-177); the
/// library moves opaque elements of any size (1/2/4/8/16-byte ones on the vector kernels).
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
        let mut p: *mut c_void = std::ptr::null_mut();
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

/// Batched view (goes in src/tensor.rs): `batch` objects of `LEN` elements each, `[[T; LEN]; batch]` contiguous,
/// in host memory (split over the initialised devices and pipelined by the library) or in HBM (`device: true`, one
/// launch).  The compile-time object shape picks the kernel template inside the library.
pub struct CudaBatch<T, const LEN: usize> {
    pub data: *mut T,
    pub batch: usize,
    pub is_trans: bool,
    owned_on_device: bool,
}

impl<T: Copy, const LEN: usize> CudaBatch<T, LEN> {
    /// `batch` objects in HBM.
    pub fn new(batch: usize) -> Self {
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { slas_cuda_malloc(&mut p, batch * LEN * std::mem::size_of::<T>()) });
        Self { data: p as *mut T, batch, is_trans: false, owned_on_device: true }
    }
    /// A view of `batch` contiguous host (or device) objects the caller owns.
    pub unsafe fn from_ptr(data: *mut T, batch: usize) -> Self { Self { data, batch, is_trans: false, owned_on_device: false } }
    pub fn upload(host: &[T]) -> Self {
        assert_eq!(host.len() % LEN, 0);
        let b = Self::new(host.len() / LEN);
        check(unsafe { slas_cuda_memcpy_h2d(b.data as *mut _, host.as_ptr() as *const _, host.len() * std::mem::size_of::<T>()) });
        b
    }
    pub fn download(&self, out: &mut [T]) {
        assert_eq!(out.len(), self.batch * LEN);
        check(unsafe { slas_cuda_memcpy_d2h(out.as_mut_ptr() as *mut _, self.data as *const _, out.len() * std::mem::size_of::<T>()) });
    }
    /// Lazy transpose of every object (src/tensor.rs:512-522): a flag, consumed as a_trans / b_trans.
    pub fn transpose(mut self) -> Self { self.is_trans = !self.is_trans; self }
    /// `Backend::transpose` per object (src/backends/rust.rs:161-177): `columns` = column count of the OUTPUT.
    pub fn transpose_into(&self, out: &mut CudaBatch<T, LEN>, columns: usize) {
        assert_eq!(self.batch, out.batch);
        check(unsafe { slas_cuda_transpose_batched(self.batch, LEN, std::mem::size_of::<T>(), self.data as *const _, out.data as *mut _, columns) });
    }
    /// `Matrix::deref_mut` per object (src/tensor.rs:546-564): materialise a lazy transpose in place.
    pub fn transpose_inplace(&mut self, columns: usize) {
        check(unsafe { slas_cuda_transpose_inplace_batched(self.batch, LEN, std::mem::size_of::<T>(), self.data as *mut _, columns) });
    }
    /// Batch-index shard `rank` of `nranks` (multi-GPU: no collective on this path).
    pub fn shard(&self, nranks: usize, rank: usize) -> CudaBatch<T, LEN> {
        let (lo, hi) = (self.batch * rank / nranks, self.batch * (rank + 1) / nranks);
        CudaBatch { data: unsafe { self.data.add(lo * LEN) }, batch: hi - lo, is_trans: self.is_trans, owned_on_device: false }
    }
}
impl<T, const LEN: usize> Drop for CudaBatch<T, LEN> {
    fn drop(&mut self) { if self.owned_on_device { unsafe { slas_cuda_free(self.data as *mut _) }; } }
}

macro_rules! impl_cuda_batch {
    ($t:ty, $gemm_batched:ident, $gemv_batched:ident, $dot_batched:ident, $nrm2_batched:ident, $normalize_batched:ident) => {
        impl<const ALEN: usize> CudaBatch<$t, ALEN> {
            /// `out[i] = op(self[i]) . op(other[i])` with `self[i]` an M x K matrix (ALEN = M * K), `other[i]` K x N,
            /// `out[i]` M x N: the per-object `Matrix::matrix_mul_buffer` (src/tensor.rs:364-407), one launch.
            pub fn matrix_mul_buffer<const BLEN: usize, const OLEN: usize>(&self, other: &CudaBatch<$t, BLEN>, out: &mut CudaBatch<$t, OLEN>, m: usize, n: usize, k: usize) {
                assert_eq!(m * k, ALEN); assert_eq!(k * n, BLEN); assert_eq!(m * n, OLEN);
                assert_eq!(self.batch, other.batch); assert_eq!(self.batch, out.batch);
                check(unsafe { $gemm_batched(self.batch, self.data, other.data, out.data, m, n, k, self.is_trans as c_int, other.is_trans as c_int) });
            }
            /// `out[i] = op(self[i]) . x[i]`, the trait's argument meaning: `m` = underlying width, `n` = underlying
            /// height (src/tensor.rs:410-438, src/backends/blas.rs:114-151).
            pub fn vector_mul_buffer<const XLEN: usize, const OLEN: usize>(&self, x: &CudaBatch<$t, XLEN>, out: &mut CudaBatch<$t, OLEN>, m: usize, n: usize) {
                assert_eq!(m * n, ALEN);
                assert_eq!(XLEN, if self.is_trans { n } else { m }); assert_eq!(OLEN, if self.is_trans { m } else { n });
                assert_eq!(self.batch, x.batch); assert_eq!(self.batch, out.batch);
                check(unsafe { $gemv_batched(self.batch, self.data, x.data, out.data, m, n, self.is_trans as c_int) });
            }
            /// `out[i] = dot(self[i], other[i])`
            pub fn dot(&self, other: &CudaBatch<$t, ALEN>, out: &mut [$t]) {
                assert_eq!(self.batch, other.batch); assert_eq!(self.batch, out.len());
                check(unsafe { $dot_batched(self.batch, ALEN, self.data, other.data, out.as_mut_ptr()) });
            }
            /// `out[i] = norm(self[i])`
            pub fn norm(&self, out: &mut [$t]) {
                assert_eq!(self.batch, out.len());
                check(unsafe { $nrm2_batched(self.batch, ALEN, self.data, out.as_mut_ptr()) });
            }
            /// `self[i] /= norm(self[i])` (true division, like src/backends/rust.rs:122-127 per object)
            pub fn normalize(&mut self) {
                check(unsafe { $normalize_batched(self.batch, ALEN, self.data) });
            }
        }
    };
}
impl_cuda_batch!(f32, slas_cuda_sgemm_batched, slas_cuda_sgemv_batched, slas_cuda_sdot_batched, slas_cuda_snrm2_batched, slas_cuda_snormalize_batched);
impl_cuda_batch!(f64, slas_cuda_dgemm_batched, slas_cuda_dgemv_batched, slas_cuda_ddot_batched, slas_cuda_dnrm2_batched, slas_cuda_dnormalize_batched);

/// One huge vector over several GPUs.  Two forms:
///  * one process per GPU (`*_sharded`): this rank's slice + one exchange of the f64 partial -- inside the reduction
///    kernel over NVLink peer memory once `p2p_init` has mapped the peers' mailboxes, else one `ncclAllReduce`;
///  * one process, several GPUs (`*_multi`, after `init_devices`): one device-resident slice per device.
pub trait CudaSharded<T> {
    fn dot_sharded<const LEN: usize>(&self, a_local: &impl StaticVec<T, LEN>, b_local: &impl StaticVec<T, LEN>) -> T;
    fn norm_sharded<const LEN: usize>(&self, a_local: &impl StaticVec<T, LEN>) -> T;
    fn normalize_sharded<const LEN: usize>(&self, a_local: &mut impl StaticVec<T, LEN>);
    /// `slices[i]` = (device pointer, length) of the part of the vector that lives on GPU i
    fn dot_multi(&self, a: &[(*const T, usize)], b: &[(*const T, usize)]) -> T;
    fn norm_multi(&self, a: &[(*const T, usize)]) -> T;
    fn normalize_multi(&self, a: &[(*mut T, usize)]);
}
macro_rules! impl_cuda_sharded {
    ($t:ty, $dot:ident, $nrm2:ident, $normalize:ident, $dot_m:ident, $nrm2_m:ident, $normalize_m:ident) => {
        impl CudaSharded<$t> for Cuda {
            fn dot_sharded<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>, b: &impl StaticVec<$t, LEN>) -> $t {
                let mut out: $t = 0.;
                check(unsafe { $dot(LEN, a.as_ptr(), b.as_ptr(), &mut out) });
                out
            }
            fn norm_sharded<const LEN: usize>(&self, a: &impl StaticVec<$t, LEN>) -> $t {
                let mut out: $t = 0.;
                check(unsafe { $nrm2(LEN, a.as_ptr(), &mut out) });
                out
            }
            fn normalize_sharded<const LEN: usize>(&self, a: &mut impl StaticVec<$t, LEN>) {
                check(unsafe { $normalize(LEN, a.as_mut_ptr()) });
            }
            fn dot_multi(&self, a: &[(*const $t, usize)], b: &[(*const $t, usize)]) -> $t {
                assert_eq!(a.len(), b.len());
                let lens: Vec<usize> = a.iter().map(|s| s.1).collect();
                assert!(b.iter().zip(&lens).all(|(s, l)| s.1 == *l));
                let (pa, pb): (Vec<*const $t>, Vec<*const $t>) = (a.iter().map(|s| s.0).collect(), b.iter().map(|s| s.0).collect());
                let mut out: $t = 0.;
                check(unsafe { $dot_m(a.len() as c_int, lens.as_ptr(), pa.as_ptr(), pb.as_ptr(), &mut out) });
                out
            }
            fn norm_multi(&self, a: &[(*const $t, usize)]) -> $t {
                let lens: Vec<usize> = a.iter().map(|s| s.1).collect();
                let pa: Vec<*const $t> = a.iter().map(|s| s.0).collect();
                let mut out: $t = 0.;
                check(unsafe { $nrm2_m(a.len() as c_int, lens.as_ptr(), pa.as_ptr(), &mut out) });
                out
            }
            fn normalize_multi(&self, a: &[(*mut $t, usize)]) {
                let lens: Vec<usize> = a.iter().map(|s| s.1).collect();
                let pa: Vec<*mut $t> = a.iter().map(|s| s.0).collect();
                check(unsafe { $normalize_m(a.len() as c_int, lens.as_ptr(), pa.as_ptr()) });
            }
        }
    };
}
impl_cuda_sharded!(f32, slas_cuda_sdot_sharded, slas_cuda_snrm2_sharded, slas_cuda_snormalize_sharded, slas_cuda_sdot_multi, slas_cuda_snrm2_multi, slas_cuda_snormalize_multi);
impl_cuda_sharded!(f64, slas_cuda_ddot_sharded, slas_cuda_dnrm2_sharded, slas_cuda_dnormalize_sharded, slas_cuda_ddot_multi, slas_cuda_dnrm2_multi, slas_cuda_dnormalize_multi);

impl Cuda {
    /// Page-lock a long-lived host allocation in place (a `Vec`, a `Box<[T; LEN]>`): its operands are then copied by the
    /// DMA engines directly instead of through the library's bounce ring.  Call `host_unregister` before freeing it.
    pub unsafe fn host_register<T>(data: &mut [T]) { check(slas_cuda_host_register(data.as_mut_ptr() as *mut c_void, std::mem::size_of_val(data))); }
    pub unsafe fn host_unregister<T>(data: &mut [T]) { check(slas_cuda_host_unregister(data.as_mut_ptr() as *mut c_void)); }
    /// Drive `n` GPUs from this one process (0 = every visible device); same as `SLAS_CUDA_DEVICES=n` in the environment.
    pub fn init_devices(n: usize) -> usize {
        check(unsafe { slas_cuda_init_devices(n as c_int) });
        let (mut nd, mut px): (c_int, c_int) = (0, 0);
        check(unsafe { slas_cuda_devices(&mut nd, &mut px) });
        nd as usize
    }
    /// [begin, end) of the slice of a `len`-element vector of `T` owned by `rank`: 16-byte aligned interior boundaries.
    pub fn shard_range<T>(len: usize, nranks: usize, rank: usize) -> (usize, usize) {
        let (mut lo, mut hi) = (0usize, 0usize);
        check(unsafe { slas_cuda_shard_range(len, std::mem::size_of::<T>(), nranks as c_int, rank as c_int, &mut lo, &mut hi) });
        (lo, hi)
    }
    /// Which kernel a large f32 `matrix_mul` runs on: 0 = exact fp32 FMA chains (what `Blas`/`Rust` compute, up to
    /// summation order), 1 = tcgen05 3xTF32 wherever its tiling applies, 2 = auto (3xTF32 from 2^27 multiply-adds;
    /// an fp32-class result, <= 2e-6 * sum|a||b| measured, but not the ascending-k fp32 sum; lo parts of operands
    /// below ~2^-115 are flushed by the TF32 MMA).  The default is 2; a crate that must reproduce the CPU backends'
    /// rounding bit-for-bit-in-class calls `Cuda::set_sgemm_path(0)` once.
    pub fn set_sgemm_path(path: i32) { check(unsafe { slas_cuda_set_sgemm_path(path as c_int) }); }
    /// Record the device-pointer ops issued inside `f` into a CUDA graph; `launch` replays them with one launch.
    pub fn record_graph(f: impl FnOnce()) -> CudaGraph {
        check(unsafe { slas_cuda_graph_begin() });
        f();
        let mut g: *mut c_void = std::ptr::null_mut();
        check(unsafe { slas_cuda_graph_end(&mut g) });
        CudaGraph(g)
    }
    pub fn synchronize() { check(unsafe { slas_cuda_synchronize() }); }
}
pub struct CudaGraph(*mut c_void);
impl CudaGraph { pub fn launch(&self) { check(unsafe { slas_cuda_graph_launch(self.0) }); } }
impl Drop for CudaGraph { fn drop(&mut self) { unsafe { slas_cuda_graph_destroy(self.0) }; } }
