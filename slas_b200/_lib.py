"""ctypes binding of libslas_cuda.so -- the C ABI declared in include/slas_cuda.h.

This is the only way Python reaches the kernels: no torch ops, no CPU fallback.  If the
library has not been built (`python -c "import __graft_entry__ as g; g.build()"` or
`make -C slas_b200/csrc`) importing the ops fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libslas_cuda.so")
CSRC = os.path.join(_HERE, "csrc")

SLAS_OK, SLAS_ERR_INVALID, SLAS_ERR_CUDA, SLAS_ERR_NCCL, SLAS_ERR_UNSUPPORTED = range(5)


class SlasCudaError(RuntimeError):
    """Raised where the Rust shim panics: any non-zero status from the C ABI."""

    def __init__(self, status: int, message: str):
        super().__init__(f"slas_cuda status {status}: {message}")
        self.status = status


def build(force: bool = False) -> str:
    """Compile every .cu under slas_b200/csrc for sm_100a into slas_b200/lib/libslas_cuda.so."""
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", CSRC, "-j", str(min(os.cpu_count() or 4, 16))], check=True,
                   stdout=subprocess.DEVNULL)
    return LIB_PATH


_sz, _i, _vp, _u64 = C.c_size_t, C.c_int, C.c_void_p, C.c_uint64
_f, _d = C.c_float, C.c_double
_ip = C.POINTER(C.c_int)
_szp = C.POINTER(C.c_size_t)

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "slas_cuda_init": [_i],
    "slas_cuda_shutdown": [],
    "slas_cuda_last_error": [],
    "slas_cuda_device_count": [_ip],
    "slas_cuda_device_info": [_ip, _szp, _szp, _ip, _ip],
    "slas_cuda_set_stream": [_vp],
    "slas_cuda_synchronize": [],
    "slas_cuda_launch_count": [],
    "slas_cuda_graph_begin": [],
    "slas_cuda_graph_end": [C.POINTER(C.c_void_p)],
    "slas_cuda_graph_launch": [_vp],
    "slas_cuda_graph_destroy": [_vp],
    "slas_cuda_set_tunable": [C.c_char_p, C.c_long],
    "slas_cuda_get_tunable": [C.c_char_p, C.POINTER(C.c_long)],
    "slas_cuda_bench_fma_peak": [_i, C.POINTER(C.c_double)],
    "slas_cuda_bench_tf32_peak": [C.POINTER(C.c_double)],
    "slas_cuda_bench_call_loop": [_i, _sz, _vp, _vp, _vp, _i, _sz],
    "slas_cuda_bench_divby_mismatches": [_vp, _sz, C.POINTER(C.c_uint64)],
    "slas_cuda_get_reduce_phases": [C.POINTER(C.c_uint64)],
    "slas_cuda_malloc": [C.POINTER(_vp), _sz],
    "slas_cuda_free": [_vp],
    "slas_cuda_malloc_host": [C.POINTER(_vp), _sz],
    "slas_cuda_free_host": [_vp],
    "slas_cuda_host_register": [_vp, _sz],
    "slas_cuda_host_unregister": [_vp],
    "slas_cuda_memcpy_h2d": [_vp, _vp, _sz],
    "slas_cuda_memcpy_d2h": [_vp, _vp, _sz],
    "slas_cuda_memcpy_d2d": [_vp, _vp, _sz],
    "slas_cuda_memset": [_vp, _i, _sz],
    "slas_cuda_is_device_ptr": [_vp, _ip],
    "slas_cuda_sfill_uniform": [_sz, _vp, _u64, _u64, _f, _f],
    "slas_cuda_dfill_uniform": [_sz, _vp, _u64, _u64, _d, _d],
    "slas_cuda_sdot": [_sz, _vp, _vp, _vp],
    "slas_cuda_ddot": [_sz, _vp, _vp, _vp],
    "slas_cuda_cdotu": [_sz, _vp, _vp, _vp],
    "slas_cuda_zdotu": [_sz, _vp, _vp, _vp],
    "slas_cuda_snrm2": [_sz, _vp, _vp],
    "slas_cuda_dnrm2": [_sz, _vp, _vp],
    "slas_cuda_scnrm2": [_sz, _vp, _vp],
    "slas_cuda_dznrm2": [_sz, _vp, _vp],
    "slas_cuda_snormalize": [_sz, _vp],
    "slas_cuda_dnormalize": [_sz, _vp],
    "slas_cuda_cnormalize": [_sz, _vp],
    "slas_cuda_znormalize": [_sz, _vp],
    **{f"slas_cuda_{t}{op}": [_sz, _vp, _vp, _vp] for t in "sd" for op in ("add", "sub", "mul", "div")},
    "slas_cuda_transpose": [_sz, _sz, _vp, _vp, _sz],
    "slas_cuda_transpose_inplace": [_sz, _sz, _vp, _sz],
    "slas_cuda_sgemm": [_vp, _vp, _vp, _sz, _sz, _sz, _sz, _sz, _sz, _i, _i],
    "slas_cuda_dgemm": [_vp, _vp, _vp, _sz, _sz, _sz, _sz, _sz, _sz, _i, _i],
    "slas_cuda_set_sgemm_path": [_i],
    "slas_cuda_get_sgemm_path": [_ip],
    "slas_cuda_sgemv": [_vp, _vp, _vp, _sz, _sz, _sz, _i],
    "slas_cuda_dgemv": [_vp, _vp, _vp, _sz, _sz, _sz, _i],
    "slas_cuda_sgemm_batched": [_sz, _vp, _vp, _vp, _sz, _sz, _sz, _i, _i],
    "slas_cuda_dgemm_batched": [_sz, _vp, _vp, _vp, _sz, _sz, _sz, _i, _i],
    "slas_cuda_sgemv_batched": [_sz, _vp, _vp, _vp, _sz, _sz, _i],
    "slas_cuda_dgemv_batched": [_sz, _vp, _vp, _vp, _sz, _sz, _i],
    "slas_cuda_transpose_batched": [_sz, _sz, _sz, _vp, _vp, _sz],
    **{f"slas_cuda_{t}{op}_dot": [_sz, _vp, _vp, _vp, _vp, _vp] for t in "sd" for op in ("add", "sub", "mul", "div")},
    "slas_cuda_sdot_norms": [_sz, _vp, _vp, _vp],
    "slas_cuda_ddot_norms": [_sz, _vp, _vp, _vp],
    "slas_cuda_snormalize_into": [_sz, _vp, _vp],
    "slas_cuda_dnormalize_into": [_sz, _vp, _vp],
    "slas_cuda_transpose_inplace_batched": [_sz, _sz, _sz, _vp, _sz],
    "slas_cuda_sdot_batched": [_sz, _sz, _vp, _vp, _vp],
    "slas_cuda_ddot_batched": [_sz, _sz, _vp, _vp, _vp],
    "slas_cuda_snrm2_batched": [_sz, _sz, _vp, _vp],
    "slas_cuda_dnrm2_batched": [_sz, _sz, _vp, _vp],
    "slas_cuda_snormalize_batched": [_sz, _sz, _vp],
    "slas_cuda_dnormalize_batched": [_sz, _sz, _vp],
    "slas_cuda_init_devices": [_i],
    "slas_cuda_devices": [_ip, _ip],
    "slas_cuda_malloc_on": [_i, C.POINTER(_vp), _sz],
    "slas_cuda_sdot_multi": [_i, _szp, C.POINTER(_vp), C.POINTER(_vp), _vp],
    "slas_cuda_ddot_multi": [_i, _szp, C.POINTER(_vp), C.POINTER(_vp), _vp],
    "slas_cuda_snrm2_multi": [_i, _szp, C.POINTER(_vp), _vp],
    "slas_cuda_dnrm2_multi": [_i, _szp, C.POINTER(_vp), _vp],
    "slas_cuda_snormalize_multi": [_i, _szp, C.POINTER(_vp)],
    "slas_cuda_dnormalize_multi": [_i, _szp, C.POINTER(_vp)],
    "slas_cuda_nccl_unique_id": [_vp],
    "slas_cuda_comm_init": [_i, _i, _vp],
    "slas_cuda_comm_destroy": [],
    "slas_cuda_comm_info": [_ip, _ip],
    "slas_cuda_p2p_handle": [_vp],
    "slas_cuda_p2p_init": [_i, _i, _vp],
    "slas_cuda_p2p_destroy": [],
    "slas_cuda_shard_range": [_sz, _sz, _i, _i, _szp, _szp],
    "slas_cuda_sdot_sharded": [_sz, _vp, _vp, _vp],
    "slas_cuda_ddot_sharded": [_sz, _vp, _vp, _vp],
    "slas_cuda_snrm2_sharded": [_sz, _vp, _vp],
    "slas_cuda_dnrm2_sharded": [_sz, _vp, _vp],
    "slas_cuda_snormalize_sharded": [_sz, _vp],
    "slas_cuda_dnormalize_sharded": [_sz, _vp],
    "slas_cuda_sdot_partial": [_sz, _vp, _vp, _vp],
    "slas_cuda_ddot_partial": [_sz, _vp, _vp, _vp],
}
_RESTYPES = {"slas_cuda_last_error": C.c_char_p, "slas_cuda_launch_count": C.c_uint64}

_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Loading needs no GPU; every compute call does."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()). "
                "slas_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header and library disagree
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != SLAS_OK:
        msg = lib().slas_cuda_last_error()
        raise SlasCudaError(status, msg.decode() if msg else "")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))
