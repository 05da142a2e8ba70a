"""Worker of tests/test_cuda_multi_gpu.py::test_single_process_several_devices: ONE ordinary process drives N GPUs
(slas_cuda_init_devices) -- what a slas program is (a backend is a ZST made by B::default(), src/static_vec.rs:76,219).
Checks the host-operand splits against the oracle / an f64 truth (bit-exact where the op is), the device-resident
*_multi reductions (peer-memory exchange inside the kernels) against the rank-ordered sum of the slices' own f64 partials,
and that an op on device operands runs on the device that owns them.  Prints MULTI_DEVICE_OK <n>."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from slas_b200 import _lib as L  # noqa: E402
from slas_b200 import sharding  # noqa: E402
from slas_b200.backends import Cuda, init_devices, launch_count, set_tunable  # noqa: E402
from slas_b200.static_vec import CudaVec  # noqa: E402


def main():
    want = int(sys.argv[1])
    n = init_devices(want)
    assert n == want, (n, want)
    nd, px = C.c_int(), C.c_int()
    L.call("slas_cuda_devices", C.byref(nd), C.byref(px))
    cuda = Cuda()
    rng = np.random.default_rng(11)
    set_tunable("multi_split_mb", 1)  # split already at test sizes (default: 64 MB)

    # ---- host operands, split over the devices ----
    batch, size = 4099, 16
    e = size * size
    A = (rng.random(batch * e, dtype=np.float32) * 2 - 1)
    B = (rng.random(batch * e, dtype=np.float32) * 2 - 1)
    Cm = np.zeros(batch * e, np.float32)
    n0 = launch_count()
    cuda.matrix_mul_batched(A, B, Cm, size, size, size)
    assert launch_count() - n0 >= n, "the batch was not split over the devices"
    ref = oracle.gemm_batched(A, B, batch, size, size, size)
    mag = np.einsum("bik,bkj->bij", np.abs(A.reshape(-1, size, size)).astype(np.float64), np.abs(B.reshape(-1, size, size)).astype(np.float64)).ravel()
    assert np.all(np.abs(Cm.astype(np.float64) - ref) <= 1e-5 * np.sqrt(size) * mag)
    T_ = np.zeros_like(A)
    cuda.transpose_batched(A, T_, e, size)
    assert np.array_equal(T_.reshape(batch, size, size), A.reshape(batch, size, size).transpose(0, 2, 1))
    ln = (1 << 21) + 5
    x = (rng.random(ln, dtype=np.float32) * 2 - 1)
    y = rng.random(ln, dtype=np.float32)
    z = np.zeros_like(x)
    for op in ("add", "sub", "mul", "div"):
        getattr(cuda, op)(x, y + 0.5, z)
        assert np.array_equal(z, oracle.ewise(op, x, (y + 0.5).astype(np.float32))), op
    scale = float(np.sum(np.abs(x.astype(np.float64) * y)))
    d = cuda.dot(x, y)
    assert abs(float(d) - float(np.dot(x.astype(np.float64), y.astype(np.float64)))) <= 1e-5 * np.sqrt(ln) * scale
    nr = cuda.norm(x)
    truth_n = float(np.sqrt(np.dot(x.astype(np.float64), x.astype(np.float64))))
    assert abs(float(nr) - truth_n) <= 1e-5 * np.sqrt(ln) * truth_n
    xn = x.copy()
    cuda.normalize(xn)
    assert np.array_equal(xn, x / nr), "host normalize over several devices is not x / norm bit for bit"
    vb, ob = (rng.random(5000 * 64, dtype=np.float32) * 2 - 1), np.zeros(5000, np.float32)
    cuda.dot_batched(vb, vb, ob, 64)
    tr = np.einsum("bi,bi->b", vb.reshape(5000, 64).astype(np.float64), vb.reshape(5000, 64).astype(np.float64))
    assert np.all(np.abs(ob - tr) <= 1e-5 * 8 * tr)

    # ---- device-resident slices, one per device: the in-process sharded reductions ----
    total = (1 << 22) + 12
    a = rng.random(total, dtype=np.float32)
    b = rng.random(total, dtype=np.float32) * 2 - 1
    sa, sb = [], []
    for r in range(n):
        lo, hi = sharding.shard_range(total, 4, n, r)
        sa.append(CudaVec.from_host(a[lo:hi], device=r))
        sb.append(CudaVec.from_host(b[lo:hi], device=r))
    part = np.zeros(1, np.float64)
    s_dot = s_sq = 0.0
    for r in range(n):  # slice order, like the exchange
        L.call("slas_cuda_sdot_partial", sa[r].LEN, sa[r].as_ptr(), sb[r].as_ptr(), part.ctypes.data)
        s_dot += float(part[0])
        L.call("slas_cuda_sdot_partial", sa[r].LEN, sa[r].as_ptr(), sa[r].as_ptr(), part.ctypes.data)
        s_sq += float(part[0])
    truth = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    sc = float(np.sum(np.abs(a.astype(np.float64) * b)))
    for rep in range(40):  # both mailbox parities, back to back
        d = cuda.dot_multi(sa, sb)
        assert np.float32(d).tobytes() == np.float32(s_dot).tobytes(), (rep, float(d), s_dot)
        nr = cuda.norm_multi(sa)
        assert np.float32(nr).tobytes() == np.float32(np.sqrt(s_sq)).tobytes(), (rep, float(nr), np.sqrt(s_sq))
    assert abs(float(d) - truth) <= 1e-5 * np.sqrt(total) * sc
    cuda.normalize_multi(sa)
    for r in range(n):
        lo, hi = sharding.shard_range(total, 4, n, r)
        assert np.array_equal(sa[r].to_host(), a[lo:hi] / nr), f"slice {r} of normalize_multi is not a / norm bit for bit"
    assert abs(float(cuda.norm_multi(sa)) - 1.0) < 1e-5
    # slices of different lengths, one of them empty
    lens = [1000, 0, 7, 4096, 1, 33, 129, 2][:n]
    ra = [CudaVec.from_host(np.full(k, 2.0, np.float32), device=r) for r, k in enumerate(lens)]
    assert float(cuda.dot_multi(ra, ra)) == 4.0 * sum(lens)

    # ---- an op on device operands runs on the device that owns them, whatever the current device is ----
    u = rng.random(100003, dtype=np.float32)
    for r in range(n):
        du, dv = CudaVec.from_host(u, device=r), CudaVec(u.size, np.float32, device=r)
        cuda.add(du, du, dv)
        assert np.array_equal(dv.to_host(), u + u), r
        assert abs(float(cuda.dot(du, du)) - float(np.dot(u.astype(np.float64), u.astype(np.float64)))) <= 1e-5 * np.sqrt(u.size) * float(np.dot(u, u))
    print("MULTI_DEVICE_OK", n, "peer_exchange" if px.value else "host_combine")


if __name__ == "__main__":
    main()
