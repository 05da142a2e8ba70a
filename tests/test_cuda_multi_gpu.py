"""Multi-GPU parity (`pytest -m gpu` on a box with >= 2 GPUs; skipped otherwise): launches
tests/_multi_gpu_worker.py under torchrun, one rank per GPU."""
import os
import subprocess
import sys

import pytest

from slas_b200.backends import device_count

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_ops_across_gpus(world):
    if device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_multi_gpu_worker.py")],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-600:])  # `pytest -s` keeps the MULTI_GPU_OK line (with the observed error / bar ratios) in the log
    assert r.returncode == 0 and f"MULTI_GPU_OK {world}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("ndev", [2, 8])
def test_single_process_several_devices(ndev):
    """ONE ordinary process driving several GPUs (slas_cuda_init_devices): host-operand splits, the in-process sharded
    reductions over peer memory, device-pointer ops on the owning device (tests/_multi_device_worker.py)"""
    if device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_multi_device_worker.py"), str(ndev)],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-600:])
    assert r.returncode == 0 and f"MULTI_DEVICE_OK {ndev}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
