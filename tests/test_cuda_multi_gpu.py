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
    assert r.returncode == 0 and f"MULTI_GPU_OK {world}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
