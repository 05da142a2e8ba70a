"""The reference's own hot-path tests (tests/src/main.rs) replayed in C++ against slas_backend::Cuda through
the typed host mirror include/slas_cuda.hpp (the reference is compiled code, so its host-side mirror is
too).  The binary is built by __graft_entry__.build() / `make -C tests/cpp`."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_reference_suite")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"], check=True)


def test_cpp_suite_builds_against_the_header():
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_cpp_reference_suite_on_gpu():
    if not os.path.exists(BIN):
        _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "test result: ok." in r.stdout and "0 failed" in r.stdout
    assert r.stdout.count(" ... ok") >= 25
