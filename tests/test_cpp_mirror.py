"""The C++ host-side mirror of the reference interface (include/plonkish_cuda.hpp): it compiles and links against
the C ABI (CPU), has no CPU path, and matches the oracle bit for bit on the GPU (tests/cpp/test_mirror.cpp)."""
import os
import subprocess

import pytest

from conftest import ROOT

CPP_DIR = os.path.join(ROOT, "tests", "cpp")


@pytest.fixture(scope="module")
def binary(oracle):
    from plonkish_b200 import build

    build.build_library()
    subprocess.run(["make", "-C", CPP_DIR], check=True, capture_output=True)
    return os.path.join(CPP_DIR, "test_mirror")


def test_mirror_compiles_links_and_has_no_cpu_path(binary):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by the gpu test")
    res = subprocess.run([binary], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2 and "plonkish_cuda_init failed" in res.stderr, (res.returncode, res.stderr)


@pytest.mark.gpu
def test_mirror_matches_the_oracle_on_the_gpu(binary):
    res = subprocess.run([binary], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "cpp mirror ok" in res.stdout, (res.returncode, res.stdout, res.stderr)
