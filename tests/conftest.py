import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def to_np(x):
    if isinstance(x, np.ndarray):
        return x
    return x.detach().cpu().numpy()


def relerr(a, b):
    """max|a-b| / max|b| over the entries where b is finite (normwise relative error)."""
    a, b = to_np(a), to_np(b)
    fin = np.isfinite(b)
    if not fin.any():
        return 0.0
    return float(np.abs(a - b)[fin].max() / max(np.abs(b[fin]).max(), 1e-300))


def _cuda_ops():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dynaphopy_b200 import ops
    return ops.default_ops()


@pytest.fixture(params=["emul", pytest.param("cuda", marks=pytest.mark.gpu)])
def kops(request):
    """The kernels under test: host emulation of the .cu sources (CPU suite) or the real
    sm_100a library through the C ABI (GPU suite)."""
    if request.param == "emul":
        from emul import emul_ops
        return emul_ops()
    return _cuda_ops()


@pytest.fixture
def cuda_ops():
    return _cuda_ops()
