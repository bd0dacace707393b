"""TEST INFRASTRUCTURE -- drive the kernels' host emulation build from numpy.

``emul_ops()`` returns a ``dynaphopy_b200.ops.Ops`` bound to tests/_emul_build/libdp_emul.so (the
SAME .cu sources compiled with -DDP_EMUL, see tests/emul_include/dp_emul.h) and a numpy
back-end, so that the CPU test-suite exercises the real kernel code paths (index algebra,
barriers, shuffles) without a GPU.  Nothing in the dynaphopy_b200 package refers to this.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dynaphopy_b200 import _lib, build, ops  # noqa: E402


class NumpyBackend:
    def empty(self, shape, dtype):
        return np.empty(tuple(int(s) for s in shape), dtype=dtype)

    def asarray(self, x, dtype):
        return np.ascontiguousarray(x, dtype=dtype)

    def is_complex(self, x):
        return np.iscomplexobj(x)

    def ptr(self, x):
        return x.ctypes.data_as(ctypes.c_void_p) if x is not None else None

    def stream(self):
        return None

    def shape(self, x):
        return tuple(x.shape)


_ops = None


def emul_ops():
    global _ops
    if _ops is None:
        path = build.build_emulator()
        _ops = ops.Ops(lib=_lib.DpLib(path), backend=NumpyBackend())
    return _ops
