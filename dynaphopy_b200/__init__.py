"""dynaphopy_b200 -- B200-native (sm_100a) implementation of DynaPhoPy's normal-mode-decomposition
hot path: wave-vector / phonon projection, FFT / maximum-entropy / direct-correlation power
spectra and the minimum-image displacement kernel, behind the reference's own API.

Layers
  csrc/            hand-written CUDA kernels + the C ABI (include/dynaphopy_b200.h)
  _lib, ops        ctypes binding and array-level operators (PyTorch supplies device memory/streams)
  structure, dynamics, parameters, projection, power_spectrum, quasiparticle
                   host-side mirror of the reference interface (same names and arguments)
  pipeline         whole q-mesh driver: time-sharded projection, one all-to-all, series-sharded spectra

There is no CPU fallback: importing works anywhere, computing needs a CUDA device.
"""
__version__ = "0.1.0"

from . import _lib, ops  # noqa: F401
