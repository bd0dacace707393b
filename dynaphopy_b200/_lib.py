"""ctypes binding of the C ABI declared in ``include/dynaphopy_b200.h``.

``load()`` returns the CUDA library (``libdynaphopy_b200.so`` next to this file, built by
``dynaphopy_b200.build``).  There is no CPU implementation behind it: if the library cannot be
loaded, or no CUDA device is present when a compute entry point is used, an exception is raised.
"""
import ctypes
import os

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_L = _c.c_int64
_D = _c.c_double
_Z = _c.c_size_t

# name -> (restype, argtypes); mirrors include/dynaphopy_b200.h one to one
SIGNATURES = {
    "dp_version": (_c.c_char_p, []),
    "dp_last_error": (_c.c_char_p, []),
    "dp_sm_count": (_I, []),
    "dp_launch_count": (_c.c_longlong, []),
    "dp_atomic_displacements": (_I, [_P, _I, _P, _P, _L, _I, _P, _P]),
    "dp_velocity_from_positions": (_I, [_P, _I, _P, _P, _P, _L, _I, _D, _P, _P, _P, _P, _P]),
    "dp_mass_weight": (_I, [_P, _I, _P, _L, _I, _P, _P]),
    "dp_displacement_moments_workspace_bytes": (_Z, [_L, _I]),
    "dp_displacement_moments": (_I, [_P, _I, _P, _P, _L, _I, _P, _P, _Z, _P]),
    "dp_project_wave_vector_workspace_bytes": (_Z, [_I, _I]),
    "dp_project_wave_vector": (_I, [_P, _I, _P, _P, _P, _I, _I, _L, _P, _I, _I, _P, _P, _Z, _P]),
    "dp_project_commensurate_workspace_bytes": (_Z, [_P, _I, _I, _I]),
    "dp_project_commensurate": (_I, [_P, _P, _I, _P, _I, _L, _P, _P, _I, _I, _P, _P, _Z, _P]),
    "dp_project_commensurate_ex": (_I, [_P, _P, _I, _P, _I, _L, _P, _P, _I, _I, _I, _P, _P, _Z, _P]),
    "dp_project_commensurate_blocks": (_I, [_P, _P, _I, _P, _I, _L, _P, _I, _I, _P, _I, _P, _Z, _P]),
    "dp_project_phonon_series_ex": (_I, [_P, _I, _P, _L, _I, _I, _I, _L, _L, _P, _P]),
    "dp_project_phonon_series_peers": (_I, [_P, _I, _P, _L, _I, _I, _I, _L, _L, _P, _I, _P]),
    "dp_project_phonon_series_mirror": (_I, [_P, _P, _P, _L, _I, _I, _I, _I, _L, _L, _P, _P, _I, _P]),
    "dp_project_phonon": (_I, [_P, _P, _L, _I, _I, _P, _P]),
    "dp_project_phonon_blocked": (_I, [_P, _P, _L, _I, _I, _I, _P, _P]),
    "dp_project_phonon_series": (_I, [_P, _P, _L, _I, _I, _I, _L, _L, _P, _P]),
    "dp_power_fft_workspace_bytes": (_Z, [_L, _I, _D, _P, _I]),
    "dp_power_fft": (_I, [_P, _L, _I, _L, _D, _P, _I, _D, _P, _P, _Z, _P]),
    "dp_power_fft_strided": (_I, [_P, _L, _I, _L, _L, _L, _L, _D, _P, _I, _D, _P, _P, _Z, _P]),
    "dp_power_fft_windows": (_I, [_P, _L, _I, _L, _L, _L, _L, _D, _P, _I, _I, _I, _P, _Z, _P]),
    "dp_power_fft_windows_ex": (_I, [_P, _L, _I, _L, _L, _L, _L, _D, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "dp_power_fft_windows_preemptible": (_I, [_P, _L, _I, _L, _L, _L, _L, _D, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "dp_power_fft_finish": (_I, [_L, _I, _D, _P, _I, _D, _P, _P, _Z, _P]),
    "dp_power_fft_pieces": (_I, [_L, _D, _P, _I, _P, _P, _P, _I]),
    "dp_power_fft_engine": (_I, [_L, _D, _P, _I]),
    "dp_power_mem_workspace_bytes": (_Z, [_L, _I, _I, _I]),
    "dp_power_mem": (_I, [_P, _L, _I, _L, _D, _P, _I, _I, _D, _I, _P, _P, _Z, _P]),
    "dp_power_mem_strided": (_I, [_P, _L, _I, _L, _L, _L, _L, _D, _P, _I, _I, _D, _I, _P, _P, _Z, _P]),
    "dp_power_correlation_strided": (_I, [_P, _L, _I, _L, _L, _L, _L, _D, _P, _I, _I, _I, _I, _D, _P, _P, _Z, _P]),
    "dp_power_correlation_workspace_bytes": (_Z, [_L, _I, _I, _I]),
    "dp_power_correlation": (_I, [_P, _L, _I, _L, _D, _P, _I, _I, _I, _I, _D, _P, _P, _Z, _P]),
    "dp_spectra_peaks": (_I, [_P, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P]),
    "dp_lammps_scan": (_I, [_c.c_char_p, _P, _P, _P, _P, _c.c_char_p, _I]),
    "dp_lammps_read": (_I, [_c.c_char_p, _L, _L, _I, _I, _P, _P, _I]),
    "dp_xdatcar_scan": (_I, [_c.c_char_p, _P, _P, _P]),
    "dp_xdatcar_read": (_I, [_c.c_char_p, _L, _L, _I, _P, _P, _L, _I]),
    "dp_fft_c2c_workspace_bytes": (_Z, [_I, _I]),
    "dp_fft_c2c": (_I, [_P, _P, _I, _I, _I, _P, _Z, _P]),
}

DP_ERR_INVALID = -1
DP_VC_ROWS = 0
DP_VC_PAIRS = 1
DP_CORR_WEIGHT_REFERENCE = 0
DP_CORR_WEIGHT_INTENDED = 1


class DpError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"dynaphopy_b200 error {code}: {text}")
        self.code = code


class DpLib:
    """A loaded C-ABI library with typed entry points; ``call`` raises DpError on failure."""

    def __init__(self, path):
        self.path = path
        self.cdll = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.cdll, name)      # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args

    def call(self, name, *args):
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise DpError(rc, self.cdll.dp_last_error().decode())

    def size(self, name, *args):
        return int(getattr(self.cdll, name)(*args))

    def version(self):
        return self.cdll.dp_version().decode()


LIB_PATH = os.environ.get("DYNAPHOPY_B200_LIB",        # development override: an alternative build of the SAME library
                          os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdynaphopy_b200.so"))
_lib = None


def load():
    """Load (building first if the .so is absent and nvcc exists) the CUDA library."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from . import build
            build.build_cuda()
        _lib = DpLib(LIB_PATH)
    return _lib
