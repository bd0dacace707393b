"""``power_spectrum`` with the reference's method selector (dynaphopy/power_spectrum/__init__.py:439-446).

Every entry keeps the reference call signature ``fn(vq, trajectory, parameters) -> (F, vq.shape[1]) float64``
(already multiplied by ``unit_conversion``, :9) and its label; ids 0 / 1 / 2 keep their meaning (direct
correlation, maximum entropy, FFT) and the FFT aliases 3 (FFTW), 4 (CUDA), 5 (CuPy) -- the same arithmetic
with another FFT library in the reference -- are served by the same kernel.  All columns are processed in one
launch instead of a Python loop over modes.
"""
import numpy as np

from . import _lib
from . import ops as _ops

unit_conversion = _ops.UNIT_CONVERSION      # u * A^2 * THz -> eV*ps


def _ops_of(trajectory):
    return getattr(trajectory, "ops", None) or _ops.default_ops()


def _division_of_data(resolution, number_of_data, time_step):
    """Window list of the FFT method (power_spectrum/__init__.py:33-51), from the library's planner."""
    grid = np.array([0.0, resolution])
    return _ops.default_ops().fft_pieces(number_of_data, time_step, grid)[1]


def get_fourier_direct_power_spectra(vq, trajectory, parameters):
    """power_spectrum/__init__.py:57-75 -> correlation.correlation_par (c/correlation.c:101-178).  The
    shipped weights exp(w*tau + 1i) are reproduced unless ``parameters.correlation_weight == 'intended'``."""
    weight = _lib.DP_CORR_WEIGHT_INTENDED if getattr(parameters, "correlation_weight", "reference") == "intended" \
        else _lib.DP_CORR_WEIGHT_REFERENCE
    out = _ops_of(trajectory).power_correlation(vq, trajectory.get_time_step_average(), parameters.frequency_range,
                                                step=parameters.correlation_function_step,
                                                integration_method=parameters.integration_method, weight=weight)
    return _ops.to_host(out)


def get_mem_power_spectra(vq, trajectory, parameters):
    """power_spectrum/__init__.py:81-103 -> mem.mem (c/mem.c:102-252), Data[N] := 0, nan_to_num applied."""
    if _ops.shape_of(vq)[0] <= parameters.number_of_coefficients_mem + 1:
        print('Number of coefficients should be smaller than the number of time steps')      # :85-87
        raise SystemExit
    out = _ops_of(trajectory).power_mem(vq, trajectory.get_time_step_average(), parameters.frequency_range,
                                        parameters.number_of_coefficients_mem)
    return _ops.to_host(out)


def get_fft_numpy_spectra(vq, trajectory, parameters):
    """power_spectrum/__init__.py:226-270 (np.correlate + np.fft + np.interp arithmetic)."""
    out = _ops_of(trajectory).power_fft(vq, trajectory.get_time_step_average(), parameters.frequency_range)
    return _ops.to_host(out)


get_fft_fftw_power_spectra = get_fft_numpy_spectra
get_fft_cuda_power_spectra = get_fft_numpy_spectra
get_fft_cupy_spectra = get_fft_numpy_spectra

power_spectrum_functions = {
    0: [get_fourier_direct_power_spectra, 'Fourier transform'],
    1: [get_mem_power_spectra, 'Maximum entropy method'],
    2: [get_fft_numpy_spectra, 'Fast Fourier transform (Numpy)'],
    3: [get_fft_fftw_power_spectra, 'Fast Fourier transform (FFTW)'],
    4: [get_fft_cuda_power_spectra, 'Fast Fourier transform (CUDA)'],
    5: [get_fft_cupy_spectra, 'Fast Fourier transform (CuPy)'],
}
