"""Parameter holder with the reference's attribute names and defaults for the fields the hot path
reads (dynaphopy/parameters.py:14-102).  The reference's own ``Parameters`` object can be used
instead; like the reference this is a process-wide singleton (parameters.py:4-14)."""
import numpy as np


class _Singleton(type):
    _instances = {}

    def __call__(cls, *args, **kwargs):
        if cls not in cls._instances:
            cls._instances[cls] = super().__call__(*args, **kwargs)
        return cls._instances[cls]


class Parameters(metaclass=_Singleton):
    def __init__(self):
        self.reset()

    def reset(self):
        self.silent = False
        self.reduced_q_vector = np.array((0, 0, 0))
        self.number_of_coefficients_mem = 1000          # parameters.py:24
        self.correlation_function_step = 10             # parameters.py:28
        self.integration_method = 1                     # 0: trapezoid  1: rectangles (parameters.py:29)
        self.zero_padding = 0
        self.power_spectra_algorithm = 1                # MEM is the reference default (parameters.py:39)
        self.spectrum_resolution = 0.05
        self.frequency_range = np.arange(0, 40.05, 0.05)
        self.use_symmetry = True
        self.project_on_atom = -1
        # extension: weights of the direct method, "reference" (as shipped) or "intended"
        self.correlation_weight = "reference"
