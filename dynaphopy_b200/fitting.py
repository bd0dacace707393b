"""The part of the reference's fitting front end that decides WHAT is fitted (host logic, no kernels):
degenerate sets of harmonic frequencies (dynaphopy/analysis/fitting/__init__.py:17-33).  The averaging and the
initial guess themselves run on the device (``Ops.spectra_peaks`` / ``dp_spectra_peaks``); Lorentzian fitting stays in
the reference's scipy code."""
import numpy as np


def degenerate_sets(freqs, cutoff=1e-4):
    """Greedy grouping in index order: j joins the set of i when it is within ``cutoff`` of ANY member so far."""
    freqs = np.asarray(freqs, dtype=float)
    indices, done = [], []
    for i in range(len(freqs)):
        if i in done:
            continue
        f_set = [i]
        done.append(i)
        for j in range(i + 1, len(freqs)):
            if (np.abs(freqs[f_set] - freqs[j]) < cutoff).any():
                f_set.append(j)
                done.append(j)
        indices.append(f_set[:])
    return indices
