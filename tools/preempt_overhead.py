"""Cost of the preemptible form of the slab spectrum kernel (one short-lived CTA per item) against the persistent form,
cfg 5-A windows, one GPU, nothing else running."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaphopy_b200 import ops as dops
ops = dops.Ops()
S, T, dt = 1480, 131072, 0.001
freq = np.arange(0, 40.05, 0.05)
x = torch.randn((S, T), dtype=torch.complex128, device=ops.be.device)
ws = ops.power_fft_begin(T, S, dt, freq)
n = len({tuple(p) for p in ops.fft_pieces(T, dt, freq)[1]})
for pre in (False, True, False, True):
    for _ in range(2):
        ops.power_fft_windows(x, dt, freq, 0, n, ws, series_major=True, preemptible=pre)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        ops.power_fft_windows(x, dt, freq, 0, n, ws, series_major=True, preemptible=pre)
    b.record(); torch.cuda.synchronize()
    print("preemptible=%s  %.3f ms per %d series x %d windows" % (pre, a.elapsed_time(b) / 3, S, n), flush=True)
