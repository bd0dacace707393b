"""Development aid: v2 FFT-spectrum kernel time vs number of persistent CTAs (L2 footprint experiment)."""
import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from dynaphopy_b200 import ops as dops
    ops = dops.Ops()
    S = int(os.environ.get("PC_S", 1480))
    x = torch.randn((S, 131072), dtype=torch.complex128, device=ops.be.device)
    freq = np.arange(0, 40.05, 0.05)
    for _ in range(2):
        ops.power_fft(x, 0.001, freq, series_major=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        ops.power_fft(x, 0.001, freq, series_major=True)
    b.record(); torch.cuda.synchronize()
    print("grid=%s  %.3f ms" % (os.environ.get("DP_S2_GRID", "148"), a.elapsed_time(b) / 3), flush=True)
else:
    for g in [int(x) for x in os.environ.get("S2_GRIDS", "148,128,111,96,74,48,37").split(",")]:
        env = dict(os.environ, DP_S2_GRID=str(g))
        subprocess.run([sys.executable, __file__, "child"], env=env)
