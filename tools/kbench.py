"""Per-kernel timing on one GPU (CUDA events, after warm-up).  Development aid, not the bench."""
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaphopy_b200 import ops as dops  # noqa: E402

ops = dops.Ops()
dev = ops.be.device
HBM = 6557.1e9


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


def report(name, t, nbytes=None, flops=None, extra=""):
    line = {"kernel": name, "ms": round(t[0] * 1e3, 4), "min_ms": round(t[1] * 1e3, 4)}
    if nbytes:
        line["GBps"] = round(nbytes / t[0] / 1e9, 1)
        line["hbm_frac"] = round(nbytes / t[0] / HBM, 3)
    if flops:
        line["TFLOPs"] = round(flops / t[0] / 1e12, 2)
    if extra:
        line["note"] = extra
    print(json.dumps(line), flush=True)


which = set(sys.argv[1:]) or {"disp", "proj", "phonon", "fft", "mem", "corr", "c2c", "dense"}
g = torch.Generator(device=dev); g.manual_seed(1)

dims = np.array([16, 16, 20]); nc = int(np.prod(dims)); N = 2 * nc
T = int(os.environ.get("KB_T", 2048))
cell = np.diag([4.0] * 3)
unit_pos = np.array([[0, 0, 0], [2.0, 2, 2]])
import itertools
pos = np.array([unit_pos[k] + np.array([x, y, z]) @ cell for k in range(2) for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])])
supercell = np.diag(dims) @ cell
sm = torch.sqrt(torch.tensor(np.repeat([28.0855, 69.723], nc), device=dev))

if "disp" in which:
    traj = torch.as_tensor(pos, device=dev)[None] + 0.05 * torch.randn((T, N, 3), dtype=torch.float64, device=dev, generator=g)
    t = timeit(lambda: ops.velocity_from_positions(traj, pos, supercell, 0.001, sqrt_mass=sm))
    report("velocity_from_positions", t, nbytes=48.0 * N * T)
    t = timeit(lambda: ops.atomic_displacements(traj, pos, supercell))
    report("atomic_displacements", t, nbytes=48.0 * N * T)
    t = timeit(lambda: ops.displacement_moments(traj, pos, supercell))
    report("displacement_moments", t, nbytes=24.0 * N * T)
    del traj

v = torch.randn((T, N, 3), dtype=torch.float64, device=dev, generator=g)
mm = np.array(list(itertools.product(range(dims[0]), range(dims[1]), range(dims[2]))), dtype=np.int32)
Q = len(mm)
q_cart = 2 * np.pi * (mm / dims) @ np.linalg.inv(cell).T
tw = torch.as_tensor(np.sqrt([28.0855, 69.723])[None, :] * np.exp(-1j * q_cart @ unit_pos.T) / np.sqrt(nc), device=dev)
qb = torch.as_tensor(mm, device=dev)
vc = None
if "proj" in which:
    t = timeit(lambda: ops.project_commensurate(v, dims, [0, 1], 2, qb, tw), reps=3, warm=1)
    report("project_commensurate", t, nbytes=(24.0 * N + 96.0 * Q) * T, extra=f"T={T} N={N} Q={Q}")
    vc = ops.project_commensurate(v, dims, [0, 1], 2, qb, tw)
if "dense" in which:
    Td = 256
    typ = np.repeat([0, 1], nc).astype(np.int32)
    t = timeit(lambda: ops.project_wave_vector(v[:Td], sm, pos, typ, 2, q_cart[:256]), reps=3, warm=1)
    report("project_wave_vector(dense)", t, flops=12.0 * N * Td * 256, extra=f"T={Td} Q=256")
    Tb, Qb = min(T, 2048), 512
    t = timeit(lambda: ops.project_wave_vector(v[:Tb], sm, pos, typ, 2, q_cart[:Qb]), reps=3, warm=1)
    report("project_wave_vector(dense)", t, flops=12.0 * N * Tb * Qb, extra=f"T={Tb} Q={Qb} (fp64 peak 36.6 TFLOP/s: frac {12.0 * N * Tb * Qb / t[0] / 36.6e12:.3f})")
    vcx = torch.randn((512, N, 3), dtype=torch.complex128, device=dev, generator=None)
    t = timeit(lambda: ops.project_wave_vector(vcx, sm, pos, typ, 2, q_cart[:Qb]), reps=3, warm=1)
    report("project_wave_vector(dense, complex v)", t, flops=24.0 * N * 512 * Qb, extra=f"T=512 Q={Qb}")
    del vcx
if "phonon" in which:
    if vc is None:
        vc = torch.randn((T, Q, 2, 3), dtype=torch.complex128, device=dev)
    eig = torch.randn((Q, 6, 6), dtype=torch.complex128, device=dev)
    t = timeit(lambda: ops.project_phonon(vc, eig), reps=3, warm=1)
    report("project_phonon", t, nbytes=2 * 96.0 * Q * T)
del v, vc
torch.cuda.empty_cache()

freq = np.arange(0, 40.05, 0.05)
if "c2c" in which:
    for n, b in ((4096, 4096), (20000, 1024), (32768, 1024), (131072, 256), (262144, 128)):
        x = torch.randn((b, n), dtype=torch.complex128, device=dev)
        t = timeit(lambda: ops.fft_c2c(x, -1), reps=3, warm=1)
        report(f"fft_c2c n={n} b={b}", t, nbytes=32.0 * n * b, flops=5.0 * n * np.log2(n) * b)
        t = timeit(lambda: torch.fft.fft(x, dim=1), reps=3, warm=1)
        report(f"cuFFT(torch) n={n} b={b}", t, nbytes=32.0 * n * b, flops=5.0 * n * np.log2(n) * b)
        del x
if "fft" in which:
    for (Ts, S, dt) in ((2500, 192, 0.002), (50001, 324, 0.001), (131072, int(os.environ.get("KB_S", 768)), 0.001)):
        x = torch.randn((Ts, S), dtype=torch.complex128, device=dev)
        t = timeit(lambda: ops.power_fft(x, dt, freq), reps=3, warm=1)
        L, pieces = ops.fft_pieces(Ts, dt, freq)
        report(f"power_fft T={Ts} S={S}", t, nbytes=16.0 * L * len(pieces) * S, extra=f"L={L} pieces={len(pieces)} series/s={S / t[0]:.1f}")
        xt = x.t().contiguous()
        t = timeit(lambda: ops.power_fft(xt, dt, freq, series_major=True), reps=3, warm=1)
        report(f"power_fft series-major T={Ts} S={S}", t, nbytes=16.0 * L * len(pieces) * S, extra=f"L={L} pieces={len(pieces)} series/s={S / t[0]:.1f}")
        del x, xt
if "fft5b" in which:
    # cfg 5-B: one window of the whole run, resolution 2^-17/dt (L = T = 131072, P = 2^18, 5244 frequencies)
    Ts, S, dt5 = 131072, int(os.environ.get("KB_S5B", 1024)), 0.001
    f5b = np.arange(5244) * (1.0 / (Ts * dt5))
    x = torch.randn((S, Ts), dtype=torch.complex128, device=dev)
    t = timeit(lambda: ops.power_fft(x, dt5, f5b, series_major=True), reps=3, warm=1)
    L, pieces = ops.fft_pieces(Ts, dt5, f5b)
    report(f"power_fft cfg5-B T={Ts} S={S}", t, nbytes=16.0 * L * S,
           extra=f"L={L} pieces={len(pieces)} engine={ops.fft_engine(Ts, dt5, f5b)} series/s={S / t[0]:.1f} full_job_30720_ms={30720 / S * t[0] * 1e3:.1f}")
    # what a library-FFT pipeline pays for its transforms alone: FFT_P, FFT_P, FFT_L per series (cuFFT through torch),
    # without |X|^2, the lag gather and the bin selection
    Sc = min(S, 256)
    xp = torch.zeros((Sc, 2 * Ts), dtype=torch.complex128, device=dev)
    xp[:, :Ts] = x[:Sc]

    def three_cufft():
        X = torch.fft.fft(xp, dim=1)
        Y = torch.fft.fft(X, dim=1)
        return torch.fft.fft(Y[:, :Ts], dim=1)
    t = timeit(three_cufft, reps=3, warm=1)
    report(f"3 x cuFFT(torch) per series, P=2^18,2^18,L=2^17, S={Sc}", t, extra=f"series/s={Sc / t[0]:.1f} full_job_30720_ms={30720 / Sc * t[0] * 1e3:.1f}")
    del x, xp
if "mem" in which:
    for (Ts, S, m) in ((2500, 192, 1000), (131072, int(os.environ.get("KB_SM", 148)), 1000)):
        x = torch.randn((Ts, S), dtype=torch.complex128, device=dev)
        t = timeit(lambda: ops.power_mem(x, 0.002, freq, m), reps=2, warm=1)
        report(f"power_mem T={Ts} S={S} m={m}", t, flops=20.0 * m * Ts * S * 2 / 2, extra=f"series/s={S / t[0]:.2f}")
        del x
if "corr" in which:
    for (Ts, S) in ((2500, 192), (32768, int(os.environ.get("KB_SC", 64))), (131072, int(os.environ.get("KB_SC", 64)))):
        x = torch.randn((Ts, S), dtype=torch.complex128, device=dev)
        t = timeit(lambda: ops.power_correlation(x, 0.001, freq, 10, 1, 1), reps=2, warm=1)
        report(f"power_correlation T={Ts} S={S}", t, flops=4.0 * Ts * Ts / 10 * S, extra=f"series/s={S / t[0]:.2f}")
        del x
