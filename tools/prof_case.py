"""One isolated call of a hot-path kernel at bench-relevant sizes, for `ncu --set full` captures.

    python tools/prof_case.py fft5a|fft5b|project|contract|dense|mem|corr [reps]
"""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaphopy_b200 import ops as dops  # noqa: E402

case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ops = dops.Ops()
dev = ops.be.device
freq = np.arange(0, 40.05, 0.05)
torch.manual_seed(0)

if case == "fft5a":
    S = int(os.environ.get("PC_S", 296))
    x = torch.randn((S, 131072), dtype=torch.complex128, device=dev)
    for _ in range(reps):
        out = ops.power_fft(x, 0.001, freq, series_major=True)
elif case == "fft5b":
    S = int(os.environ.get("PC_S", 64))
    x = torch.randn((S, 131072), dtype=torch.complex128, device=dev)
    f5b = np.arange(5244) * (1.0 / (131072 * 0.001))
    for _ in range(reps):
        out = ops.power_fft(x, 0.001, f5b, series_major=True)
elif case in ("project", "contract"):
    dims = np.array([16, 16, 20]); nc = int(np.prod(dims)); T = int(os.environ.get("PC_T", 4096))
    cell = np.diag([4.0] * 3); unit_pos = np.array([[0, 0, 0], [2.0, 2, 2]])
    mm = np.array([[x, y, z] for z in range(20) for y in range(16) for x in range(16)], dtype=np.int32)
    q_cart = 2 * np.pi * (mm / dims) @ np.linalg.inv(cell).T
    tw = torch.as_tensor(np.sqrt([28.0855, 69.723])[None, :] * np.exp(-1j * q_cart @ unit_pos.T) / np.sqrt(nc), device=dev)
    qb = torch.as_tensor(mm, device=dev)
    v = torch.randn((T, 2 * nc, 3), dtype=torch.float64, device=dev)
    eig = torch.randn((len(mm), 6, 6), dtype=torch.complex128, device=dev)
    # as the pipeline runs it: bare lattice DFT of one wave vector per pair (q, -q) in the pair-major intermediate
    # layout, both contracted (PC_FULL=1: the whole mesh)
    from dynaphopy_b200.pipeline import mirror_pairs
    qmap = mirror_pairs(mm, dims)
    e_np = eig.cpu().numpy()
    eig2 = np.zeros((len(qmap), 2, 6, 6), dtype=complex)
    eig2[:, 0] = e_np[qmap[:, 0]]
    has = qmap[:, 1] >= 0
    eig2[has, 1] = np.conj(e_np[qmap[has, 1]])
    eig2 = torch.as_tensor(eig2, device=dev)
    qmap_d = torch.as_tensor(qmap, device=dev)
    qb_half = torch.as_tensor(mm[qmap[:, 0]], device=dev)
    full = os.environ.get("PC_FULL") == "1"
    for _ in range(reps):
        vc = ops.project_commensurate(v, dims, [0, 1], 2, qb if full else qb_half, None, pair_major=True)
        if case == "contract":
            vq = ops.project_phonon_series(vc, eig, pair_major=True) if full else \
                ops.project_phonon_series_mirror(vc, eig2, qmap_d, len(mm))
elif case == "dense":
    # arbitrary (non-commensurate) wave vectors: the DMMA GEMM form of projection.py:24-31
    T, N, Q = int(os.environ.get("PC_T", 2048)), 10240, int(os.environ.get("PC_Q", 512))
    rng = np.random.default_rng(1)
    pos = torch.as_tensor(rng.random((N, 3)) * 64.0, device=dev)
    typ = np.repeat(np.arange(2), N // 2).astype(np.int32)
    v = torch.randn((T, N, 3), dtype=torch.float64, device=dev)
    qc = rng.normal(size=(Q, 3))
    for _ in range(reps):
        vc = ops.project_wave_vector(v, None, pos, typ, 2, qc)
elif case == "mem":
    T, S = (int(os.environ.get("PC_T", 2500)), int(os.environ.get("PC_S", 592)))
    x = torch.randn((T, S), dtype=torch.complex128, device=dev)
    for _ in range(reps):
        out = ops.power_mem(x, 0.002, freq, 1000)
elif case == "corr":
    T, S = (int(os.environ.get("PC_T", 32768)), int(os.environ.get("PC_S", 64)))
    x = torch.randn((T, S), dtype=torch.complex128, device=dev)
    for _ in range(reps):
        out = ops.power_correlation(x, 0.001, freq, 10, 1, 1)
torch.cuda.synchronize()
print("done", case)
