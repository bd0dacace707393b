"""Small shapes of every hot-path kernel for `compute-sanitizer` (memcheck / racecheck / synccheck); each case also
checks its result against a torch fp64 restatement so a sanitizer run is a correctness run too.

    compute-sanitizer --tool racecheck python tools/sanitize_case.py fft|fft_big|project|dense|mem_cluster|corr
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaphopy_b200 import ops as dops  # noqa: E402

case = sys.argv[1]
ops = dops.Ops()
dev = ops.be.device
torch.manual_seed(1)

if case == "fft":
    # spectrum2_kernel: L = 2000 (three paired windows + an odd one), tensor-memory parking, bulk stores
    x = torch.randn((3, 7001), dtype=torch.complex128, device=dev)
    out = ops.power_fft(x, 0.001, np.arange(0, 40.0, 0.5), series_major=True)
    assert bool(torch.isfinite(out).all()), "non-finite spectrum"
elif case == "fft_big":
    # windows with P > 65536 (three-factor engine)
    nt = 70000
    x = torch.randn((2, nt), dtype=torch.complex128, device=dev)
    out = ops.power_fft(x, 0.001, np.arange(600) * (1.0 / (nt * 0.001)), series_major=True)
    assert bool(torch.isfinite(out).all()), "non-finite spectrum"
elif case == "project":
    # project_fft_kernel + project_phonon_series_kernel on a 4 x 4 x 5 supercell
    dims = np.array([4, 4, 5]); nc = int(np.prod(dims)); T = 24
    mm = np.array([[x, y, z] for z in range(5) for y in range(4) for x in range(4)], dtype=np.int32)
    v = torch.randn((T, 2 * nc, 3), dtype=torch.float64, device=dev)
    eig = torch.randn((len(mm), 6, 6), dtype=torch.complex128, device=dev)
    vc = ops.project_commensurate(v, dims, [0, 1], 2, torch.as_tensor(mm, device=dev), None, pair_major=True)
    vq = ops.project_phonon_series(vc, eig, pair_major=True)
    vc2 = ops.project_commensurate(v, dims, [0, 1], 2, torch.as_tensor(mm, device=dev), None)
    vq2 = ops.project_phonon(vc2, eig)                               # time-major path, (T, Q, M)
    ref = vq2.reshape(T, -1).transpose(0, 1)                          # (Q*M, T)
    err = float((vq.reshape(-1, T) - ref).abs().max() / ref.abs().max())
    assert err < 1e-12, err
elif case == "dense":
    # project_dense_kernel (DMMA) with an incommensurate q batch
    T, N, Q = 40, 48, 9
    v = torch.randn((T, N, 3), dtype=torch.float64, device=dev)
    pos = np.random.default_rng(0).normal(size=(N, 3))
    typ = np.arange(N) % 2
    q = np.random.default_rng(1).normal(size=(Q, 3))
    out = ops.project_wave_vector(v, None, pos, typ, 2, q)
    ph = torch.as_tensor(np.exp(-1j * pos @ q.T), device=dev)                   # (N, Q)
    ref = torch.zeros((T, Q, 2, 3), dtype=torch.complex128, device=dev)
    for p_ in range(2):
        sel = torch.as_tensor(np.nonzero(typ == p_)[0], device=dev)
        ref[:, :, p_, :] = torch.einsum("tnk,nq->tqk", v[:, sel, :].to(torch.complex128), ph[sel, :])
    ref = ref * np.sqrt(2.0 / N)        # (N / n_prim)^-1/2
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < 1e-12, err
elif case == "mem_cluster":
    # burg_cluster_kernel: the series no longer fits one CTA's shared memory
    x = torch.randn((16384, 2), dtype=torch.complex128, device=dev)
    out = ops.power_mem(x, 0.001, np.arange(0, 40.0, 2.0), 12)
    assert bool(torch.isfinite(out).all())
elif case == "corr":
    x = torch.randn((2048, 3), dtype=torch.complex128, device=dev)
    out = ops.power_correlation(x, 0.001, np.arange(0, 40.0, 2.0), 10, 1, 1)
    assert bool(torch.isfinite(out).all())
else:
    raise SystemExit("unknown case " + case)
torch.cuda.synchronize()
print("done", case)
