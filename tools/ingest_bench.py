"""f4 end to end at the size of cfg 3 (GaN 3x3x3: 108 atoms, 4 atoms per primitive cell, 50 001 frames, 27 commensurate
q x 12 modes, 626 frequencies): LAMMPS velocity dump (text) -> native reader into a pinned buffer ->
MeshPipeline.run_from_host -> spectra, with the time of every leg.

    python tools/ingest_bench.py [frames]
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_dump(path, frames, n_atoms, box, rng, block=512):
    """Velocity dump as LAMMPS writes it (ITEM: ATOMS vx vy vz), white noise + one spectral line."""
    t0 = time.perf_counter()
    header = ("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n" + str(n_atoms) + "\nITEM: BOX BOUNDS pp pp pp\n" +
              "\n".join("0.0 %.6f" % b for b in box) + "\nITEM: ATOMS vx vy vz\n")
    phase = rng.uniform(0, 2 * np.pi, size=(n_atoms, 3))
    with open(path, "w") as f:
        for s0 in range(0, frames, block):
            n = min(block, frames - s0)
            t = (np.arange(s0, s0 + n) * 0.001)[:, None, None]
            v = rng.normal(scale=2.0, size=(n, n_atoms, 3)) + 1.5 * np.cos(2 * np.pi * 9.0 * t + phase[None])
            out = []
            for k in range(n):
                out.append(header % (s0 + k))
                out.append("\n".join("%.8f %.8f %.8f" % (a, b, c) for a, b, c in v[k]))
                out.append("\n")
            f.write("".join(out))
    return time.perf_counter() - t0


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 50001
    import torch
    from dynaphopy_b200 import iofile, ops as dops, pipeline
    from dynaphopy_b200.structure import Structure
    rng = np.random.default_rng(3)
    dims = np.array([3, 3, 3])
    cell = np.diag([3.19, 5.52, 5.19])                       # orthorhombic setting of a 4-atom cell
    frac = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.33, 0.5], [0.5, 0.83, 0.5]])
    masses = np.array([69.723, 69.723, 14.007, 14.007])
    st = Structure(cell=cell, positions=frac @ cell, masses=masses)
    nc = int(np.prod(dims))
    n_atoms = 4 * nc
    path = os.path.join(tempfile.gettempdir(), "dp_cfg3_vel.lammpstrj")
    t_write = write_dump(path, frames, n_atoms, np.diag(np.diag(dims) @ cell), rng)
    size = os.path.getsize(path)
    torch.zeros(1, device="cuda")                            # CUDA context and the pinned allocator exist before the clock starts
    torch.empty(1 << 20, dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    dyn = iofile.read_lammps_trajectory(path, st, time_step=0.001, pinned=True)
    t_read = time.perf_counter() - t0
    vel = np.asarray(dyn._velocity)                          # the reader's pinned host buffer itself (no device round trip)
    vel = np.ascontiguousarray(vel.real if np.iscomplexobj(vel) else vel)
    assert vel.shape == (frames, n_atoms, 3), vel.shape
    # atoms in a LAMMPS dump of a replicated cell: unit atom fastest is NOT assumed -- positions / types as the pipeline wants
    grid = np.array([[x, y, z] for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])], float) @ cell
    pos = ((frac @ cell)[:, None, :] + grid[None]).reshape(-1, 3)
    typ = np.repeat(np.arange(4), nc).astype(np.int32)
    mass = np.repeat(masses, nc)
    m = np.array([[x, y, z] for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])])
    q_cart = (m / dims) @ (2.0 * np.pi * np.linalg.inv(cell).T)
    plan = pipeline.CommensuratePlan(cell, dims, pos, mass, typ, 4, q_cart)
    assert plan.usable() and plan.commensurate.all()
    eig = np.stack([np.linalg.qr(rng.normal(size=(12, 12)) + 1j * rng.normal(size=(12, 12)))[0].T for _ in m])
    freq = np.arange(0, 31.3, 0.05)
    ops = dops.Ops()
    pipe = pipeline.MeshPipeline(plan, eig, freq, 0.001, ops=ops, algorithm=2)
    vh = torch.from_numpy(vel)
    if not vh.is_pinned():
        vh = vh.pin_memory()
    pipe.run_from_host(vh, chunk_frames=2048)                 # warm-up (allocations, first launches)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ps = pipe.run_from_host(vh, chunk_frames=2048)
    torch.cuda.synchronize()
    t_run = time.perf_counter() - t0
    peak = float(freq[int(np.argmax(np.asarray(ps).sum(axis=1)))])
    print(json.dumps({"case": "cfg 3 geometry: %d atoms, %d frames, 27 q x 12 modes, %d frequencies" % (n_atoms, frames, freq.size),
                      "dump_bytes": size, "write_dump_s": t_write, "read_lammps_trajectory_s": t_read,
                      "read_MBps_text": size / t_read / 1e6, "host_threads": os.cpu_count(),
                      "run_from_host_s": t_run, "velocity_bytes": int(vel.nbytes),
                      "text_to_spectra_s": t_read + t_run, "text_to_spectra_MBps": size / (t_read + t_run) / 1e6,
                      "peak_THz_of_summed_spectra": peak, "spectra_shape": list(np.asarray(ps).shape)}))
    os.remove(path)


if __name__ == "__main__":
    main()
