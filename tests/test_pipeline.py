"""Whole-mesh pipeline (projection -> contraction -> [all-to-all] -> spectra) against the oracle,
including the time-sharded multi-rank path: 2 gloo ranks on CPU drive the emulated kernels and the
real torch.distributed all-to-all; on the GPU box the single-rank CUDA path is checked too."""
import os
import sys

import numpy as np
import pytest

from conftest import relerr, to_np
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_case(nt=64, dims=(4, 2, 3), seed=5):
    from dynaphopy_b200 import pipeline
    rng = np.random.default_rng(seed)
    cell = np.diag([4.0, 4.1, 3.9])
    dims = np.array(dims)
    nc = int(np.prod(dims))
    unit_pos = np.array([[0, 0, 0], [0.5, 0.5, 0.5]]) @ cell
    pos = port.supercell_positions(unit_pos, cell, dims)
    typ = port.supercell_replicate([0, 1], dims).astype(np.int32)
    mass = port.supercell_replicate([28.0855, 69.723], dims)
    m = np.array([[x, y, z] for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])])
    q_cart = 2 * np.pi * (m / dims) @ np.linalg.inv(cell).T
    t = np.arange(nt) * 0.002
    v = rng.normal(size=(nt, 2 * nc, 3)) + 2.0 * np.cos(2 * np.pi * 7.3 * t)[:, None, None]
    eig = np.stack([np.linalg.qr(rng.normal(size=(6, 6)) + 1j * rng.normal(size=(6, 6)))[0].T for _ in m])
    freq = np.arange(0, 40.05, 2.0)
    plan = pipeline.CommensuratePlan(cell, dims, pos, mass, typ, 2, q_cart)
    vm = port.velocity_mass_average(v, mass)
    ref_vq = np.stack([port.project_onto_phonon(port.project_onto_wave_vector(vm, pos, typ, 2, q), eig[i].reshape(6, 2, 3))
                       for i, q in enumerate(q_cart)], axis=1)                        # (T, Q, M)
    ref_ps = port.fft_spectra(ref_vq.reshape(nt, -1), 0.002, freq)
    return dict(plan=plan, v=v, eig=eig, freq=freq, ref_vq=ref_vq, ref_ps=ref_ps, dt=0.002)


def test_plan_classifies_q():
    c = make_case()
    assert c["plan"].usable() and c["plan"].commensurate.all()
    from dynaphopy_b200 import pipeline
    p = c["plan"]
    q2 = np.vstack([p.q_cart[:3], p.q_cart[3] * 1.0001])
    cell = np.diag([4.0, 4.1, 3.9])
    pos = port.supercell_positions(np.array([[0, 0, 0], [0.5, 0.5, 0.5]]) @ cell, cell, p.dims)
    plan2 = pipeline.CommensuratePlan(cell, p.dims, pos, np.ones(len(pos)), np.repeat([0, 1], p.nc), 2, q2)
    assert list(plan2.commensurate) == [True, True, True, False]
    # shuffled sites are not a lattice in reference order -> DFT form refused
    plan3 = pipeline.CommensuratePlan(cell, p.dims, pos[::-1], np.ones(len(pos)), np.repeat([0, 1], p.nc), 2, q2)
    assert not plan3.usable()


def test_mesh_pipeline_single_rank(kops):
    from dynaphopy_b200 import pipeline
    c = make_case()
    pipe = pipeline.MeshPipeline(c["plan"], c["eig"], c["freq"], c["dt"], ops=kops, algorithm=2)
    vc = pipe.project(c["v"])
    vq = pipe.contract(vc)
    assert relerr(to_np(vq).reshape(c["ref_vq"].shape), c["ref_vq"]) < 1e-10
    ser = to_np(pipe.series(c["v"], chunk_frames=24))                    # (1, Q*M, T) series-major
    assert relerr(ser[0].T.reshape(c["ref_vq"].shape), c["ref_vq"]) < 1e-10
    ps = to_np(pipe.run(c["v"], chunk_frames=24))
    assert relerr(ps, c["ref_ps"]) < 1e-10
    assert np.array_equal(ps.argmax(0), c["ref_ps"].argmax(0))
    ps_c = to_np(pipe.run_cyclic(c["v"], chunk_frames=16))                # one rank: chunk-cyclic == contiguous
    assert relerr(ps_c, c["ref_ps"]) < 1e-10
    for alg, ref in ((1, port.mem_spectra(c["ref_vq"].reshape(64, -1)[:, :6], c["dt"], c["freq"], 10)),):
        pipe1 = pipeline.MeshPipeline(c["plan"], c["eig"], c["freq"], c["dt"], ops=kops, algorithm=alg, mem_coefficients=10)
        assert relerr(to_np(pipe1.spectra(to_np(vq)[:, :6])), ref) < 1e-10


def test_project_q_list_mixed(kops):
    """Commensurate and non-commensurate q in one list: DFT kernel + dense kernel, stitched."""
    from dynaphopy_b200 import pipeline
    c = make_case(nt=9, dims=(4, 4, 4))
    p = c["plan"]
    cell = np.diag([4.0, 4.1, 3.9])
    pos = port.supercell_positions(np.array([[0, 0, 0], [0.5, 0.5, 0.5]]) @ cell, cell, p.dims)
    typ = np.repeat([0, 1], p.nc).astype(np.int32)
    mass = np.repeat([28.0855, 69.723], p.nc)
    q = np.vstack([p.q_cart[:10], [[0.1, 0.2, 0.3]], p.q_cart[10:12]])
    plan = pipeline.CommensuratePlan(cell, p.dims, pos, mass, typ, 2, q)
    assert plan.commensurate.sum() == 12
    v = c["v"]
    vc = to_np(pipeline.project_q_list(kops, v, np.sqrt(mass), plan, pos, typ))
    vm = port.velocity_mass_average(v, mass)
    ref = np.stack([port.project_onto_wave_vector(vm, pos, typ, 2, qq) for qq in q], axis=1)
    assert relerr(vc, ref) < 1e-10
    for method in ("dense", "fft"):
        plan_c = pipeline.CommensuratePlan(cell, p.dims, pos, mass, typ, 2, p.q_cart[:12])
        out = to_np(pipeline.project_q_list(kops, v, np.sqrt(mass), plan_c, pos, typ, method=method))
        assert relerr(out, ref[:, [*range(10), 11, 12]]) < 1e-10


def _rank_main(rank, world, port_no, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["DP_EMUL_THREADS"] = "2"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port_no}", rank=rank, world_size=world)
    from emul import emul_ops
    from dynaphopy_b200 import pipeline
    c = make_case()
    pipe = pipeline.MeshPipeline(c["plan"], c["eig"], c["freq"], c["dt"], ops=emul_ops(), algorithm=2)
    assert pipe.world == world and pipe.q_block == 24 // world
    nt = c["v"].shape[0]
    tl = nt // world
    v_local = c["v"][rank * tl:(rank + 1) * tl]
    send = pipe.contract(pipe.project(v_local))                       # numpy (G, Tl, Q/G, M)  time-major path
    series = pipe.exchange(torch.from_numpy(np.ascontiguousarray(send))).numpy()
    ps_tm = pipe.spectra(series)
    send2 = pipe.series(v_local, chunk_frames=5)                      # numpy (G, S_loc, Tl)  series-major path
    blocks = pipe.exchange_series(torch.from_numpy(np.ascontiguousarray(send2))).numpy()
    ps = pipe.spectra_series(blocks)
    # (the two paths may pick different projection kernels: equal up to rounding, not bitwise)
    assert relerr(ps, ps_tm) < 1e-12
    assert relerr(blocks.transpose(0, 2, 1).reshape(series.shape), series) < 1e-12
    # overlapped path: chunked asynchronous all-to-all, (G*n_chunks, S_loc, Tc) time blocks
    blocks3 = pipe.series_exchanged(np.ascontiguousarray(v_local), chunk_frames=tl // 2)
    assert tuple(blocks3.shape) == (world * 2, pipe.q_block * pipe.M, tl // 2)
    ps3 = pipe.spectra_series(blocks3)
    assert relerr(ps3, ps) < 1e-12
    np.save(os.path.join(out_dir, f"ps{rank}.npy"), ps)
    np.save(os.path.join(out_dir, f"series{rank}.npy"), series)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_time_sharded_all_to_all_gloo(tmp_path, world):
    """N ranks, time-sharded projection, one all-to-all, series-sharded spectra == oracle."""
    import torch.multiprocessing as mp
    port_no = 29600 + os.getpid() % 300 + world
    mp.spawn(_rank_main, args=(world, port_no, str(tmp_path)), nprocs=world, join=True)
    c = make_case()
    nt = c["ref_vq"].shape[0]
    qb = 24 // world
    for r in range(world):
        series = np.load(tmp_path / f"series{r}.npy")
        ref_series = c["ref_vq"][:, r * qb:(r + 1) * qb].reshape(nt, -1)
        assert relerr(series, ref_series) < 1e-10
        ps = np.load(tmp_path / f"ps{r}.npy")
        assert relerr(ps, c["ref_ps"][:, r * qb * 6:(r + 1) * qb * 6]) < 1e-10


@pytest.mark.parametrize("dims,world", [((4, 2, 3), 2), ((4, 2, 3), 3), ((4, 2, 3), 4), ((4, 2, 3), 8), ((3, 3, 3), 3),
                                        ((3, 3, 3), 9), ((5, 4, 3), 4), ((2, 2, 2), 2), ((6, 5, 4), 8)])
def test_intermediate_exchange_plan_deals_pairs_evenly(dims, world):
    """MeshPipeline.intermediate_exchange_plan (host logic of the multi-GPU intermediate exchange): every rank gets
    exactly Q/G wave vectors, both members of a pair (q, -q) on the same rank, every q exactly once, the local tables
    consistent with the global pairing -- or None when the pairs cannot be dealt evenly (the series exchange is used)."""
    from emul import emul_ops
    from dynaphopy_b200 import pipeline
    dims = np.array(dims)
    nc = int(np.prod(dims))
    cell = np.diag([4.0, 4.1, 3.9])
    unit = np.array([[0, 0, 0], [0.5, 0.5, 0.5]]) @ cell
    grid = np.array([[x, y, z] for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])], float) @ cell
    pos = (unit[:, None, :] + grid[None]).reshape(-1, 3)
    typ = np.repeat(np.arange(2), nc).astype(np.int32)
    mass = np.repeat([28.0855, 69.723], nc)
    m = np.array([[x, y, z] for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])])
    q_cart = (m / dims) @ (2.0 * np.pi * np.linalg.inv(cell).T)
    rng = np.random.default_rng(int(dims.sum()) + world)
    eig = rng.normal(size=(nc, 6, 6)) + 1j * rng.normal(size=(nc, 6, 6))
    plan = pipeline.CommensuratePlan(cell, dims, pos, mass, typ, 2, q_cart)
    pipe = pipeline.MeshPipeline(plan, eig, np.arange(0, 10.0, 0.5), 0.002, ops=emul_ops(), algorithm=2)
    if nc % world:
        return
    qmap = pipeline.mirror_pairs(plan.qbin, plan.dims)
    partner = {int(a): int(b) for a, b in qmap if b >= 0}
    partner.update({b: a for a, b in list(partner.items())})
    n_self = int((qmap[:, 1] < 0).sum())
    seen = []
    for rank in range(world):
        pipe.world, pipe.rank, pipe.q_block, pipe._vcx = world, rank, nc // world, False
        vcx = pipe.intermediate_exchange_plan()
        qb = nc // world
        feasible = pipe.mirror and world <= 16 and (n_self - world * (qb % 2)) >= 0 and (n_self - world * (qb % 2)) % 2 == 0
        if not feasible:
            assert vcx is None
            return
        assert vcx is not None
        lq = vcx["local_q"][rank]
        assert len(lq) == qb and len(set(lq.tolist())) == qb
        for q in lq.tolist():                                  # both members of a pair on this rank
            assert q not in partner or partner[q] in lq.tolist()
        qm = np.asarray(vcx["d_qmap_loc"])
        e2 = np.asarray(vcx["d_eig2_loc"])
        ef = pipe._ef_host
        bins = np.asarray(vcx["d_qbin_all"]).reshape(world, vcx["hb"], 3)[rank]
        for k in range(vcx["hb"]):
            a, b = int(qm[k, 0]), int(qm[k, 1])
            if a < 0:
                assert b < 0                                   # padding row
                continue
            assert np.array_equal(bins[k], np.asarray(plan.qbin)[lq[a]])
            assert np.array_equal(e2[k, 0], ef[lq[a]])
            if b >= 0:
                assert partner[int(lq[a])] == int(lq[b]) and np.array_equal(e2[k, 1], np.conj(ef[lq[b]]))
            else:
                assert int(lq[a]) not in partner
        seen += lq.tolist()
    assert sorted(seen) == list(range(nc))


def _rank_cyclic(rank, world, port_no, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["DP_EMUL_THREADS"] = "2"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port_no}", rank=rank, world_size=world)
    from emul import emul_ops
    from dynaphopy_b200 import pipeline
    c = make_case(nt=768, seed=7)
    pipe = pipeline.MeshPipeline(c["plan"], c["eig"], c["freq"], c["dt"], ops=emul_ops(), algorithm=2)
    chunks = pipe.cyclic_chunks(768, 96)
    assert chunks[0] == (rank * 96, rank * 96 + 96) and len(chunks) == 768 // (world * 96)
    v_local = np.concatenate([c["v"][a:b] for a, b in chunks])
    ps = pipe.run_cyclic(v_local, chunk_frames=96)
    assert pipe.intermediate_exchange_plan() is not None and sorted(np.concatenate(pipe._vcx["local_q"])) == list(range(24))
    np.save(os.path.join(out_dir, f"psc{rank}.npy"), np.asarray(ps))
    # the series-exchange form of the same step (DP_NO_VC_EXCHANGE) gives the same spectra
    os.environ["DP_NO_VC_EXCHANGE"] = "1"
    pipe2 = pipeline.MeshPipeline(c["plan"], c["eig"], c["freq"], c["dt"], ops=emul_ops(), algorithm=2)
    ps2 = pipe2.run_cyclic(v_local, chunk_frames=96)
    assert pipe2.intermediate_exchange_plan() is None
    assert np.abs(np.asarray(ps2) - np.asarray(ps)).max() <= 1e-13 * np.abs(np.asarray(ps)).max()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_chunk_cyclic_sharding_gloo(tmp_path, world):
    """Chunk-cyclic time sharding (round c completes frames [0, (c+1)*G*chunk) on every rank; the spectra of finished
    windows are launched round by round): 2 ranks x 4 rounds / 4 ranks x 2 rounds, 4 overlapping windows of 250 samples
    == oracle.  The ranks exchange the half-mesh INTERMEDIATE (MeshPipeline.intermediate_exchange_plan: every pair
    (q, -q) lives on one rank, which contracts both members) and restore the natural q order with one small all-to-all."""
    import torch.multiprocessing as mp
    port_no = 29900 + os.getpid() % 90 + world
    mp.spawn(_rank_cyclic, args=(world, port_no, str(tmp_path)), nprocs=world, join=True)
    c = make_case(nt=768, seed=7)
    assert len(set(map(tuple, port.division_of_data(c["freq"][1] - c["freq"][0], 768, c["dt"])))) == 4
    qb = 24 // world
    for r in range(world):
        ps = np.load(tmp_path / f"psc{r}.npy")
        ref = c["ref_ps"][:, r * qb * 6:(r + 1) * qb * 6]
        assert relerr(ps, ref) < 1e-10
        assert np.array_equal(ps.argmax(0), ref.argmax(0))


@pytest.mark.gpu
def test_run_from_host_streams(cuda_ops):
    """Pinned-host trajectory, chunked H2D overlapped with projection == device-resident run."""
    import torch
    from dynaphopy_b200 import pipeline
    c = make_case(nt=96)
    pipe = pipeline.MeshPipeline(c["plan"], c["eig"], c["freq"], c["dt"], ops=cuda_ops, algorithm=2)
    vh = torch.from_numpy(c["v"]).pin_memory()
    ps = pipe.run_from_host(vh, chunk_frames=20)
    ref = port.fft_spectra(np.stack([port.project_onto_phonon(
        port.project_onto_wave_vector(port.velocity_mass_average(c["v"], np.repeat([28.0855, 69.723], 24)),
                                      port.supercell_positions(np.array([[0, 0, 0], [0.5, 0.5, 0.5]]) @ np.diag([4.0, 4.1, 3.9]),
                                                               np.diag([4.0, 4.1, 3.9]), c["plan"].dims),
                                      np.repeat([0, 1], 24), 2, q), c["eig"][i].reshape(6, 2, 3))
        for i, q in enumerate(c["plan"].q_cart)], axis=1).reshape(96, -1), c["dt"], c["freq"])
    assert relerr(ps, ref) < 1e-10
    assert relerr(ps, to_np(pipe.run(c["v"]))) < 1e-13


@pytest.mark.gpu
def test_full_mesh_parseval_unitarity_linearity(cuda_ops):
    """BASELINE cfg 5 geometry at full width (10 240 atoms, all 5120 commensurate q, 6 modes), where the oracle is out
    of reach: size-independent properties of the path.  (1) Parseval of the lattice DFT over the full mesh:
    sum_q |vc[t,q,p,k]|^2 = sum_cells m_p v[t,cell,p,k]^2.  (2) The contraction with unitary eigenvectors preserves
    the norm per (t, q).  (3) The whole projection + contraction is linear in the velocities.  (4) The series-major
    pipeline output is the transpose of the time-major one."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from dynaphopy_b200 import pipeline
    dev = cuda_ops.be.device
    wl = bench.workload(1024)
    plan = pipeline.CommensuratePlan(wl["cell"], bench.DIMS, wl["pos"], wl["mass"], wl["typ"], wl["n_prim"], wl["q_cart"])
    assert plan.usable() and plan.commensurate.all()
    eig = bench.random_unitary_eigenvectors(torch, dev, wl["Q"], wl["M"])
    pipe = pipeline.MeshPipeline(plan, eig, bench.FREQ, bench.DT, ops=cuda_ops, algorithm=2)
    nt = 256
    v1 = bench.gen_chunk_torch(torch, dev, wl, 0, nt)
    v2 = torch.randn(v1.shape, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    vc = pipe.project(v1)                                                   # (T, Q, n_p, 3), twiddles and sqrt(m) applied
    lhs = (vc.real ** 2 + vc.imag ** 2).sum(dim=1)                          # (T, n_p, 3)
    mass = torch.as_tensor(wl["mass"], device=dev).reshape(1, 2, -1, 1)
    rhs = (mass * v1.reshape(nt, 2, -1, 3) ** 2).sum(dim=2)
    assert float(((lhs - rhs).abs().max() / rhs.abs().max()).item()) < 1e-12
    n_vc = (vc.real ** 2 + vc.imag ** 2).sum(dim=(2, 3))                    # (T, Q)
    vq = pipe.contract(vc.clone())                                          # (T, Q*M) time-major, in place on the copy
    vq = vq.reshape(nt, wl["Q"], wl["M"])
    n_vq = (vq.real ** 2 + vq.imag ** 2).sum(dim=2)
    assert float(((n_vq - n_vc).abs().max() / n_vc.abs().max()).item()) < 1e-12
    s1 = pipe.series(v1, chunk_frames=128)                                  # (1, Q*M, T) series-major, folded twiddle path
    assert float(((s1[0].t().reshape(nt, wl["Q"], wl["M"]) - vq).abs().max() / vq.abs().max()).item()) < 1e-11
    a, b = 0.75, -1.5
    s2 = pipe.series(v2, chunk_frames=128)
    s12 = pipe.series(a * v1 + b * v2, chunk_frames=128)
    assert float(((s12 - (a * s1 + b * s2)).abs().max() / s12.abs().max()).item()) < 1e-12
