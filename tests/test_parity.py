"""Parity of the CUDA path with the reference, through the C ABI.

Every test runs twice: on the host emulation of the kernel sources (CPU suite, `-m "not gpu"`)
and on the real sm_100a library (`-m gpu`).  Expected values come from
  * tests/golden/*.npz -- outputs of the REFERENCE ITSELF (oracle/gen_golden.py), and
  * oracle.port        -- the CPU restatement pinned against the reference in
                          tests/test_oracle_pinning.py,
never from the code under test.  Tolerances: bit-exact for the displacement / velocity maps,
1e-10 normwise relative (BASELINE.json north_star) for projections and spectra, exact argmax
(peak-frequency bin) for spectra.
"""
import numpy as np
import pytest

from conftest import golden, relerr, to_np
from oracle import port

TOL = 1e-10


def is_cuda(kops):
    return type(kops.be).__name__ == "TorchCudaBackend"


# ------------------------------------------------------------------------------- A8 / A9 / A1
def test_displacements_bit_exact(kops):
    g = golden("displacements_triclinic")
    u = to_np(kops.atomic_displacements(g["trajectory"], g["positions"], g["supercell"]))
    assert np.array_equal(u, g["relative"])
    # complex128 input (the reference's native dtype): only the real part is used
    uc = to_np(kops.atomic_displacements(g["trajectory"].astype(complex), g["positions"], g["supercell"]))
    assert np.array_equal(uc, g["relative"])


def test_velocity_pipeline_bit_exact(kops):
    g = golden("displacements_triclinic")
    dt = float(g["dt"])
    u, vm = kops.velocity_from_positions(g["trajectory"], g["positions"], g["supercell"], dt,
                                         sqrt_mass=np.sqrt(g["masses"]), want_displacements=True)
    assert np.array_equal(to_np(u), g["relative"])
    assert np.array_equal(to_np(vm), g["vm"])
    v = to_np(kops.velocity_from_positions(g["trajectory"], g["positions"], g["supercell"], dt))
    assert np.array_equal(v, g["velocity"])
    assert np.array_equal(to_np(kops.mass_weight(g["velocity"], np.sqrt(g["masses"]))), g["vm"])


@pytest.mark.parametrize("cut", [1, 7, 48, 95])
def test_velocity_time_shard_halo(kops, cut):
    """Time-sharded evaluation with a one-frame halo equals the single-block result."""
    g = golden("displacements_triclinic")
    dt = float(g["dt"])
    tr = g["trajectory"]
    a = to_np(kops.velocity_from_positions(tr[:cut], g["positions"], g["supercell"], dt, next_frame=tr[cut]))
    b = to_np(kops.velocity_from_positions(tr[cut:], g["positions"], g["supercell"], dt, prev_frame=tr[cut - 1]))
    assert np.array_equal(np.concatenate([a, b]), g["velocity"])


@pytest.mark.parametrize("tag", ["xdatcar", "lammps"])
def test_reading_test_known_answers(kops, tag):
    """unittest/reading_test.py:23-102 -- the reference's only known-answer test on this path."""
    g = golden("reading_test")
    u = to_np(kops.atomic_displacements(g[tag + "_trajectory"], g[tag + "_positions"], g[tag + "_supercell"]))
    assert np.array_equal(u, g[tag + "_relative"])
    # Dynamics.get_mean_displacement_matrix (dynamics.py:207-239) on top of the kernel output
    typ = g[tag + "_atom_type"]
    ntypes = typ.max() + 1
    mats = np.zeros((ntypes, 3, 3))
    counts = np.bincount(typ)
    for i in range(u.shape[1]):
        mats[typ[i]] += u[:, i, :].T @ u[:, i, :] / counts[typ[i]]
    mean = np.average(mats / u.shape[0], axis=0)
    assert np.allclose(mean, g[tag + "_mean_matrix"], rtol=1e-12, atol=0)
    assert np.allclose(mean, g[tag + "_mean_matrix_literal"], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("tag", ["xdatcar", "lammps"])
def test_displacement_moments(kops, tag):
    """Fused displacement + per-atom sums (feeds dynamics.py:207-295) against the reference's relative trajectory."""
    g = golden("reading_test")
    u = g[tag + "_relative"]
    sums = to_np(kops.displacement_moments(g[tag + "_trajectory"], g[tag + "_positions"], g[tag + "_supercell"]))
    i, j = np.triu_indices(3)
    ref = np.concatenate([u.sum(0), (u[:, :, i] * u[:, :, j]).sum(0)], axis=1)
    assert sums.shape == (u.shape[1], 9)
    assert np.abs(sums - ref).max() <= 1e-13 * np.abs(ref).max()
    # long run, several time splits, complex128 positions (real part used), wrapped atoms
    rng = np.random.default_rng(31)
    cell = np.array([[8.0, 0.3, 0.0], [0.0, 9.0, 0.4], [0.5, 0.0, 7.0]])
    pos = rng.random((37, 3)) @ cell
    tr = pos[None] + rng.normal(scale=0.1, size=(2000, 37, 3)) + rng.integers(-2, 3, size=(2000, 37, 3)) @ cell
    u = np.stack([port.atomic_displacements(tr[:, a, :], pos[a], cell) for a in range(37)], axis=1)
    ref = np.concatenate([u.sum(0), (u[:, :, i] * u[:, :, j]).sum(0)], axis=1)
    for arr in (tr, tr + 1j * rng.normal(size=tr.shape)):
        sums = to_np(kops.displacement_moments(arr, pos, cell))
        assert np.abs(sums - ref).max() <= 1e-12 * np.abs(ref).max()


def test_displacements_empty_and_errors(kops):
    g = golden("displacements_triclinic")
    out = kops.atomic_displacements(np.zeros((0, 24, 3)), g["positions"], g["supercell"])
    assert tuple(out.shape) == (0, 24, 3)
    from dynaphopy_b200._lib import DpError
    with pytest.raises(DpError):      # np.gradient needs two frames
        kops.velocity_from_positions(g["trajectory"][:1], g["positions"], g["supercell"], 0.002)


# ------------------------------------------------------------------------------- A2 / A3
def _commensurate_inputs(g, q_indices=None):
    dims = g["dims"]
    nc = int(np.prod(dims))
    cell = g["unit_cell"]
    n_prim = int(g["n_prim"])
    unit_type = g["atom_type"][::nc]
    tau = g["positions"][::nc]
    keep, qb, tw = [], [], []
    for i in (range(len(g["q_cart"])) if q_indices is None else q_indices):
        m = g["q_cart"][i] @ cell.T / (2 * np.pi) * dims
        if np.abs(m - np.round(m)).max() > 1e-6:
            continue
        keep.append(i)
        qb.append(np.mod(np.round(m).astype(int), dims))
        tw.append(np.sqrt(g["unit_masses"]) * np.exp(-1j * (tau @ g["q_cart"][i])) / np.sqrt(len(g["masses"]) / n_prim))
    return dims, unit_type, n_prim, np.array(qb, dtype=np.int32), np.array(tw), keep


@pytest.mark.parametrize("name", ["cubic2_232", "si_conv_222", "cubic2_222_complex"])
def test_projection_dense_vs_reference(kops, name):
    g = golden(name)
    n_prim = int(g["n_prim"])
    sm = np.sqrt(g["masses"])
    vc = to_np(kops.project_wave_vector(g["velocity"], sm, g["positions"], g["atom_type"], n_prim, g["q_cart"]))
    assert vc.shape == (g["velocity"].shape[0], len(g["q_cart"]), n_prim, 3)
    assert relerr(vc.transpose(1, 0, 2, 3), g["vc"]) < TOL
    # already mass-weighted input, one q at a time (the reference's calling pattern)
    vc1 = to_np(kops.project_wave_vector(g["vm"], None, g["positions"], g["atom_type"], n_prim, g["q_cart"][2]))
    assert relerr(vc1[:, 0], g["vc"][2]) < TOL
    vca = to_np(kops.project_wave_vector(g["velocity"], sm, g["positions"], g["atom_type"], n_prim, g["q_cart"][1],
                                         project_on_atom=0))
    assert relerr(vca[:, 0], g["vc_atom0"]) < TOL
    assert np.all(vca[:, 0, 1:] == 0)
    vq = to_np(kops.project_phonon(vc, g["eigenvectors"]))
    assert relerr(vq.transpose(1, 0, 2), g["vq"]) < TOL


@pytest.mark.parametrize("name", ["cubic2_232", "si_conv_222"])
def test_projection_commensurate_vs_reference(kops, name):
    g = golden(name)
    dims, unit_type, n_prim, qb, tw, keep = _commensurate_inputs(g)
    assert len(keep) >= 4
    vc = to_np(kops.project_commensurate(g["velocity"], dims, unit_type, n_prim, qb, tw))
    assert relerr(vc.transpose(1, 0, 2, 3), g["vc"][keep]) < TOL
    dims, unit_type, n_prim, qb, tw, keep = _commensurate_inputs(g, [1])
    vca = to_np(kops.project_commensurate(g["velocity"], dims, unit_type, n_prim, qb, tw, project_on_atom=0))
    assert relerr(vca[:, 0], g["vc_atom0"]) < TOL
    assert np.all(vca[:, 0, 1:] == 0)


def _random_lattice_case(rng, dims, n_unit, n_prim, nt):
    """Seeded synthetic structure (SURVEY 8d style) + oracle projection for every commensurate q."""
    dims = np.array(dims)
    nc = int(np.prod(dims))
    cell = np.array([[4.0, 0.0, 0.0], [0.3, 4.2, 0.0], [-0.2, 0.5, 3.9]])
    frac = rng.random((n_unit, 3))
    unit_pos = frac @ cell
    unit_type = np.arange(n_unit) % n_prim
    unit_mass = 20.0 + 10.0 * unit_type
    pos = port.supercell_positions(unit_pos, cell, dims)
    typ = port.supercell_replicate(unit_type, dims).astype(np.int32)
    mass = port.supercell_replicate(unit_mass, dims)
    v = rng.normal(size=(nt, n_unit * nc, 3))
    return dict(dims=dims, cell=cell, unit_pos=unit_pos, unit_type=unit_type.astype(np.int32), unit_mass=unit_mass,
                pos=pos, typ=typ, mass=mass, v=v, n_prim=n_prim, nc=nc)


@pytest.mark.parametrize("dims,n_unit,n_prim", [((1, 1, 1), 2, 2), ((2, 1, 3), 1, 1), ((3, 5, 2), 3, 3),
                                                ((7, 2, 2), 4, 2), ((4, 6, 5), 2, 2), ((11, 1, 2), 2, 1),
                                                ((16, 16, 20), 2, 2), ((13, 9, 8), 3, 3), ((32, 2, 1), 2, 2),
                                                ((17, 3, 3), 1, 1), ((12, 10, 6), 5, 2)])
def test_projection_commensurate_shapes(kops, dims, n_unit, n_prim):
    """All pencil lengths / unit-cell layouts: FFT form == dense form == oracle on random q subsets."""
    rng = np.random.default_rng(hash((dims, n_unit)) % (2 ** 31))
    nt = 3 if not is_cuda(kops) else 17
    c = _random_lattice_case(rng, dims, n_unit, n_prim, nt)
    nq = min(c["nc"], 24)
    m = np.stack([rng.integers(0, d, nq) for d in c["dims"]], axis=1)
    m[0] = 0
    shift = rng.integers(-1, 2, size=m.shape)                     # q outside the first cell too
    q_unit = (m + shift * c["dims"]) / c["dims"]
    q_cart = 2 * np.pi * q_unit @ np.linalg.inv(c["cell"]).T
    tw = np.sqrt(c["unit_mass"])[None, :] * np.exp(-1j * q_cart @ c["unit_pos"].T) / np.sqrt(c["nc"] * n_unit / n_prim)
    vc = to_np(kops.project_commensurate(c["v"], c["dims"], c["unit_type"], n_prim, m.astype(np.int32), tw))
    vm = port.velocity_mass_average(c["v"], c["mass"])
    ref = np.stack([port.project_onto_wave_vector(vm, c["pos"], c["typ"], n_prim, q) for q in q_cart], axis=1)
    assert relerr(vc, ref) < TOL
    dense = to_np(kops.project_wave_vector(c["v"], np.sqrt(c["mass"]), c["pos"], c["typ"], n_prim, q_cart))
    assert relerr(dense, ref) < TOL


def test_projection_ragged_and_empty(kops):
    g = golden("cubic2_232")
    n_prim = int(g["n_prim"])
    sm = np.sqrt(g["masses"])
    # T and Q not multiples of any tile size; 37 copies of 5 q
    q = np.tile(g["q_cart"], (37, 1))[:181]
    vc = to_np(kops.project_wave_vector(g["velocity"][:77], sm, g["positions"], g["atom_type"], n_prim, q))
    assert relerr(vc[:, :5].transpose(1, 0, 2, 3), g["vc"][:, :77]) < TOL
    assert np.array_equal(vc[:, 5:10], vc[:, :5])
    out = kops.project_wave_vector(g["velocity"][:0], sm, g["positions"], g["atom_type"], n_prim, q)
    assert tuple(out.shape) == (0, 181, n_prim, 3)
    from dynaphopy_b200._lib import DpError
    with pytest.raises(DpError):
        kops.project_wave_vector(g["velocity"], sm, g["positions"], g["atom_type"], n_prim, q, project_on_atom=5)
    dims, unit_type, n_prim, qb, tw, _ = _commensurate_inputs(g)
    with pytest.raises(DpError):      # DFT bin outside the grid, bins on the host: checked before the launch
        kops.project_commensurate(g["velocity"], dims, unit_type, n_prim, qb + 7, tw)
    # bins already on the device: checked by the kernel (no host round trip), the offending q comes back as NaN
    bad = qb.copy()
    bad[1] = [0, 7, 0]
    out = to_np(kops.project_commensurate(g["velocity"], dims, unit_type, n_prim, bad, tw, validate_bins=False))
    good = to_np(kops.project_commensurate(g["velocity"], dims, unit_type, n_prim, qb, tw))
    assert np.isnan(out[:, 1]).all()
    keep = [i for i in range(len(qb)) if i != 1]
    assert np.array_equal(out[:, keep], good[:, keep])


def test_pair_major_intermediate_layout(kops):
    """Bare lattice DFT in the pair-major layout + twiddle folded into the eigenvectors == projection with
    twiddle + contraction in the reference layout (what MeshPipeline relies on)."""
    g = golden("cubic2_232")
    dims, unit_type, n_prim, qb, tw, keep = _commensurate_inputs(g)
    v = g["velocity"]
    nq, m = len(keep), 3 * n_prim
    rng = np.random.default_rng(9)
    eig = rng.normal(size=(nq, m, n_prim, 3)) + 1j * rng.normal(size=(nq, m, n_prim, 3))
    vc = kops.project_commensurate(v, dims, unit_type, n_prim, qb, tw)
    ref = to_np(kops.project_phonon_series(vc, eig))
    raw = kops.project_commensurate(v, dims, unit_type, n_prim, qb, None, pair_major=True)
    assert tuple(raw.shape) == (v.shape[0], m // 2, nq, 2)
    folded = eig * np.conj(np.asarray(tw))[:, None, :, None]
    out = to_np(kops.project_phonon_series(raw, folded, pair_major=True))
    assert relerr(out, ref) < 1e-13
    # and the rows layout of the bare DFT carries the same numbers
    raw_rows = to_np(kops.project_commensurate(v, dims, unit_type, n_prim, qb, None))
    assert np.array_equal(to_np(raw).transpose(0, 2, 1, 3).reshape(raw_rows.shape[0], nq, -1),
                          raw_rows.reshape(raw_rows.shape[0], nq, -1))


@pytest.mark.parametrize("dims,n_prim,nt,drop", [((4, 3, 5), 2, 70, 0), ((2, 2, 2), 2, 33, 0), ((3, 4, 2), 4, 45, 3),
                                                 ((6, 5, 4), 2, 64, 7)])
def test_half_mesh_contraction_is_bit_identical(kops, dims, n_prim, nt, drop):
    """dp_project_phonon_series_mirror: one wave vector of every pair (q, -q) projected and stored, both contracted --
    the same bits as projecting and contracting the whole list (real velocities: vc[-q] = conj(vc[q]); projection.py:24-54).
    Lists with self-conjugate points, with partners missing (`drop` wave vectors removed) and with q blocks."""
    from dynaphopy_b200.pipeline import mirror_pairs
    rng = np.random.default_rng(sum(dims) + nt)
    dims = np.array(dims, dtype=np.int32)
    nc = int(np.prod(dims))
    m = 3 * n_prim
    unit_type = np.arange(n_prim, dtype=np.int32)
    v = rng.normal(size=(nt, n_prim * nc, 3))
    qb = np.array([[x, y, z] for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])], dtype=np.int32)
    if drop:
        qb = np.delete(qb, rng.choice(len(qb), size=drop, replace=False), axis=0)
    nq = len(qb)
    eig = rng.normal(size=(nq, m, m)) + 1j * rng.normal(size=(nq, m, m))
    raw = kops.project_commensurate(v, dims, unit_type, n_prim, qb, None, pair_major=True)
    q_block = nq // 2 if nq % 2 == 0 else nq
    ref = to_np(kops.project_phonon_series(raw, eig, q_block=q_block, pair_major=True))
    qmap = mirror_pairs(qb, dims)
    assert sorted(int(i) for i in qmap.ravel() if i >= 0) == list(range(nq))          # every q exactly once
    assert len(qmap) < nq or tuple(dims) == (2, 2, 2)          # 2 x 2 x 2: every wave vector is its own negative
    eig2 = np.zeros((len(qmap), 2, m, m), dtype=complex)
    eig2[:, 0] = eig[qmap[:, 0]]
    has = qmap[:, 1] >= 0
    eig2[has, 1] = np.conj(eig[qmap[has, 1]])
    half = kops.project_commensurate(v, dims, unit_type, n_prim, qb[qmap[:, 0]], None, pair_major=True)
    out = to_np(kops.project_phonon_series_mirror(half, eig2, qmap, nq, q_block=q_block))
    assert out.shape == ref.shape
    assert np.array_equal(out, ref)
    # the same intermediate written as per-destination blocks (dp_project_commensurate_blocks): block g holds the
    # columns [g*hb, (g+1)*hb) of every (frame, slot pair) row
    nh = len(qmap)
    for hb in [d for d in range(1, nh + 1) if nh % d == 0 and nh // d <= 16]:
        blocks = to_np(kops.project_commensurate_blocks(v, dims, unit_type, n_prim, qb[qmap[:, 0]], hb))
        assert blocks.shape == (nh // hb, nt, m // 2, hb, 2)
        assert np.array_equal(np.concatenate(list(blocks), axis=2), to_np(half))
    # chunked in time, as the pipeline calls it
    out2 = kops.project_phonon_series_mirror(half[:20], eig2, qmap, nq, q_block=q_block, t_total=nt)
    out2 = to_np(kops.project_phonon_series_mirror(half[20:], eig2, qmap, nq, q_block=q_block, out=out2, t_offset=20, t_total=nt))
    assert np.array_equal(out2, ref)


def test_project_phonon_wide(kops):
    """M = 12 and 30 modes (GaN / larger cells), ragged frame count."""
    rng = np.random.default_rng(5)
    for n_prim, nt, nq in ((4, 131, 3), (10, 70, 2)):
        m = 3 * n_prim
        vc = rng.normal(size=(nt, nq, n_prim, 3)) + 1j * rng.normal(size=(nt, nq, n_prim, 3))
        eig = rng.normal(size=(nq, m, n_prim, 3)) + 1j * rng.normal(size=(nq, m, n_prim, 3))
        vq = to_np(kops.project_phonon(vc, eig))
        ref = np.stack([port.project_onto_phonon(vc[:, q], eig[q]) for q in range(nq)], axis=1)
        assert relerr(vq, ref) < TOL
        # series-major output (one contiguous series per (q, s)), written in two time chunks
        out = kops.project_phonon_series(vc[:50], eig, t_total=nt)
        out = to_np(kops.project_phonon_series(vc[50:], eig, out=out, t_offset=50, t_total=nt))
        assert np.array_equal(out[0].T.reshape(nt, nq, m), vq)


# ------------------------------------------------------------------------------- FFT engine
@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 16, 30, 49, 121, 243, 300, 1000, 2500, 3000, 4096, 4100, 5000, 8192,
                               20000, 32768, 2 * 61 * 61])
def test_fft_engine(kops, n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=(3, n)) + 1j * rng.normal(size=(3, n))
    assert relerr(kops.fft_c2c(x, -1), np.fft.fft(x, axis=1)) < 1e-14
    assert relerr(kops.fft_c2c(x, +1), np.fft.ifft(x, axis=1) * n) < 1e-14


def test_fft_engine_linearity_and_unsupported(kops):
    rng = np.random.default_rng(0)
    n = 6000
    x = rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))
    y = to_np(kops.fft_c2c(x, -1))
    back = to_np(kops.fft_c2c(y, +1)) / n
    assert relerr(back, x) < 1e-14                                  # round trip
    z = to_np(kops.fft_c2c(x[:1] * 2.0 - 1j * x[1:], -1))
    assert relerr(z, 2.0 * y[:1] - 1j * y[1:]) < 1e-14               # linearity
    from dynaphopy_b200._lib import DpError
    with pytest.raises(DpError):
        kops.fft_c2c(x[:, :4099], -1)                                # prime > 61


# ------------------------------------------------------------------------------- A7 / A6 / A5
@pytest.mark.parametrize("name", ["cubic2_232", "si_conv_222", "cubic2_222_complex"])
def test_spectra_vs_reference_golden(kops, name):
    g = golden(name)
    vq = g["vq"][1]
    dt = float(g["time_step_average"])
    f = g["freq"]
    L, pieces = kops.fft_pieces(vq.shape[0], dt, f)
    assert pieces == port.division_of_data(f[1] - f[0], vq.shape[0], dt)
    ps = to_np(kops.power_fft(vq, dt, f))
    assert relerr(ps, g["ps_fft"]) < TOL
    assert np.array_equal(ps.argmax(0), g["ps_fft"].argmax(0))
    cols = g["vc"][1].swapaxes(1, 2).reshape(-1, vq.shape[1])          # dynaphopy/__init__.py:764 layout
    assert relerr(kops.power_fft(cols, dt, f), g["ps_fft_wave_vector_cols"]) < TOL
    pm = to_np(kops.power_mem(vq, dt, f, int(g["mem_order"])))
    assert relerr(pm, g["ps_mem"]) < TOL
    assert np.array_equal(pm.argmax(0), g["ps_mem"].argmax(0))
    for method in (1, 0):
        ref = g["ps_corr_m%d" % method]
        pc = to_np(kops.power_correlation(vq, dt, f, 10, method, 0))
        assert np.array_equal(np.isfinite(pc), np.isfinite(ref))        # non-finite where the reference is
        assert relerr(pc, ref) < TOL
        ref = g["ps_corr_intended_m%d" % method]
        pc = to_np(kops.power_correlation(vq, dt, f, 10, method, 1, scale=1.0))
        assert relerr(pc, ref) < TOL


def _series(rng, nt, ns, dt):
    t = np.arange(nt) * dt
    x = rng.normal(size=(nt, ns)) + 1j * rng.normal(size=(nt, ns))
    for f0, a0 in ((3.83, 3.0), (7.11, 2.0), (15.53, 1.5)):
        x += a0 * np.exp(2j * np.pi * f0 * t[:, None] + 1j * rng.uniform(0, 2 * np.pi, size=(1, ns)))
    return x


@pytest.mark.parametrize("nt,dt,res,fmax", [(2500, 0.002, 0.05, 18.2),      # cfg 1/2: single window L = T
                                            (3000, 0.001, 0.05, 25.0),      # cfg 4 (Ag2Cu2O4)
                                            (1201, 0.002, 2.0, 40.0),       # L = 250, 5 windows, odd T
                                            (700, 0.001, 3.0, 60.0),        # odd L = 333, grid beyond Nyquist? no: clamps
                                            (640, 0.002, 1.7, 300.0),       # grid far beyond +Nyquist (np.interp clamps)
                                            (2000, 0.001, 8.0 / 3.0, 60.0),  # odd L = 375, 6 windows: paired windows, no lag -L/2
                                            (4000, 0.001, 1.0, 40.0),        # even L = 1000, 4 windows: two pairs
                                            (2300, 0.002, 1.0, 40.0)])       # L = 500, 5 windows: two pairs + a single
def test_power_fft_vs_oracle(kops, nt, dt, res, fmax):
    rng = np.random.default_rng(nt)
    ns = 3 if not is_cuda(kops) else 12
    x = _series(rng, nt, ns, dt)
    f = np.arange(0, fmax + res, res)
    ps = to_np(kops.power_fft(x, dt, f))
    ref = port.fft_spectra(x, dt, f)
    assert ps.shape == ref.shape
    assert relerr(ps, ref) < TOL
    assert np.array_equal(ps.argmax(0), ref.argmax(0))


def test_power_fft_negative_grid_and_ld(kops):
    """Grid with negative frequencies and a strided (T, ld) parent array."""
    rng = np.random.default_rng(3)
    x = _series(rng, 900, 5, 0.002)
    f = np.arange(-30.0, 30.0, 0.7)
    ps = to_np(kops.power_fft(x[:, 1:4].copy(), 0.002, f))
    ref = port.fft_spectra(x[:, 1:4], 0.002, f)
    assert relerr(ps, ref) < TOL


def test_power_fft_long_windows(kops):
    """cfg-3/5 window geometry: L = 20000 (four-step, radix 5) with overlapping windows."""
    nt = 50001 if is_cuda(kops) else 20011
    rng = np.random.default_rng(7)
    x = _series(rng, nt, 2 if not is_cuda(kops) else 6, 0.001)
    f = np.arange(0, 31.3, 0.05)
    L, pieces = kops.fft_pieces(nt, 0.001, f)
    assert L == 20000 and pieces == port.division_of_data(0.05, nt, 0.001)
    ps = to_np(kops.power_fft(x, 0.001, f))
    ref = port.fft_spectra(x, 0.001, f, use_fft_correlate=True)      # FFT-based np.correlate restatement
    assert relerr(ps, ref) < TOL
    assert np.array_equal(ps.argmax(0), ref.argmax(0))


@pytest.mark.parametrize("nt,dt,res,nf", [(6000, 0.001, None, 2500),      # P = 16384 but 2500 kept bins: too wide for the v2 bin pass
                                          (9000, 0.002, 0.125, 2400),     # L = 4000, three windows, wide bin range
                                          (45000, 0.001, None, 300),      # P = 131072 = 32 x 64 x 64, n1' = 150 (15 x 10)
                                          (100000, 0.001, 1.0 / 45.056, 700)])   # L = 45056 = 2^12 * 11 (n1' = 256), three windows
def test_power_fft_long_window_engine(kops, nt, dt, res, nf):
    """Windows with P > 65536 or wide kept-bin ranges: the three-factor batched engine (dp_spectrum_fft3.cuh)."""
    rng = np.random.default_rng(nt + 1)
    ns = 2 if not is_cuda(kops) else 19
    x = _series(rng, nt, ns, dt)
    if res is None:
        res = 1.0 / (nt * dt)
    f = np.arange(nf) * res
    assert kops.fft_engine(nt, dt, f) == 3
    ps = to_np(kops.power_fft(x, dt, f))
    ref = port.fft_spectra(x, dt, f, use_fft_correlate=True)
    assert ps.shape == ref.shape
    assert relerr(ps, ref) < TOL
    assert np.array_equal(ps.argmax(0), ref.argmax(0))
    # series-major and blocked-time layouts (what the multi-GPU exchange delivers) give the same bits
    xs = np.ascontiguousarray(x.T)
    ps2 = to_np(kops.power_fft(xs, dt, f, series_major=True))
    assert np.array_equal(ps, ps2)
    if nt % 4 == 0:
        xb = np.ascontiguousarray(xs.reshape(ns, 4, nt // 4).transpose(1, 0, 2))       # (G, S, Tb)
        ps3 = to_np(kops.power_fft(xb, dt, f, series_major=True))
        assert np.array_equal(ps, ps3)


@pytest.mark.gpu
def test_power_fft_cfg5b_single_window(cuda_ops):
    """cfg 5-B: resolution 2^-17/dt so that ONE window covers the whole 2^17-sample run
    (the reference's piece list is the same window twice, power_spectrum/__init__.py:33-51), 5244 frequencies.
    Against the oracle with the FFT-based restatement of np.correlate (the O(L^2) form takes minutes per series)."""
    nt, dt = 131072, 0.001
    rng = np.random.default_rng(517)
    x = _series(rng, nt, 3, dt)
    f = np.arange(5244) * (1.0 / (nt * dt))
    L, pieces = cuda_ops.fft_pieces(nt, dt, f)
    assert L == nt and pieces == [[0, nt], [0, nt]] == port.division_of_data(f[1] - f[0], nt, dt)
    ref = port.fft_spectra(x, dt, f, use_fft_correlate=True)
    for lay, arr in (("time-major", x), ("series-major", np.ascontiguousarray(x.T))):
        ps = to_np(cuda_ops.power_fft(arr, dt, f, series_major=(lay == "series-major")))
        assert ps.shape == ref.shape == (5244, 3)
        assert relerr(ps, ref) < TOL, lay
        assert np.array_equal(ps.argmax(0), ref.argmax(0)), lay


@pytest.mark.gpu
def test_c_abi_calls_do_not_block_the_host(cuda_ops):
    """include/dynaphopy_b200.h: every entry point only ENQUEUES work.  A long-running job is put on the stream first;
    the calls that used to synchronise (commensurate projection, dense projection, FFT finish, direct correlation,
    MEM) are then queued behind it 20 times each -- if any of them waited for the stream, the host loop would take at
    least as long as the job in front of it."""
    import time
    import torch
    ops = cuda_ops
    rng = np.random.default_rng(5)
    g = golden("cubic2_232")
    dims, unit_type, n_prim, qb, tw, _ = _commensurate_inputs(g)
    dev = ops.be.device
    v = torch.as_tensor(g["velocity"], device=dev)
    d_qb = torch.as_tensor(qb, device=dev)
    d_tw = torch.as_tensor(tw, device=dev)
    sm = torch.as_tensor(np.sqrt(g["masses"]), device=dev)
    pos = torch.as_tensor(g["positions"], device=dev)
    x = torch.as_tensor(_series(rng, 4000, 4, 0.002), device=dev)
    f = np.arange(0, 30.0, 0.5)
    big = torch.randn((296, 131072), dtype=torch.complex128, device=dev)
    fbig = np.arange(0, 40.05, 0.05)

    def small_calls():
        ops.project_commensurate(v, dims, unit_type, n_prim, d_qb, d_tw)
        ops.project_wave_vector(v, sm, pos, g["atom_type"], n_prim, g["q_cart"])
        ops.power_fft(x, 0.002, f)
        ops.power_correlation(x, 0.002, f, 10, 1, 1)
        ops.power_mem(x, 0.002, f, 50)

    small_calls()
    ops.power_fft(big, 0.001, fbig, series_major=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(24):
        ops.power_fft(big, 0.001, fbig, series_major=True)          # ~30 ms of device work in front
    h0 = time.perf_counter()
    for _ in range(20):
        small_calls()
    host_ms = (time.perf_counter() - h0) * 1e3
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
    assert host_ms < 0.5 * dev_ms, "host enqueue loop %.1f ms vs %.1f ms of device work: a call waited" % (host_ms, dev_ms)


def test_power_mem_vs_oracle(kops):
    cuda = is_cuda(kops)
    nt, m, ns = (2500, 1000, 6) if cuda else (900, 120, 2)          # cfg 2 on the GPU
    rng = np.random.default_rng(11)
    x = _series(rng, nt, ns, 0.002)
    x[:, -1] = x[:, -1].real                                         # Im part all zero -> skipped part (xms == 0)
    f = np.arange(0, 18.25, 0.05)
    pm = to_np(kops.power_mem(x, 0.002, f, m))
    ref = port.mem_spectra(x, 0.002, f, m)
    assert relerr(pm, ref) < TOL
    assert np.array_equal(pm.argmax(0), ref.argmax(0))
    from dynaphopy_b200._lib import DpError
    with pytest.raises(DpError, match="Number of coefficients"):    # power_spectrum/__init__.py:85-87
        kops.power_mem(x[: m + 1], 0.002, f, m)


@pytest.mark.gpu
def test_power_mem_long_series_cluster(cuda_ops):
    """Series too long for one CTA's shared memory: the thread-block-cluster Burg kernel (f in registers, b in
    shared memory, one cluster barrier per order) for its three block sizes, against the oracle."""
    rng = np.random.default_rng(43)
    f = np.arange(0, 40.0, 0.25)
    for nt, m, ns in ((20001, 64, 3), (40000, 200, 2), (65537, 100, 2), (131072, 1000, 2), (131073, 300, 1), (100003, 50, 20)):
        x = _series(rng, nt, ns, 0.001)
        if ns > 1:
            x[:, -1] = x[:, -1].real                                  # skipped part (xms == 0)
        pm = to_np(cuda_ops.power_mem(x, 0.001, f, m))
        ref = port.mem_spectra(x, 0.001, f, m)
        assert relerr(pm, ref) < TOL, (nt, m)
        assert np.array_equal(pm.argmax(0), ref.argmax(0))
    # beyond the cluster's capacity (8 x 512 x 32 samples): the streaming single-CTA kernel
    x = _series(rng, 140000, 1, 0.001)
    assert relerr(cuda_ops.power_mem(x, 0.001, f, 30), port.mem_spectra(x, 0.001, f, 30)) < TOL


def test_power_correlation_vs_oracle(kops):
    cuda = is_cuda(kops)
    nt, ns = (2500, 6) if cuda else (777, 2)
    rng = np.random.default_rng(13)
    x = _series(rng, nt, ns, 0.002)
    f = np.arange(0, 40.05, 0.05 if cuda else 0.5)
    for method in (1, 0):
        for step in (10, 3):
            ref = port.correlation_spectra(x, 0.002, f, step, method, "reference")
            pc = to_np(kops.power_correlation(x, 0.002, f, step, method, 0))
            fin_ref, fin = np.isfinite(ref), np.isfinite(pc)
            # overflow boundary: the factored sum may cross DBL_MAX one grid point earlier/later
            mismatch = fin_ref != fin
            assert mismatch.sum() <= 2 * ns
            ok = fin_ref & fin & (np.abs(ref) < 1e290)
            assert ok.any()
            assert np.abs(pc - ref)[ok].max() / np.abs(ref[ok]).max() < TOL
            assert np.all(np.abs(pc[ok] - ref[ok]) <= 1e-9 * np.abs(ref[ok]) + 1e-300)   # element-wise: huge dynamic range
            ref = port.correlation_spectra(x, 0.002, f, step, method, "intended")
            pc = to_np(kops.power_correlation(x, 0.002, f, step, method, 1))
            assert relerr(pc, ref) < TOL
    from dynaphopy_b200._lib import DpError
    with pytest.raises(DpError, match="Integration method"):         # c/correlation.c:138-141
        kops.power_correlation(x, 0.002, f, 10, 2, 0)


def test_power_correlation_lag_tiling_edges(kops):
    """Register-tiled lag sums: several warps / several CTAs per series, ragged residue lengths, T <= step."""
    rng = np.random.default_rng(29)
    f = np.array([0.0, 3.3, 17.0])
    for nt, step in ((3000, 1), (2111, 2), (1237, 1), (300, 7), (523, 10), (21, 10), (11, 10)):
        x = _series(rng, nt, 2, 0.002)
        for method in (1, 0):
            ref = port.correlation_spectra(x, 0.002, f, step, method, "intended")
            pc = to_np(kops.power_correlation(x, 0.002, f, step, method, 1))
            assert relerr(pc, ref) < TOL, (nt, step, method)
    # no lag has an origin (N <= step): 0 * dt / (N div step), NaN when N < step -- c/correlation.c:153,177
    for nt in (10, 5):
        x = _series(rng, nt, 2, 0.002)
        ref = port.correlation_spectra(x, 0.002, f, 10, 1, "intended")
        pc = to_np(kops.power_correlation(x, 0.002, f, 10, 1, 1))
        assert np.array_equal(np.isnan(pc), np.isnan(ref))
        assert np.array_equal(np.nan_to_num(pc), np.nan_to_num(ref))


def test_spectra_single_column_and_two_frequencies(kops):
    rng = np.random.default_rng(17)
    x = _series(rng, 512, 1, 0.002)
    f = np.array([1.0, 9.0])
    assert relerr(kops.power_fft(x, 0.002, f), port.fft_spectra(x, 0.002, f)) < TOL
    assert relerr(kops.power_mem(x, 0.002, f, 40), port.mem_spectra(x, 0.002, f, 40)) < TOL
    assert relerr(kops.power_correlation(x, 0.002, f, 10, 1, 1), port.correlation_spectra(x, 0.002, f, 10, 1, "intended")) < TOL


def test_power_fft_series_major_and_time_blocks(kops):
    """Series-major (S, T) input and the blocked time axis (G, S, Tb) the all-to-all delivers give the
    same spectra as the reference's (T, S) layout."""
    rng = np.random.default_rng(23)
    nt, ns = 1200, 3
    x = _series(rng, nt, ns, 0.002)
    f = np.arange(0, 40.0, 2.0)
    ref = port.fft_spectra(x, 0.002, f)
    a = to_np(kops.power_fft(x, 0.002, f))
    b = to_np(kops.power_fft(np.ascontiguousarray(x.T), 0.002, f, series_major=True))
    blocks = np.ascontiguousarray(x.reshape(4, nt // 4, ns).transpose(0, 2, 1))       # (G, S, Tb)
    c = to_np(kops.power_fft(blocks, 0.002, f, series_major=True))
    assert relerr(a, ref) < TOL
    assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.parametrize("nt,dt,res", [(4096, 0.002, 0.5), (8192, 0.001, 1.0 / 3.0), (2048, 0.002, 2.0)])
def test_power_fft_series_rows_as_tensor_boxes(kops, nt, dt, res):
    """Series-major input whose rows start on multiples of the row length of the padded transform: pass 1 fetches its
    window tiles as boxes of the series tensor (bulk copies); every window start (any offset inside a row), the last
    window touching the end of the array, and the result identical to the per-lane load path of the (T, S) layout."""
    rng = np.random.default_rng(nt)
    ns = 3
    x = _series(rng, nt, ns, dt)
    f = np.arange(0, 30.0, res)
    L, pieces = kops.fft_pieces(nt, dt, f)
    assert len(pieces) >= 2 and any(p[0] % 64 for p in pieces)
    ref = port.fft_spectra(x, dt, f)
    a = to_np(kops.power_fft(x, dt, f))
    b = to_np(kops.power_fft(np.ascontiguousarray(x.T), dt, f, series_major=True))
    assert relerr(a, ref) < TOL and relerr(b, ref) < TOL
    assert np.array_equal(a, b)
    assert np.array_equal(b.argmax(0), ref.argmax(0))


def test_power_mem_and_correlation_series_major_and_time_blocks(kops):
    """dp_power_mem_strided / dp_power_correlation_strided: series-major (S, T) input and the blocked time axis
    (G, S, Tb) of the multi-GPU exchange give exactly the spectra of the reference's (T, S) layout."""
    rng = np.random.default_rng(31)
    nt, ns, dt = 1200, 5, 0.002
    x = _series(rng, nt, ns, dt)
    f = np.arange(0, 40.0, 2.0)
    xt = np.ascontiguousarray(x.T)
    blocks = np.ascontiguousarray(x.reshape(4, nt // 4, ns).transpose(0, 2, 1))       # (G, S, Tb)
    a = to_np(kops.power_mem(x, dt, f, 40))
    assert relerr(a, port.mem_spectra(x, dt, f, 40)) < TOL
    assert np.array_equal(a, to_np(kops.power_mem(xt, dt, f, 40, series_major=True)))
    assert np.array_equal(a, to_np(kops.power_mem(blocks, dt, f, 40, series_major=True)))
    for method in (0, 1):
        c = to_np(kops.power_correlation(x, dt, f, 10, method, 1))
        assert relerr(c, port.correlation_spectra(x, dt, f, 10, method, "intended")) < TOL
        assert np.array_equal(c, to_np(kops.power_correlation(xt, dt, f, 10, method, 1, series_major=True)))
        assert np.array_equal(c, to_np(kops.power_correlation(blocks, dt, f, 10, method, 1, series_major=True)))


def test_power_fft_preemptible_form_is_bit_identical(kops):
    """dp_power_fft_windows_preemptible (one short-lived CTA per (series, window pair), slabs from a free list) gives the
    bits of the persistent form, for pairs, a trailing single window and an incremental window range."""
    rng = np.random.default_rng(77)
    nt, dt, ns = 2600, 0.002, 21
    x = np.ascontiguousarray(_series(rng, nt, ns, dt).T)                       # (S, T) series-major
    f = np.arange(0, 30.0, 0.5)
    L, pieces = kops.fft_pieces(nt, dt, f)
    npieces = len({tuple(pc) for pc in pieces})
    assert kops.fft_engine(nt, dt, f) == 2 and npieces >= 3
    ref = to_np(kops.power_fft(x, dt, f, series_major=True))
    ws = kops.power_fft_begin(nt, ns, dt, f)
    kops.power_fft_windows(x, dt, f, 0, 2, ws, series_major=True, preemptible=True)
    kops.power_fft_windows(x, dt, f, 2, npieces - 2, ws, series_major=True, preemptible=True, max_ctas=3)
    out = to_np(kops.power_fft_finish(nt, ns, dt, f, ws))
    assert np.array_equal(out, ref)


def test_power_fft_incremental_windows(kops):
    """dp_power_fft_windows / dp_power_fft_finish: windows transformed in several calls (any split, even while
    later samples are still garbage) give exactly the one-shot spectrum."""
    rng = np.random.default_rng(29)
    nt, ns, dt = 2300, 3, 0.002
    x = _series(rng, nt, ns, dt)
    f = np.arange(0, 40.0, 1.0)                              # L = 500, 5 windows
    L, pieces = kops.fft_pieces(nt, dt, f)
    assert len(pieces) == 5
    ref = to_np(kops.power_fft(x, dt, f))
    xs = np.ascontiguousarray(x.T)                           # series-major
    for split in ([2, 2, 1], [1, 1, 1, 1, 1], [4, 1], [5]):
        ws = kops.power_fft_begin(nt, ns, dt, f)
        part = xs.copy()
        first = 0
        for n in split:
            end = pieces[first + n - 1][1]
            part[:, end:] = np.nan                           # frames after the requested windows do not exist yet
            kops.power_fft_windows(part, dt, f, first, n, ws, series_major=True)
            part = xs.copy()
            first += n
        out = to_np(kops.power_fft_finish(nt, ns, dt, f, ws))
        assert relerr(out, ref) < 1e-13
        if split == [2, 2, 1]:
            assert np.array_equal(out, ref)                  # same pairing as the one-shot call: bitwise


def test_spectra_peaks_degenerate_average_and_guess(kops):
    """dp_spectra_peaks against analysis/fitting/__init__.py:17-39, 60-61 (degenerate_sets / average_phonon /
    np.max + np.argmax), restated with numpy exactly as the reference writes them."""
    from dynaphopy_b200.fitting import degenerate_sets
    rng = np.random.default_rng(41)
    F, Q, M = 157, 5, 6
    ps = rng.random((F, Q * M)) ** 4
    ps[40, 3] = ps[90, 3] = 7.0                # tie: the first maximum wins (np.argmax)
    ps[11, 8] = np.nan                         # NaN wins np.max / np.argmax
    ps[50, 8] = np.nan
    freqs = np.array([[0, 0, 0, 3.1, 3.1 + 5e-5, 9.0], [1, 2, 3, 4, 5, 6], [2, 2, 2, 2, 2, 2], [1, 5, 1, 5, 1, 5],
                      [0.0, 0.00009, 0.00018, 7, 8, 8]])
    sets_q = [degenerate_sets(f) for f in freqs]
    assert sets_q[0] == [[0, 1, 2], [3, 4], [5]] and sets_q[4] == [[0, 1, 2], [3], [4, 5]]    # chained membership
    flat = [[i * M + j for j in g] for i, sq in enumerate(sets_q) for g in sq]
    avg, bins, heights = (to_np(a) for a in kops.spectra_peaks(ps, flat))
    for i in range(Q):
        for j in range(M):
            members = [g for g in sets_q[i] if j in g][0]
            ref = np.average([ps[:, i * M + k] for k in members], axis=0)           # average_phonon
            col = i * M + j
            assert np.array_equal(avg[:, col], ref, equal_nan=True)
            assert bins[col] == np.argmax(ref)
            assert np.array_equal(heights[col], np.max(ref), equal_nan=True)
    # without sets: plain guess per series, no averaged output
    avg0, bins0, heights0 = kops.spectra_peaks(ps)
    assert avg0 is None
    assert np.array_equal(to_np(bins0), np.argmax(ps, axis=0))
    assert np.array_equal(to_np(heights0), np.max(ps, axis=0), equal_nan=True)
    with pytest.raises(ValueError):
        kops.spectra_peaks(ps, flat[:-1])


def test_project_phonon_series_peer_blocks(kops):
    """dp_project_phonon_series_peers: every destination block at its own base pointer (peer receive buffers on a
    multi-GPU box) holds exactly what the single-array send layout holds for that block."""
    rng = np.random.default_rng(47)
    T, Q, M, qb = 70, 12, 6, 3
    vc = rng.normal(size=(T, Q, M)) + 1j * rng.normal(size=(T, Q, M))
    eig = rng.normal(size=(Q, M, M)) + 1j * rng.normal(size=(Q, M, M))
    t_total, t_off = 100, 20
    ref = to_np(kops.project_phonon_series(vc, eig, q_block=qb, t_offset=t_off, t_total=t_total,
                                           out=kops.be.asarray(np.zeros((Q // qb, qb * M, t_total), complex), np.complex128)))
    blocks = [kops.be.asarray(np.zeros((qb * M, t_total), complex), np.complex128) for _ in range(Q // qb)]
    assert kops.project_phonon_series(vc, eig, q_block=qb, t_offset=t_off, t_total=t_total,
                                      block_ptrs=[kops.be.ptr(b).value for b in blocks]) is None
    for g, b in enumerate(blocks):
        assert np.array_equal(to_np(b), ref[g])
    assert np.abs(ref[:, :, t_off:t_off + T]).min() > 0 and not ref[:, :, :t_off].any()
    from dynaphopy_b200._lib import DpError
    with pytest.raises(DpError, match="block pointers"):
        kops.project_phonon_series(vc, eig, q_block=qb, t_total=T, block_ptrs=[kops.be.ptr(blocks[0]).value])


def test_small_and_ragged_shapes_of_the_reductions(kops):
    """Edge shapes of the two reduction kernels: fewer frames than one time split, atom counts that do not fill a
    CTA, fewer frequencies than a warp, a single series, all-equal and all-NaN spectra."""
    rng = np.random.default_rng(53)
    cell = np.diag([5.0, 6.0, 7.0])
    i, j = np.triu_indices(3)
    for nt, na in ((1, 1), (3, 5), (17, 257), (40, 300)):
        pos = rng.random((na, 3)) @ cell
        tr = pos[None] + rng.normal(scale=0.2, size=(nt, na, 3)) + rng.integers(-1, 2, size=(nt, na, 3)) @ cell
        u = np.stack([port.atomic_displacements(tr[:, a, :], pos[a], cell) for a in range(na)], axis=1)
        ref = np.concatenate([u.sum(0), (u[:, :, i] * u[:, :, j]).sum(0)], axis=1)
        sums = to_np(kops.displacement_moments(tr, pos, cell))
        assert np.abs(sums - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1e-300), (nt, na)
    for F, S in ((1, 1), (5, 3), (31, 2), (33, 70)):
        ps = rng.random((F, S))
        _, bins, heights = kops.spectra_peaks(ps)
        assert np.array_equal(to_np(bins), ps.argmax(0)) and np.array_equal(to_np(heights), ps.max(0))
    flat = np.full((40, 2), 3.0)                      # all equal: the first bin wins
    flat[:, 1] = np.nan                               # all NaN: bin 0, NaN height
    _, bins, heights = kops.spectra_peaks(flat)
    assert list(to_np(bins)) == [0, 0]
    h = to_np(heights)
    assert h[0] == 3.0 and np.isnan(h[1])


def test_randomised_shapes_property(kops):
    """Randomly drawn shapes (seeded) for the kernels whose index arithmetic depends on ragged sizes: decimated lag
    sums (any T / step), per-destination contraction blocks (any T, Q, q_block, both vc layouts), degenerate-set
    partitions."""
    rng = np.random.default_rng(79)
    f = np.array([0.5, 7.0, 21.0])
    cuda = is_cuda(kops)
    for _ in range(6 if cuda else 4):
        nt = int(rng.integers(12, 1400 if cuda else 500))
        step = int(rng.integers(1, 13))
        x = _series(rng, nt, 2, 0.002)
        for method in (1, 0):
            ref = port.correlation_spectra(x, 0.002, f, step, method, "intended")
            pc = to_np(kops.power_correlation(x, 0.002, f, step, method, 1))
            if np.isnan(ref).any():
                assert np.array_equal(np.isnan(pc), np.isnan(ref)), (nt, step)
            else:
                assert relerr(pc, ref) < TOL, (nt, step, method)
    for _ in range(4):
        M = int(rng.choice([3, 6, 12]))
        nb = int(rng.integers(1, 5))
        qb = int(rng.integers(1, 7))
        Q, T = nb * qb, int(rng.integers(1, 90))
        vc = rng.normal(size=(T, Q, M)) + 1j * rng.normal(size=(T, Q, M))
        eig = rng.normal(size=(Q, M, M)) + 1j * rng.normal(size=(Q, M, M))
        ref = np.einsum("tqj,qsj->qst", vc, np.conj(eig)).reshape(nb, qb * M, T)
        blocks = [kops.be.asarray(np.zeros((qb * M, T), complex), np.complex128) for _ in range(nb)]
        kops.project_phonon_series(vc, eig, q_block=qb, t_total=T, block_ptrs=[kops.be.ptr(b).value for b in blocks])
        for g, b in enumerate(blocks):
            assert relerr(to_np(b), ref[g]) < 1e-13, (T, Q, M, qb)
    for _ in range(4):
        F, S = int(rng.integers(1, 70)), int(rng.integers(1, 40))
        ps = rng.random((F, S))
        perm = rng.permutation(S)
        cuts = np.sort(rng.choice(np.arange(1, S), size=min(S - 1, int(rng.integers(0, 6))), replace=False)) if S > 1 else []
        sets = [list(map(int, g)) for g in np.split(perm, cuts)]
        avg, bins, heights = (to_np(a) for a in kops.spectra_peaks(ps, sets))
        for g in sets:
            ref = np.average([ps[:, k] for k in g], axis=0)
            for k in g:
                assert np.array_equal(avg[:, k], ref) and bins[k] == ref.argmax() and heights[k] == ref.max()
