"""The reference-facing Python API (Dynamics / projection / power_spectrum selector / Quasiparticle) on top
of the kernels, driven the way the reference's own scripts and unit tests drive it
(unittest/Si_test.py:29-41: build a Dynamics, make a Quasiparticle, set q, select the algorithm, read the
spectra), and compared with what the REFERENCE returned for the same inputs (tests/golden)."""
import os

import numpy as np
import pytest

from conftest import golden, relerr

TOL = 1e-10


def _patch_default_ops(monkeypatch, kops):
    from dynaphopy_b200 import ops as dops
    monkeypatch.setattr(dops, "_default", kops)


def _dynamics(g, kops):
    from dynaphopy_b200.dynamics import Dynamics
    from dynaphopy_b200.structure import Structure
    st = Structure(cell=g["unit_cell"], positions=g["unit_pos"], masses=g["unit_masses"],
                   primitive_matrix=g["prim_matrix"])
    return Dynamics(structure=st, velocity=g["velocity"].astype(complex), time=g["time"], supercell=g["supercell"],
                    ops=kops)


@pytest.mark.parametrize("name", ["cubic2_232", "si_conv_222", "cubic2_222_complex"])
def test_quasiparticle_path_vs_reference(kops, monkeypatch, name):
    from dynaphopy_b200.quasiparticle import Quasiparticle
    from dynaphopy_b200.parameters import Parameters
    _patch_default_ops(monkeypatch, kops)
    g = golden(name)
    Parameters().reset()
    dyn = _dynamics(g, kops)
    assert dyn.get_time_step_average() == float(g["time_step_average"])
    assert relerr(dyn.get_velocity_mass_average(), g["vm"]) < 1e-15
    calc = Quasiparticle(dyn)
    calc.parameters.use_symmetry = False
    calc.set_frequency_limits([g["freq"][0], g["freq"][-1]])
    calc.parameters.frequency_range = g["freq"]
    for i, q in enumerate(g["q_reduced"]):
        calc.set_reduced_q_vector(q)
        assert np.allclose(calc.get_q_vector(), g["q_cart"][i], rtol=0, atol=1e-13)
        vc = calc.get_vc()
        assert vc.shape == g["vc"][i].shape and vc.dtype == np.complex128
        assert relerr(vc, g["vc"][i]) < TOL
        calc.set_eigenvectors(g["eigenvectors"][i])
        assert relerr(calc.get_vq(), g["vq"][i]) < TOL
    # spectra of q index 1 with the three methods of the selector
    calc.set_reduced_q_vector(g["q_reduced"][1])
    calc.set_eigenvectors(g["eigenvectors"][1])
    calc.set_number_of_mem_coefficients(int(g["mem_order"]))
    for alg, key in ((2, "ps_fft"), (1, "ps_mem"), (3, "ps_fft")):
        calc.select_power_spectra_algorithm(alg)
        ps = calc.get_power_spectrum_phonon()
        assert ps.shape == g[key].shape
        assert relerr(ps, g[key]) < TOL
        assert np.array_equal(ps.argmax(0), g[key].argmax(0))
    calc.select_power_spectra_algorithm(0)
    ps = calc.get_power_spectrum_phonon()
    assert np.array_equal(np.isfinite(ps), np.isfinite(g["ps_corr_m1"]))
    assert relerr(ps, g["ps_corr_m1"]) < TOL
    calc.select_power_spectra_algorithm(2)
    assert relerr(calc.get_power_spectrum_wave_vector(), np.nansum(g["ps_fft_wave_vector_cols"], axis=1)) < TOL
    # project_on_atom keeps the normalisation (projection.py:26-28, 38-39)
    calc.set_projection_onto_atom_type(0)
    calc.full_clear()
    assert relerr(calc.get_vc(), g["vc_atom0"]) < TOL
    Parameters().reset()


@pytest.mark.parametrize("name", ["cubic2_232", "si_conv_222", "cubic2_222_complex"])
def test_full_and_partial_spectra_vs_reference(kops, monkeypatch, name):
    """Quasiparticle.get_power_spectrum_full / _partials (dynaphopy/__init__.py:781-854) against what the reference
    itself returned (oracle/gen_golden.py::_full_and_partial_spectra): all atoms, one atom type, one atom type and
    one coordinate, `silent` both ways, FFT and MEM methods; plus the reference's exits and its caching."""
    from dynaphopy_b200.quasiparticle import Quasiparticle
    from dynaphopy_b200.parameters import Parameters
    _patch_default_ops(monkeypatch, kops)
    g = golden(name)
    pr = Parameters()
    pr.reset()
    dyn = _dynamics(g, kops)
    calc = Quasiparticle(dyn)
    pr.frequency_range = g["freq"]
    pr.number_of_coefficients_mem = int(g["mem_order"])

    def full(alg, atom, coord, silent):
        pr.power_spectra_algorithm = alg
        pr.project_on_atom = atom
        pr.silent = silent
        calc.power_spectra_clear()
        return calc.get_power_spectrum_full(projection_on_coordinate=coord)

    cases = [("ps_full_fft", 2, -1, -1, False), ("ps_full_fft_silent", 2, -1, -1, True),
             ("ps_full_fft_coord1_ignored", 2, -1, 1, False), ("ps_full_fft_atom0", 2, 0, -1, False),
             ("ps_full_fft_atom1_silent", 2, 1, -1, True), ("ps_full_fft_atom0_coord1", 2, 0, 1, False),
             ("ps_full_fft_atom1_coord2_silent", 2, 1, 2, True), ("ps_full_mem", 1, -1, -1, False),
             ("ps_full_mem_atom0_coord0", 1, 0, 0, False)]
    for key, alg, atom, coord, silent in cases:
        ps = full(alg, atom, coord, silent)
        assert ps.shape == g[key].shape == (len(g["freq"]),), key
        assert relerr(ps, g[key]) < TOL, key
        assert int(np.argmax(ps)) == int(np.argmax(g[key])), key
    # cached whatever the arguments until power_spectra_clear (the reference's `is None` test, :789)
    again = calc.get_power_spectrum_full(projection_on_coordinate=2)
    assert again is calc.get_power_spectrum_full()
    # the reference's exits: unknown atom type, coordinate >= number of dimensions (only with an atom type)
    pr.project_on_atom = 5
    calc.power_spectra_clear()
    with pytest.raises(SystemExit):
        calc.get_power_spectrum_full()
    pr.project_on_atom = 0
    with pytest.raises(SystemExit):
        calc.get_power_spectrum_full(projection_on_coordinate=3)
    pr.project_on_atom = -1
    pr.silent = False
    for alg, key in ((2, "ps_partials_fft"), (1, "ps_partials_mem")):
        pr.power_spectra_algorithm = alg
        calc.power_spectra_clear()
        pp = calc.get_power_spectrum_partials()
        assert pp.shape == g[key].shape
        assert relerr(pp, g[key]) < TOL, key
        assert np.array_equal(pp.argmax(0), g[key].argmax(0)), key
    pr.reset()


def test_selector_contract_and_errors(kops, monkeypatch):
    from dynaphopy_b200 import power_spectrum
    from dynaphopy_b200.quasiparticle import Quasiparticle
    from dynaphopy_b200.parameters import Parameters
    _patch_default_ops(monkeypatch, kops)
    assert sorted(power_spectrum.power_spectrum_functions) == [0, 1, 2, 3, 4, 5]
    assert power_spectrum.power_spectrum_functions[1][1] == 'Maximum entropy method'
    g = golden("cubic2_232")
    Parameters().reset()
    calc = Quasiparticle(_dynamics(g, kops))
    with pytest.raises(SystemExit):
        calc.select_power_spectra_algorithm(9)             # reference prints the menu and exit()s
    with pytest.raises(SystemExit):
        calc.set_projection_onto_atom_type(7)
    calc.parameters.number_of_coefficients_mem = 10 ** 6
    with pytest.raises(SystemExit):                        # power_spectrum/__init__.py:85-87
        power_spectrum.get_mem_power_spectra(g["vq"][0], calc.dynamic, calc.parameters)
    # cropping keeps the last frames and invalidates the caches (dynamics.py:65-107)
    calc2 = Quasiparticle(_dynamics(g, kops), last_steps=100)
    assert calc2.dynamic.get_velocity_mass_average().shape[0] == 100
    assert relerr(calc2.dynamic.get_velocity_mass_average(), g["vm"][-100:]) < 1e-15
    assert power_spectrum._division_of_data(0.05, 50001, 0.001) == [[0, 20000], [15000, 35000], [30001, 50001]]
    Parameters().reset()


@pytest.mark.parametrize("tag", ["xdatcar", "lammps"])
def test_dynamics_displacement_statistics_vs_reference(kops, tag):
    """Dynamics.get_mean_displacement_matrix / average_positions (dynamics.py:207-295) as unittest/reading_test.py
    drives them, against what the reference returned (and its literals, reading_test.py:54-60, 94-99)."""
    from dynaphopy_b200.dynamics import Dynamics
    from dynaphopy_b200.structure import Structure
    g = golden("reading_test")
    st = Structure(cell=g["unit_cell"], positions=g["unit_pos"], masses=g["unit_masses"], primitive_matrix=g["prim_matrix"])
    assert np.array_equal(st.get_atom_type_index(), g["unit_atom_type"])

    def make():
        return Dynamics(structure=st, trajectory=g[tag + "_trajectory"].astype(complex), supercell=g[tag + "_supercell"],
                        time=np.arange(12) * float(g[tag + "_time_step"]), ops=kops)
    dyn = make()
    assert np.array_equal(dyn.get_supercell_matrix(), [2, 2, 2])
    mats = dyn.get_mean_displacement_matrix()
    assert relerr(mats, g[tag + "_mean_matrix_types"]) < 1e-12
    assert np.allclose(np.average(mats, axis=0), g[tag + "_mean_matrix_literal"], rtol=1e-3, atol=1e-6)
    assert relerr(make().get_mean_displacement_matrix(use_average_positions=True), g[tag + "_mean_matrix_avgpos"]) < 1e-12
    ap = dyn.average_positions()
    assert ap.dtype == np.complex128 and relerr(ap, g[tag + "_average_positions"]) < 1e-14
    assert relerr(dyn.average_positions(to_unit_cell=True), g[tag + "_average_positions_unit"]) < 1e-14
    # the reference's rel_traj check is a difference of O(10) numbers compared at 1e-9: only meaningful to ~1e-14 absolute
    rel_traj = np.array([np.average(ap.real - traj, axis=0).real for traj in g[tag + "_trajectory"]])
    assert np.abs(rel_traj - g[tag + "_rel_traj"]).max() < 1e-13


def test_commensurate_points_batch_matches_q_loop(kops, monkeypatch):
    """Quasiparticle.get_commensurate_points_spectra (all q in one pass) against the reference-style loop over q
    (set_reduced_q_vector -> get_power_spectrum_phonon, dynaphopy/__init__.py:1130-1174) on the same object."""
    from dynaphopy_b200.quasiparticle import Quasiparticle
    from dynaphopy_b200.parameters import Parameters
    from dynaphopy_b200.fitting import degenerate_sets
    _patch_default_ops(monkeypatch, kops)
    g = golden("cubic2_232")
    Parameters().reset()
    dyn = _dynamics(g, kops)
    calc = Quasiparticle(dyn)
    calc.parameters.frequency_range = g["freq"]
    calc.parameters.use_symmetry = True
    calc.select_power_spectra_algorithm(2)
    com = dyn.structure.get_commensurate_points(dyn.get_supercell_matrix())
    assert com.shape == (12, 3)
    rng = np.random.default_rng(3)
    eig = np.stack([np.linalg.qr(rng.normal(size=(6, 6)) + 1j * rng.normal(size=(6, 6)))[0].T.reshape(6, 2, 3) for _ in com])
    fr = np.tile(np.array([0.0, 0.0, 0.0, 5.0, 5.0, 9.0]), (len(com), 1))
    data = calc.get_commensurate_points_spectra(eig, frequencies=fr, com_points=com)
    assert data["power_spectra"].shape == (12, g["freq"].size, 6)
    for i, q in enumerate(com):
        calc.set_reduced_q_vector(q)
        calc.set_eigenvectors(eig[i], frequencies=fr[i])
        ps = calc.get_power_spectrum_phonon()
        assert relerr(data["power_spectra"][i], ps) < TOL
        sets = degenerate_sets(fr[i])
        assert sets == data["degenerate_sets"][i] == [[0, 1, 2], [3, 4], [5]]
        for j in range(6):
            members = [s for s in sets if j in s][0]
            ref = np.average([data["power_spectra"][i][:, k] for k in members], axis=0)
            assert np.array_equal(data["averaged"][i][:, j], ref)
            assert data["guess_position"][i, j] == g["freq"][np.argmax(ref)]
            assert data["guess_height"][i, j] == np.max(ref)
    # parameters.project_on_atom reaches the batched projection like it reaches get_vc in the reference's q loop
    calc.set_projection_onto_atom_type(1)
    data1 = calc.get_commensurate_points_spectra(eig, frequencies=fr, com_points=com)
    for i in (0, 5, 11):
        calc.set_reduced_q_vector(com[i])
        calc.full_clear()
        calc.set_eigenvectors(eig[i], frequencies=fr[i])
        assert relerr(data1["power_spectra"][i], calc.get_power_spectrum_phonon()) < TOL
    assert relerr(data1["power_spectra"], data["power_spectra"]) > 1e-3
    Parameters().reset()


def test_star_of_q_average_matches_loop(kops, monkeypatch):
    """use_symmetry: get_power_spectrum_phonon averages the spectra over the symmetry-equivalent q points
    (dynaphopy/__init__.py:725-739).  phonopy's star and eigenvectors are injected; the batched evaluation (one
    projection pass, one spectrum launch for the whole star) must equal the reference's loop over the members."""
    from dynaphopy_b200.quasiparticle import Quasiparticle
    from dynaphopy_b200.parameters import Parameters
    _patch_default_ops(monkeypatch, kops)
    g = golden("cubic2_232")
    Parameters().reset()
    calc = Quasiparticle(_dynamics(g, kops))
    calc.parameters.frequency_range = g["freq"]
    calc.select_power_spectra_algorithm(2)
    qs = [np.array(q, dtype=float) for q in g["q_reduced"][1:4]]          # three commensurate points play the star

    def harmonic(q):
        i = [k for k, qq in enumerate(g["q_reduced"]) if np.allclose(qq, q)][0]
        return g["eigenvectors"][i], np.arange(6.0)

    calc.set_harmonic_provider(harmonic)
    # the loop of the reference, member by member, without symmetry
    calc.parameters.use_symmetry = False
    looped = []
    for q in qs:
        calc.set_reduced_q_vector(q)
        assert relerr(calc.get_eigenvectors(), harmonic(q)[0]) == 0          # lazily fetched for the current q
        looped.append(calc.get_power_spectrum_phonon())
    # with symmetry: star provider + batched evaluation
    calc.parameters.use_symmetry = True
    calc.set_star_provider(lambda q: qs)
    calc.set_reduced_q_vector(qs[0])
    calc.power_spectra_clear()
    ps = calc.get_power_spectrum_phonon()
    assert ps.shape == looped[0].shape
    assert relerr(ps, np.average(looped, axis=0)) < 1e-12
    assert np.array_equal(calc.get_reduced_q_vector(), qs[0])
    assert calc.get_power_spectrum_phonon() is ps                            # cached until q or the algorithm changes


def _write_dump(path, frames, steps, bounds_lines, labels, extra_cols=0, rng=None):
    with open(path, "w") as f:
        for fr, st in zip(frames, steps):
            f.write("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\n" % (st, fr.shape[0]))
            f.write("ITEM: BOX BOUNDS xy xz yz pp pp pp\n" + "\n".join(bounds_lines) + "\n")
            f.write("ITEM: ATOMS " + labels + " \n")
            for row in fr:
                cols = ["%.10f" % v for v in row] + ["%d" % rng.integers(0, 9) for _ in range(extra_cols)]
                f.write("   " + "   ".join(cols) + " \n")


def test_lammps_ingest_page_aligned_file_without_trailing_newline(tmp_path):
    """A dump whose size is an exact multiple of the page size and whose last byte belongs to a number (no trailing
    newline, e.g. a file still being written): strtod must find a terminator behind the mapping."""
    import mmap as _mmap
    from dynaphopy_b200 import iofile
    from dynaphopy_b200.structure import Structure
    rng = np.random.default_rng(7)
    st = Structure(cell=np.eye(3) * 4.0, positions=np.array([[0.0, 0, 0], [2.0, 2.0, 2.0]]), masses=[28.0855, 69.723])
    bounds = ["0.0 8.0", "0.0 8.0", "0.0 8.0"]
    frames = [rng.normal(scale=3.0, size=(16, 3)) for _ in range(40)]
    path = str(tmp_path / "aligned.lammpstrj")
    with open(path, "w") as f:
        for i, fr in enumerate(frames):
            f.write("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n16\nITEM: BOX BOUNDS pp pp pp\n" % i + "\n".join(bounds) + "\nITEM: ATOMS x y z\n")
            f.write("\n".join(" ".join("%.10f" % v for v in row) for row in fr))
            if i + 1 < len(frames):
                f.write("\n")
    size = os.path.getsize(path)
    pad = (-size) % _mmap.PAGESIZE
    # lengthen the last number with digits that do not change its value class (more decimals) up to the page boundary
    with open(path, "a") as f:
        f.write("1" * pad)
    assert os.path.getsize(path) % _mmap.PAGESIZE == 0
    dyn = iofile.read_lammps_trajectory(path, st, time_step=0.001, pinned=False)
    assert dyn.trajectory.shape == (40, 16, 3)
    last = float(("%.10f" % frames[-1][-1, -1]) + "1" * pad)
    assert dyn.trajectory[-1, -1, -1].real == last
    assert np.array_equal(dyn.trajectory[:-1].real, np.array([[[float("%.10f" % v) for v in row] for row in fr] for fr in frames[:-1]]))


def test_lammps_ingest_synthetic(tmp_path):
    """dp_lammps_scan / dp_lammps_read + the read_lammps_trajectory mirror on dumps written here: positions with a
    tilted box and extra columns, a velocity dump, frame ranges, atom template, and the reference's error paths;
    checked against a line-by-line restatement of the reference's parser (oracle/port.py)."""
    from oracle import port
    from dynaphopy_b200 import iofile
    from dynaphopy_b200._lib import DpError
    from dynaphopy_b200.structure import Structure
    rng = np.random.default_rng(61)
    cell = np.array([[4.0, 0.0, 0.0], [0.4, 4.1, 0.0], [0.3, 0.2, 3.9]])           # rows: lattice vectors (LAMMPS triclinic)
    st = Structure(cell=cell, positions=np.array([[0.0, 0, 0], [2.0, 2.0, 2.0]]), masses=[28.0855, 69.723])
    sc = np.diag([2, 2, 2]) @ cell
    xy, xz, yz = sc[1, 0], sc[2, 0], sc[2, 1]
    bounds = ["%.8f %.8f %.8f" % (min(0, xy, xz, xy + xz), sc[0, 0] + max(0, xy, xz, xy + xz), xy),
              "%.8f %.8f %.8f" % (min(0, yz), sc[1, 1] + max(0, yz), xz), "%.8f %.8f %.8f" % (0.0, sc[2, 2], yz)]
    frames = [rng.normal(scale=3.0, size=(16, 3)) for _ in range(9)]
    steps = [1000 + 5 * i for i in range(9)]
    pos_dump = str(tmp_path / "pos.lammpstrj")
    _write_dump(pos_dump, frames, steps, bounds, "x y z ix iy", extra_cols=2, rng=rng)
    assert iofile.get_trajectory_parser(pos_dump) is iofile.read_lammps_trajectory
    ref, ref_steps, _, ref_labels = port.read_lammps_frames(pos_dump, 3, initial_cut=2, end_cut=7)
    dyn = iofile.read_lammps_trajectory(pos_dump, st, time_step=0.002, initial_cut=2, end_cut=7, pinned=False)
    assert ref.shape == (6, 16, 3) and b"x y" in ref_labels
    sc_ref, tmat = iofile.lammps_supercell(np.array([[float(x) for x in l.split()] for l in bounds]).ravel(), 3, st)
    assert np.allclose(sc_ref, sc, atol=1e-6)
    assert np.array_equal(dyn.trajectory, np.array([np.dot(fr, tmat) for fr in ref]))
    assert np.array_equal(dyn.get_time(), ref_steps * 0.002)
    assert dyn._velocity is None
    # whole file, atom template (argsort reordering), last_steps crop
    template = rng.permutation(16)
    dyn2 = iofile.read_lammps_trajectory(pos_dump, st, time_step=0.002, template=template, last_steps=4, pinned=False, threads=3)
    full, _, _, _ = port.read_lammps_frames(pos_dump, 3)
    assert np.array_equal(dyn2.trajectory, np.array([np.dot(fr[np.argsort(template)], tmat) for fr in full])[-4:])
    # velocity dump
    vel_dump = str(tmp_path / "vel.lammpstrj")
    _write_dump(vel_dump, frames[:3], steps[:3], bounds, "vx vy vz", rng=rng)
    dyn3 = iofile.read_lammps_trajectory(vel_dump, st, pinned=False)                  # default time step warning path
    assert dyn3._velocity is not None and dyn3._velocity.shape == (3, 16, 3)
    # error behaviour: missing file -> exit(); truncated frame -> "Error reading step"; unknown columns -> exit()
    with pytest.raises(SystemExit):
        iofile.read_lammps_trajectory(str(tmp_path / "nope"), st)
    text = open(pos_dump).read().split("\n")
    open(pos_dump, "w").write("\n".join(text[:-6]))
    with pytest.raises(DpError, match="Error reading step 9"):
        iofile.read_lammps_trajectory(pos_dump, st, pinned=False)
    assert iofile.read_lammps_trajectory(pos_dump, st, end_cut=8, pinned=False).trajectory.shape == (8, 16, 3)
    odd = str(tmp_path / "odd.lammpstrj")
    _write_dump(odd, frames[:1], steps[:1], bounds, "fx fy fz", rng=rng)
    with pytest.raises(SystemExit):
        iofile.read_lammps_trajectory(odd, st, pinned=False)


@pytest.mark.skipif(not os.path.isfile("/root/reference/unittest/Si_data/si.lammpstrj"), reason="reference data not mounted")
def test_lammps_ingest_matches_reference_parser():
    """The dump the reference's own reading test uses (unittest/reading_test.py:64-70): bit-identical to what the
    reference's parser returned (tests/golden/reading_test.npz was written by it)."""
    from dynaphopy_b200 import iofile
    from dynaphopy_b200.structure import Structure
    g = golden("reading_test")
    st = Structure(cell=g["unit_cell"], positions=g["unit_pos"], masses=g["unit_masses"], primitive_matrix=g["prim_matrix"])
    dyn = iofile.read_lammps_trajectory("/root/reference/unittest/Si_data/si.lammpstrj", st, initial_cut=3, end_cut=14,
                                        time_step=0.001, pinned=False)
    assert np.array_equal(dyn.trajectory, g["lammps_trajectory"])
    assert np.array_equal(dyn.get_supercell(), g["lammps_supercell"])
    assert dyn.get_time_step_average() == float(g["lammps_time_step"])


def test_xdatcar_ingest_synthetic(tmp_path):
    """dp_xdatcar_scan / dp_xdatcar_read + the read_VASP_XDATCAR mirror on a file written here (two species, several
    frames): frame ranges, the reference's habit of keeping the times of skipped frames, template, truncation error."""
    from dynaphopy_b200 import iofile
    from dynaphopy_b200._lib import DpError
    from dynaphopy_b200.structure import Structure
    rng = np.random.default_rng(67)
    cell = np.array([[8.0, 0.0, 0.0], [0.5, 8.2, 0.0], [0.0, 0.3, 7.8]])
    st = Structure(cell=cell / 2, positions=np.array([[0.0, 0, 0], [2.0, 2.0, 2.0]]), masses=[28.0855, 69.723])
    frames = [rng.random((16, 3)) for _ in range(7)]
    path = str(tmp_path / "XDATCAR")
    with open(path, "w") as f:
        f.write("unknown system\n           1\n")
        for row in cell:
            f.write("  %12.6f %12.6f %12.6f\n" % tuple(row))
        f.write("   Si   Ga\n   8   8\n")
        for k, fr in enumerate(frames):
            f.write("Direct configuration= %5d\n" % (k + 1))
            for row in fr:
                f.write("   %.8f  %.8f  %.8f\n" % tuple(row))
    parsed = np.array([[[float("%.8f" % v) for v in row] for row in fr] for fr in frames])
    dyn = iofile.read_VASP_XDATCAR(path, st, time_step=0.0005, initial_cut=2, end_cut=6, pinned=False, threads=2)
    assert np.array_equal(dyn._scaled_trajectory, parsed[1:6])
    assert np.array_equal(dyn.get_supercell(), np.array([[float("%12.6f" % v) for v in row] for row in cell]))
    assert np.array_equal(dyn.get_time(), np.arange(1, 7) * 0.0005)                  # six times for five frames (:418 vs :429)
    assert np.array_equal(np.asarray(dyn.trajectory), np.dot(parsed[1:6], dyn.get_supercell()))
    template = rng.permutation(16)
    dyn2 = iofile.read_VASP_XDATCAR(path, st, time_step=0.0005, template=template, last_steps=3, pinned=False)
    assert np.array_equal(dyn2._scaled_trajectory, parsed[:, np.argsort(template)][-3:])
    text = open(path).read().split("\n")
    open(path, "w").write("\n".join(text[:-4]))
    with pytest.raises(DpError, match="Error reading step 7"):
        iofile.read_VASP_XDATCAR(path, st, pinned=False)
    with pytest.raises(SystemExit):
        iofile.read_VASP_XDATCAR(str(tmp_path / "nope"), st)


@pytest.mark.skipif(not os.path.isfile("/root/reference/unittest/Si_data/XDATCAR"), reason="reference data not mounted")
def test_xdatcar_ingest_matches_reference_parser():
    """unittest/reading_test.py:23-35: same file, same cuts and template -> bit-identical to the reference's parser."""
    from dynaphopy_b200 import iofile
    from dynaphopy_b200.structure import Structure
    g = golden("reading_test")
    st = Structure(cell=g["unit_cell"], positions=g["unit_pos"], masses=g["unit_masses"], primitive_matrix=g["prim_matrix"])
    dyn = iofile.read_VASP_XDATCAR("/root/reference/unittest/Si_data/XDATCAR", st, initial_cut=3, end_cut=14,
                                   time_step=0.0005, template=g["xdatcar_template"], pinned=False)
    assert np.array_equal(np.asarray(dyn.trajectory), g["xdatcar_trajectory"])
    assert np.array_equal(dyn.get_supercell(), g["xdatcar_supercell"])
    assert np.array_equal(dyn.get_time(), g["xdatcar_time"])
    assert dyn.get_time_step_average() == float(g["xdatcar_time_step"])


def test_atom_order_template_vs_reference_search(tmp_path):
    """get_correct_arrangement (O(N)) against the reference's O(N^2) search (restated in oracle/port.py) on shuffled,
    thermally displaced supercells, and check_atoms_order end to end through the XDATCAR reader."""
    from oracle import port
    from dynaphopy_b200 import iofile
    from dynaphopy_b200.structure import Structure
    rng = np.random.default_rng(71)
    cell = np.array([[4.0, 0.0, 0.0], [0.3, 4.2, 0.0], [0.0, 0.2, 3.9]])
    unit = np.array([[0.1, 0.1, 0.1], [2.3, 2.6, 2.4], [0.3, 2.9, 1.2]])      # (the reference's supercell-size guess
    # round(max(scaled)) needs an atom beyond the middle of the cell in every direction, and no exact half-cell pairs)
    st = Structure(cell=cell, positions=unit, masses=[28.0855, 69.723, 15.999])
    sp = unit @ np.linalg.inv(cell)
    for dims in ((2, 2, 2), (3, 2, 4)):
        pos = st.get_positions(supercell=dims) + rng.normal(scale=0.08, size=(3 * int(np.prod(dims)), 3))
        perm = rng.permutation(len(pos))
        shuffled = pos[perm]
        template = np.asarray(iofile.get_correct_arrangement(shuffled, st))
        assert np.array_equal(template, port.get_correct_arrangement(shuffled, cell, sp))
        assert np.array_equal(shuffled[np.argsort(template)], pos)              # the readers' reordering (:304-306)
    # end to end: an XDATCAR in shuffled order comes back in DynaPhoPy's order
    dims = (2, 2, 2)
    sc = np.diag(dims) @ cell
    pos = st.get_positions(supercell=dims)
    perm = rng.permutation(len(pos))
    path = str(tmp_path / "XDATCAR")
    with open(path, "w") as f:
        f.write("shuffled\n 1\n")
        for row in sc:
            f.write("  %14.8f %14.8f %14.8f\n" % tuple(row))
        f.write(" A B C\n 8 8 8\n")
        for k in range(3):
            f.write("Direct configuration= %5d\n" % (k + 1))
            for row in (pos[perm] + 0.01 * k) @ np.linalg.inv(sc):
                f.write("   %.10f  %.10f  %.10f\n" % tuple(row))
    assert iofile.get_trajectory_parser(path) is iofile.read_VASP_XDATCAR          # keyword sniffing (:139-164)
    template = iofile.check_atoms_order(path, lambda *a, **k: iofile.read_VASP_XDATCAR(*a, pinned=False, **k), st)
    dyn = iofile.read_VASP_XDATCAR(path, st, time_step=0.001, template=template, pinned=False)
    assert np.abs(np.asarray(dyn.trajectory)[0] - pos).max() < 1e-8


def test_dump_to_spectrum_end_to_end(kops, monkeypatch, tmp_path):
    """The whole drop-in path in one go, as a user of the reference would drive it: LAMMPS velocity dump -> native
    reader -> Dynamics -> Quasiparticle (q, eigenvectors, FFT selector) -> phonon power spectra, against the oracle
    fed with the numbers written to the file."""
    from oracle import port
    from dynaphopy_b200 import iofile
    from dynaphopy_b200.quasiparticle import Quasiparticle
    from dynaphopy_b200.parameters import Parameters
    from dynaphopy_b200.structure import Structure
    _patch_default_ops(monkeypatch, kops)
    rng = np.random.default_rng(73)
    cell = np.diag([4.0, 4.1, 3.9])
    st = Structure(cell=cell, positions=np.array([[0.0, 0, 0], [2.0, 2.05, 1.95]]), masses=[28.0855, 69.723])
    dims = (2, 2, 2)
    nt, dt = 128, 0.002
    v = rng.normal(size=(nt, 16, 3))
    sc = np.diag(dims) @ cell
    bounds = ["0.0 %.8f" % sc[0, 0], "0.0 %.8f" % sc[1, 1], "0.0 %.8f" % sc[2, 2]]
    path = str(tmp_path / "vel.lammpstrj")
    _write_dump(path, list(v), [100 + i for i in range(nt)], bounds, "vx vy vz", rng=rng)
    v_file = np.array([[[float("%.10f" % x) for x in row] for row in fr] for fr in v])
    Parameters().reset()
    dyn = iofile.read_lammps_trajectory(path, st, time_step=dt, pinned=False, ops=kops)
    assert dyn.get_time_step_average() == dt and np.array_equal(dyn.get_supercell_matrix(), dims)
    calc = Quasiparticle(dyn)
    calc.parameters.use_symmetry = False
    calc.parameters.frequency_range = np.arange(0, 40.05, 2.5)
    calc.select_power_spectra_algorithm(2)
    q_red = np.array([0.5, 0.0, 0.5])
    calc.set_reduced_q_vector(q_red)
    eig = np.linalg.qr(rng.normal(size=(6, 6)) + 1j * rng.normal(size=(6, 6)))[0].T.reshape(6, 2, 3)
    calc.set_eigenvectors(eig)
    ps = calc.get_power_spectrum_phonon()
    pos = st.get_positions(supercell=dims)
    typ = st.get_atom_type_index(supercell=dims)
    vm = port.velocity_mass_average(v_file, st.get_masses(supercell=dims))
    vq = port.project_onto_phonon(port.project_onto_wave_vector(vm, pos, typ, 2, calc.get_q_vector()), eig)
    ref = port.fft_spectra(vq, dt, calc.parameters.frequency_range)
    assert relerr(ps, ref) < TOL
    assert np.array_equal(ps.argmax(0), ref.argmax(0))


def test_poscar_structure_reader(tmp_path):
    from dynaphopy_b200 import iofile
    cell = np.array([[5.4661639157319968, 0, 0], [0, 5.4661639157319968, 0], [0, 0, 5.4661639157319968]])
    frac = np.array([[.875, .875, .875], [.875, .375, .375], [.125, .125, .125]])
    new = str(tmp_path / "POSCAR")
    with open(new, "w") as f:
        f.write(" Si O\n   1.5\n" + "".join("  %.16f %.16f %.16f\n" % tuple(r) for r in cell) + "   Si O\n   2 1\nDirect\n")
        f.write("".join("  %.16f  %.16f  %.16f\n" % tuple(r) for r in frac))
    st = iofile.read_from_file_structure_poscar(new, masses={"Si": 28.0855, "O": 15.9994})
    assert np.array_equal(st.get_cell(), cell * 1.5)
    assert np.array_equal(st.get_positions(), np.dot(frac, cell * 1.5))
    assert list(st.get_masses()) == [28.0855, 28.0855, 15.9994] and st.get_atomic_elements() == ["Si", "Si", "O"]
    old = str(tmp_path / "POSCAR_old")                                   # old style: elements on the title line, Cartesian
    with open(old, "w") as f:
        f.write(" Si O\n   1.0\n" + "".join("  %.16f %.16f %.16f\n" % tuple(r) for r in cell) + "   2 1\nCartesian\n")
        f.write("".join("  %.16f  %.16f  %.16f\n" % tuple(r) for r in frac @ cell))
    st2 = iofile.read_from_file_structure_poscar(old, masses=[28.0855, 28.0855, 15.9994])
    assert np.array_equal(st2.get_positions(), np.array([[float("%.16f" % v) for v in r] for r in frac @ cell]))
    assert st2.get_atomic_elements() == ["Si", "Si", "O"]
    with pytest.raises(SystemExit):
        iofile.read_from_file_structure_poscar(str(tmp_path / "nope"), masses={})
    with pytest.raises(ValueError, match="masses"):
        iofile.read_from_file_structure_poscar(new)


@pytest.mark.skipif(not os.path.isfile("/root/reference/unittest/Si_data/POSCAR"), reason="reference data not mounted")
def test_poscar_reader_matches_reference_structure():
    from dynaphopy_b200 import iofile
    g = golden("reading_test")
    st = iofile.read_from_file_structure_poscar("/root/reference/unittest/Si_data/POSCAR", masses={"Si": 28.0855})
    assert np.array_equal(st.get_cell(), g["unit_cell"]) and np.array_equal(st.get_positions(), g["unit_pos"])
    assert np.array_equal(st.get_masses(), g["unit_masses"])
    st.set_primitive_matrix(g["prim_matrix"])
    assert np.array_equal(st.get_atom_type_index(), g["unit_atom_type"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/unittest/Si_data"), reason="reference data not mounted")
@pytest.mark.parametrize("fname,dt,tag", [("XDATCAR", 0.0005, "xdatcar"), ("si.lammpstrj", 0.001, "lammps")])
def test_reading_test_of_the_reference_line_by_line(kops, fname, dt, tag):
    """unittest/reading_test.py:11-102 rewritten against this package with the same calls in the same order:
    structure from POSCAR, parser picked by keyword, atom-order template, trajectory with the same cuts, and the
    reference's own assertions (time step, mean displacement matrix literals; the rel_traj literals are compared with
    what the reference itself computed, they are differences of O(10) numbers at the 1e-9 level)."""
    from dynaphopy_b200 import iofile as io
    data = "/root/reference/unittest/Si_data/"
    structure = io.read_from_file_structure_poscar(data + "POSCAR", masses={"Si": 28.0855})
    structure.set_primitive_matrix([[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]])
    parser = io.get_trajectory_parser(data + fname)
    kw = dict(pinned=False, ops=kops)
    if tag == "xdatcar":
        template = io.check_atoms_order(data + fname, lambda *a, **k: parser(*a, pinned=False, **k), structure)
        trajectory = parser(data + fname, structure, initial_cut=3, end_cut=14, time_step=dt, template=template, **kw)
    else:
        trajectory = parser(data + fname, structure, initial_cut=3, end_cut=14, time_step=dt, **kw)
    g = golden("reading_test")
    rel_traj = np.array([np.average(trajectory.average_positions().real - traj, axis=0).real
                         for traj in np.asarray(trajectory.trajectory)])
    assert np.abs(rel_traj - g[tag + "_rel_traj"]).max() < 1e-13
    assert float(trajectory.get_time_step_average()) == float(dt)
    mean_matrix = np.average(trajectory.get_mean_displacement_matrix(), axis=0)
    assert np.allclose(mean_matrix, g[tag + "_mean_matrix_literal"], rtol=1e-3, atol=1.e-6)
    assert np.allclose(mean_matrix, g[tag + "_mean_matrix"], rtol=1e-12, atol=0)
