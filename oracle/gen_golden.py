"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz from the REFERENCE ITSELF.

Run in the dev container only (needs /root/reference and `make -C oracle ref`):

    python -m oracle.gen_golden

Every array written here is either a seeded input or the output of the reference's own,
unmodified code (Python imported in place via oracle.ref_python, C kernels compiled by
oracle/Makefile).  MEM uses the deterministic build (Data[N] := 0);
direct correlation is recorded for the as-shipped build and for the intended-weight
build (oracle/shims/fixcexp.h).  The fixtures travel to the GPU box; the reference does not.
"""
import os
import sys

import numpy as np

from . import ref, ref_python

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Params:
    """Stand-in for dynaphopy.parameters.Parameters: only the fields the selector reads."""

    def __init__(self, frequency_range, mem=100, step=10, method=1):
        self.frequency_range = frequency_range
        self.silent = True
        self.zero_padding = 0
        self.number_of_coefficients_mem = mem
        self.correlation_function_step = step
        self.integration_method = method


def _unitary_eigenvectors(rng, n_prim):
    m = 3 * n_prim
    u = np.linalg.qr(rng.normal(size=(m, m)) + 1j * rng.normal(size=(m, m)))[0]
    return u.T.reshape(m, n_prim, 3)          # (mode, prim_atom, xyz), phonopy_link.py:189-192


def hotpath_case(name, unit_cell, unit_pos, masses, numbers, prim_matrix, dims, nt, dt, q_reduced,
                 freq, mem_order, seed, complex_velocity=False):
    from dynaphopy.atoms import Structure
    from dynaphopy.dynamics import Dynamics
    from dynaphopy import projection, power_spectrum

    rng = np.random.default_rng(seed)
    st = Structure(cell=unit_cell, positions=unit_pos, masses=masses, atomic_numbers=numbers,
                   primitive_matrix=prim_matrix)
    n_unit = len(masses)
    natoms = n_unit * int(np.prod(dims))
    supercell = np.dot(np.diagflat(dims), unit_cell)
    t = np.arange(nt) * dt
    vel = rng.normal(size=(nt, natoms, 3))
    for f0, a0 in ((3.83, 1.0), (7.11, 0.7), (15.53, 0.4)):
        vel += a0 * np.cos(2 * np.pi * f0 * t[:, None, None] + rng.uniform(0, 2 * np.pi, size=(1, natoms, 3)))
    vel = vel.astype(complex)
    if complex_velocity:
        vel = vel + 1j * rng.normal(size=vel.shape)
    dyn = Dynamics(structure=st, velocity=vel.copy(), time=t, supercell=supercell)
    n_prim = st.get_number_of_primitive_atoms()
    pos = st.get_positions(dyn.get_supercell_matrix())
    typ = np.array(st.get_atom_type_index(supercell=dyn.get_supercell_matrix()), dtype=np.int32)
    mass_sc = np.array(st.get_masses(supercell=dyn.get_supercell_matrix()))
    prim_cell = st.get_primitive_cell()

    q_cart, vcs, vqs, eigs = [], [], [], []
    for q in q_reduced:
        qc = np.dot(np.array(q, float), 2.0 * np.pi * np.linalg.inv(prim_cell).T)   # __init__.py:167-169
        vc = projection.project_onto_wave_vector(dyn, qc)
        e = _unitary_eigenvectors(rng, n_prim)
        vq = projection.project_onto_phonon(vc, e)
        q_cart.append(qc); vcs.append(vc); vqs.append(vq); eigs.append(e)
    vc_atom0 = projection.project_onto_wave_vector(dyn, q_cart[1], project_on_atom=0)

    p = _Params(freq, mem=mem_order)
    vq0 = vqs[1]
    fns = power_spectrum.power_spectrum_functions
    out = dict(
        unit_cell=unit_cell, unit_pos=unit_pos, unit_masses=np.array(masses, float),
        prim_matrix=np.array(prim_matrix, float), dims=np.array(dims), dt=dt, time=t,
        velocity=vel if complex_velocity else vel.real, positions=pos, atom_type=typ, masses=mass_sc,
        supercell=supercell, primitive_cell=prim_cell, n_prim=n_prim,
        q_reduced=np.array(q_reduced, float), q_cart=np.array(q_cart), eigenvectors=np.array(eigs),
        vc=np.array(vcs), vq=np.array(vqs), vc_atom0=vc_atom0, vm=dyn.get_velocity_mass_average(),
        freq=freq, mem_order=mem_order, time_step_average=dyn.get_time_step_average(),
        ps_fft=fns[2][0](vq0, dyn, p), ps_mem=fns[1][0](vq0, dyn, p),
        ps_corr_m1=fns[0][0](vq0, dyn, p),
    )
    p0 = _Params(freq, mem=mem_order, method=0)
    out["ps_corr_m0"] = fns[0][0](vq0, dyn, p0)
    ci = ref.correlation_intended()
    out["ps_corr_intended_m1"] = np.array(
        [ci.correlation_par(freq, np.ascontiguousarray(vq0[:, i]), dyn.get_time_step_average(),
                            step=10, integration_method=1) for i in range(vq0.shape[1])]).T
    out["ps_corr_intended_m0"] = np.array(
        [ci.correlation_par(freq, np.ascontiguousarray(vq0[:, i]), dyn.get_time_step_average(),
                            step=10, integration_method=0) for i in range(vq0.shape[1])]).T
    # wave-vector spectrum input layout (dynaphopy/__init__.py:764): vc.swapaxes(1,2).reshape(-1, 3 n_p)
    out["ps_fft_wave_vector_cols"] = fns[2][0](vcs[1].swapaxes(1, 2).reshape(-1, 3 * n_prim), dyn, p)
    out.update(_full_and_partial_spectra(dyn, freq, mem_order))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, {k: np.shape(v) for k, v in out.items() if np.ndim(v) > 1})


def _full_and_partial_spectra(dyn, freq, mem_order):
    """Quasiparticle.get_power_spectrum_full / get_power_spectrum_partials of the reference itself
    (dynaphopy/__init__.py:781-854) for every branch: all atoms / one atom type / one atom type and one
    coordinate, the `silent` column-by-column path and the single-call path, FFT and MEM methods."""
    import dynaphopy
    pr = dynaphopy.parameters.Parameters()
    saved = (pr.silent, pr.project_on_atom, pr.power_spectra_algorithm, pr.frequency_range,
             pr.number_of_coefficients_mem)
    res = {}
    try:
        pr.frequency_range = freq
        pr.number_of_coefficients_mem = mem_order

        def full(alg, atom, coord, silent):
            pr.power_spectra_algorithm = alg
            pr.project_on_atom = atom
            pr.silent = silent
            calc = dynaphopy.Quasiparticle(dyn)
            return np.array(calc.get_power_spectrum_full(projection_on_coordinate=coord))

        res["ps_full_fft"] = full(2, -1, -1, False)
        res["ps_full_fft_silent"] = full(2, -1, -1, True)
        res["ps_full_fft_coord1_ignored"] = full(2, -1, 1, False)       # coordinate is honoured only with an atom type
        res["ps_full_fft_atom0"] = full(2, 0, -1, False)
        res["ps_full_fft_atom1_silent"] = full(2, 1, -1, True)
        res["ps_full_fft_atom0_coord1"] = full(2, 0, 1, False)
        res["ps_full_fft_atom1_coord2_silent"] = full(2, 1, 2, True)
        res["ps_full_mem"] = full(1, -1, -1, False)
        res["ps_full_mem_atom0_coord0"] = full(1, 0, 0, False)
        pr.project_on_atom = -1
        pr.silent = False
        for alg, key in ((2, "ps_partials_fft"), (1, "ps_partials_mem")):
            pr.power_spectra_algorithm = alg
            res[key] = np.array(dynaphopy.Quasiparticle(dyn).get_power_spectrum_partials())
    finally:
        pr.silent, pr.project_on_atom, pr.power_spectra_algorithm, pr.frequency_range, \
            pr.number_of_coefficients_mem = saved
    return res


def displacement_case(name, seed):
    """positions -> c/displacements.c -> np.gradient -> mass weighting, through Dynamics."""
    from dynaphopy.atoms import Structure
    from dynaphopy.dynamics import Dynamics
    rng = np.random.default_rng(seed)
    unit_cell = np.array([[4.1, 0.0, 0.0], [0.6, 3.9, 0.0], [0.3, -0.5, 4.4]])
    unit_pos = np.dot(np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]]), unit_cell)
    masses = [28.0855, 69.723]
    dims = (3, 2, 2)
    st = Structure(cell=unit_cell, positions=unit_pos, masses=masses, atomic_numbers=[14, 31],
                   primitive_matrix=np.eye(3))
    natoms = 2 * int(np.prod(dims))
    nt, dt = 96, 0.002
    supercell = np.dot(np.diagflat(dims), unit_cell)
    dyn0 = Dynamics(structure=st, trajectory=np.zeros((2, natoms, 3), complex), time=np.arange(2) * dt,
                    supercell=supercell)
    pos = st.get_positions(dyn0.get_supercell_matrix())
    traj = pos[None] + rng.normal(scale=0.05, size=(nt, natoms, 3))
    wrap = rng.random(size=(nt, natoms)) < 0.05
    shifts = rng.integers(-2, 3, size=(nt, natoms, 3))
    traj = traj + (wrap[..., None] * shifts) @ supercell
    t = np.arange(nt) * dt
    dyn = Dynamics(structure=st, trajectory=traj.astype(complex), time=t, supercell=supercell)
    rel = dyn.get_relative_trajectory()
    vel = dyn.velocity
    vm = dyn.get_velocity_mass_average()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), unit_cell=unit_cell, supercell=supercell,
                        positions=pos, trajectory=traj, time=t, dt=dyn.get_time_step_average(),
                        masses=np.array(st.get_masses(supercell=dyn.get_supercell_matrix())),
                        relative=rel.real, velocity=vel.real, vm=vm.real,
                        rel_imag_max=np.abs(rel.imag).max(), vel_imag_max=np.abs(vel.imag).max())
    print("wrote", name, rel.shape)


def reading_test_case():
    """Known-answer fixtures of unittest/reading_test.py:23-102 (the only reference test that
    isolates c/displacements.c + Dynamics).  Inputs parsed by the reference's own parsers."""
    import dynaphopy.interface.iofile as io
    cwd = os.getcwd()
    os.chdir(os.path.join(ref_python.REFERENCE_ROOT, "unittest"))
    try:
        st = io.read_from_file_structure_poscar('Si_data/POSCAR')
        st.set_primitive_matrix([[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]])
        out = {}
        for tag, fname, ts in (("xdatcar", 'Si_data/XDATCAR', 0.0005), ("lammps", 'Si_data/si.lammpstrj', 0.001)):
            parser = io.get_trajectory_parser(fname)
            if tag == "xdatcar":
                template = io.check_atoms_order(fname, parser, st)
                dyn = parser(fname, st, initial_cut=3, end_cut=14, time_step=ts, template=template)
                out["xdatcar_template"] = np.array(template)
            else:
                dyn = parser(fname, st, initial_cut=3, end_cut=14, time_step=ts)
            sc = dyn.get_supercell_matrix()
            out[tag + "_trajectory"] = np.array(dyn.trajectory).real
            out[tag + "_supercell"] = np.array(dyn.get_supercell())
            out[tag + "_positions"] = st.get_positions(sc)
            out[tag + "_atom_type"] = np.array(st.get_atom_type_index(supercell=sc), dtype=np.int32)
            out[tag + "_relative"] = np.array(dyn.get_relative_trajectory()).real
            out[tag + "_mean_matrix"] = np.average(dyn.get_mean_displacement_matrix(), axis=0)
            out[tag + "_time_step"] = dyn.get_time_step_average()
            out[tag + "_time"] = np.array(dyn.get_time())
            # dynamics.py:207-295 in full: per-type matrices, the average-position variant, average positions
            out[tag + "_mean_matrix_types"] = np.array(dyn.get_mean_displacement_matrix())
            dyn._mean_displacement_matrix = None
            out[tag + "_mean_matrix_avgpos"] = np.array(dyn.get_mean_displacement_matrix(use_average_positions=True))
            dyn._mean_displacement_matrix = None
            out[tag + "_average_positions"] = np.array(dyn.average_positions())
            out[tag + "_average_positions_unit"] = np.array(dyn.average_positions(to_unit_cell=True))
            # the quantity reading_test.py:44-45, 84-85 compares with its rel_traj_ref literals
            out[tag + "_rel_traj"] = np.array([np.average(dyn.average_positions().real - traj, axis=0).real
                                               for traj in dyn.trajectory])
        out["unit_cell"] = np.array(st.get_cell())
        out["unit_pos"] = np.array(st.get_positions())
        out["unit_masses"] = np.array(st.get_masses())
        out["prim_matrix"] = np.array(st.get_primitive_matrix())
        out["unit_atom_type"] = np.array(st.get_atom_type_index(), dtype=np.int32)
        # literals asserted by the reference test (reading_test.py:54-56, 94-96), rtol=1e-3, atol=1e-6
        out["xdatcar_mean_matrix_literal"] = np.array([[0.00065253, 0.00016701, 0.00033313],
                                                       [0.00016701, 0.00063413, 0.00023646],
                                                       [0.00033313, 0.00023646, 0.00072747]])
        out["lammps_mean_matrix_literal"] = np.array([[0.00263396, 0.00053198, 0.00040298],
                                                      [0.00053198, 0.00383308, 0.00034559],
                                                      [0.00040298, 0.00034559, 0.00381229]])
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "reading_test.npz"), **out)
    print("wrote reading_test", {k: np.shape(v) for k, v in out.items()})


def main():
    if not (ref_python.available() and ref.available()):
        sys.exit("needs /root/reference and `make -C oracle ref`")
    os.makedirs(OUT, exist_ok=True)
    ref_python.load()
    if "--reading-only" in sys.argv:
        reading_test_case()
        return
    a = 4.0
    # identity-primitive 2-atom cubic cell, 2x3x2 supercell; commensurate q incl. Gamma, zone
    # boundary, a q outside the first zone, and one non-commensurate q
    hotpath_case("cubic2_232", np.diag([a, a, a]), np.array([[0, 0, 0], [a / 2, a / 2, a / 2]]),
                 [28.0855, 69.723], [14, 31], np.eye(3), (2, 3, 2), nt=256, dt=0.002,
                 q_reduced=[[0, 0, 0], [0.5, 1 / 3, 0], [0.5, 2 / 3, 0.5], [1.5, -1 / 3, 0.5], [0.13, 0.27, 0.41]],
                 freq=np.arange(0, 40.05, 5.0), mem_order=20, seed=11)
    # same cell, longer run, default grid spacing 0.05 -> one window (L = T), complex velocities
    hotpath_case("cubic2_222_complex", np.diag([a, a, a]), np.array([[0, 0, 0], [a / 2, a / 2, a / 2]]),
                 [28.0855, 69.723], [14, 31], np.eye(3), (2, 2, 2), nt=300, dt=0.002,
                 q_reduced=[[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0]],
                 freq=np.arange(0, 25.05, 0.05), mem_order=50, seed=12, complex_velocity=True)
    # Si-like conventional fcc cell (8 atoms, primitive matrix of unittest/Si_test.py:21-23),
    # 2x2x2 supercell = the cfg-1 geometry (64 atoms, 2 primitive atoms, 4 unit atoms per type)
    a = 5.45
    frac = np.array([[.875, .875, .875], [.875, .375, .375], [.375, .875, .375], [.375, .375, .875],
                     [.125, .125, .125], [.125, .625, .625], [.625, .125, .625], [.625, .625, .125]])
    hotpath_case("si_conv_222", np.diag([a, a, a]), frac * a, [28.0855] * 8, [14] * 8,
                 [[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]], (2, 2, 2), nt=400, dt=0.002,
                 q_reduced=[[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0], [0.25, 0.25, 0.5]],
                 freq=np.arange(0, 20.05, 2.5), mem_order=30, seed=13)
    displacement_case("displacements_triclinic", seed=21)
    reading_test_case()


if __name__ == "__main__":
    main()
