"""``Quasiparticle`` -- the callers of the hot path, with the reference's method names and caching contract
(dynaphopy/__init__.py:23-186, 610-854) for everything that does not need phonopy: q handling, projections,
power spectra of phonon modes / wave vector / all atoms.  Eigenvectors come from phonopy in the reference
(`get_eigenvectors`, :172-186); here they are injected with ``set_eigenvectors`` (same (M, n_p, 3) layout,
``phonopy_link.py:189-192``).  Peak fitting, renormalised force constants, plotting and I/O stay in the
reference.
"""
import numpy as np

from . import projection
from .parameters import Parameters
from .power_spectrum import power_spectrum_functions


class Quasiparticle:
    def __init__(self, dynamic, last_steps=None, vc=None):
        self._dynamic = dynamic
        self._vc = vc
        self._vq = None
        self._eigenvectors = None
        self._frequencies = None
        self._power_spectrum_phonon = None
        self._power_spectrum_wave_vector = None
        self._power_spectrum_direct = None
        self._power_spectrum_partials = None
        self._parameters = Parameters()
        self._harmonic_provider = None             # q_reduced -> (eigenvectors (M, n_p, 3), frequencies (M,))
        self._star_provider = None                 # q_reduced -> list of symmetry-equivalent reduced q points
        self.crop_trajectory(last_steps)

    # -- state -----------------------------------------------------------------------------------------
    @property
    def dynamic(self):
        return self._dynamic

    @property
    def parameters(self):
        return self._parameters

    def crop_trajectory(self, last_steps):
        if last_steps is not None:
            self._dynamic.crop_trajectory(last_steps)
        print("Using {0} steps".format(len(self._dynamic.get_time())))

    def power_spectra_clear(self):                       # :47-51
        self._power_spectrum_phonon = None
        self._power_spectrum_wave_vector = None
        self._power_spectrum_direct = None
        self._power_spectrum_partials = None

    def full_clear(self):                                # :58-65
        self._eigenvectors = None
        self._frequencies = None
        self._vc = None
        self._vq = None
        self.power_spectra_clear()

    # -- parameters ------------------------------------------------------------------------------------
    def set_number_of_mem_coefficients(self, coefficients):                  # :128-130
        self.power_spectra_clear()
        self.parameters.number_of_coefficients_mem = coefficients

    def set_projection_onto_atom_type(self, atom_type):                      # :132-137
        if atom_type in range(self.dynamic.structure.get_number_of_primitive_atoms()):
            self.parameters.project_on_atom = atom_type
        else:
            print('Atom type {} does not exist'.format(atom_type))
            raise SystemExit

    def set_frequency_limits(self, limits):                                  # :139-142
        self.power_spectra_clear()
        self.parameters.frequency_range = np.arange(limits[0], limits[1] + self.parameters.spectrum_resolution,
                                                    self.parameters.spectrum_resolution)

    def set_spectra_resolution(self, resolution):                            # :144-150
        self.power_spectra_clear()
        limits = [self.get_frequency_range()[0], self.get_frequency_range()[-1]]
        self.parameters.spectrum_resolution = resolution
        self.parameters.frequency_range = np.arange(limits[0], limits[1] + resolution, resolution)

    def get_frequency_range(self):
        return self.parameters.frequency_range

    def set_reduced_q_vector(self, q_vector):                                # :157-162
        q_vector = np.array(q_vector, dtype=float)
        if not np.array_equal(q_vector, np.asarray(self.parameters.reduced_q_vector, dtype=float)):
            self.full_clear()
        self.parameters.reduced_q_vector = q_vector

    def get_reduced_q_vector(self):
        return self.parameters.reduced_q_vector

    def get_q_vector(self):                                                  # :167-169
        return np.dot(self.parameters.reduced_q_vector,
                      2.0 * np.pi * np.linalg.inv(self.dynamic.structure.get_primitive_cell()).T)

    def check_commensurate(self, q_point, decimals=4):                       # :610-623
        supercell = self.dynamic.get_supercell_matrix()
        primitive_matrix = self.dynamic.structure.get_primitive_matrix()
        transform = np.dot(q_point, np.linalg.inv(primitive_matrix)) * supercell
        transform = np.round(transform, decimals=decimals)
        return bool(np.all(np.equal(np.mod(transform, 1), 0)))

    def select_power_spectra_algorithm(self, algorithm):                     # :697-707
        if algorithm in power_spectrum_functions.keys():
            if algorithm != self.parameters.power_spectra_algorithm:
                self.power_spectra_clear()
                self.parameters.power_spectra_algorithm = algorithm
            print("Using {0} function".format(power_spectrum_functions[algorithm][1]))
        else:
            print("Power spectrum algorithm number not found!\nPlease select:")
            for i in power_spectrum_functions.keys():
                print('{0} : {1}'.format(i, power_spectrum_functions[i][1]))
            raise SystemExit

    def get_algorithm_list(self):                                            # :1097-1098
        return power_spectrum_functions.values()

    # -- harmonic inputs (phonopy in the reference) -------------------------------------------------------
    def set_eigenvectors(self, eigenvectors, frequencies=None):
        """(M, n_p, 3) complex eigenvectors of the current q, arranged as phonopy_link.py:189-192."""
        self._eigenvectors = np.asarray(eigenvectors, dtype=complex)
        self._frequencies = frequencies
        self._vq = None
        self._power_spectrum_phonon = None

    def set_harmonic_provider(self, provider):
        """Callable ``q_reduced -> (eigenvectors (M, n_p, 3), frequencies (M,))`` standing in for
        ``pho_interface.obtain_eigenvectors_and_frequencies`` (dynaphopy/__init__.py:172-186): with it the eigenvectors
        follow ``set_reduced_q_vector`` lazily, exactly as in the reference."""
        self._harmonic_provider = provider

    def set_star_provider(self, provider):
        """Callable ``q_reduced -> [q_reduced, ...]`` standing in for
        ``pho_interface.get_equivalent_q_points_by_symmetry`` (used at dynaphopy/__init__.py:727-728)."""
        self._star_provider = provider

    def _harmonic(self):
        if self._eigenvectors is None and self._harmonic_provider is not None:
            e, f = self._harmonic_provider(np.array(self.parameters.reduced_q_vector, dtype=float))
            self._eigenvectors, self._frequencies = np.asarray(e, dtype=complex), (None if f is None else np.asarray(f))

    def get_eigenvectors(self):
        self._harmonic()
        if self._eigenvectors is None:
            raise RuntimeError("eigenvectors come from phonopy in the reference (dynaphopy/__init__.py:172-186); "
                               "inject them with set_eigenvectors()")
        return self._eigenvectors

    def get_frequencies(self):
        self._harmonic()
        return self._frequencies

    # -- projections -----------------------------------------------------------------------------------------
    def get_vc(self):                                                        # :626-640
        if self._vc is None:
            print("Projecting into wave vector")
            if not self.check_commensurate(self.get_reduced_q_vector()):
                print("warning! This wave vector is not a commensurate q-point in MD supercell")
            self._vc = projection.project_onto_wave_vector(self.dynamic, self.get_q_vector(),
                                                           project_on_atom=self.parameters.project_on_atom)
        return self._vc

    def get_vq(self):                                                        # :642-646
        if self._vq is None:
            print("Projecting into phonon mode")
            self._vq = projection.project_onto_phonon(self.get_vc(), self.get_eigenvectors())
        return self._vq

    # -- power spectra -----------------------------------------------------------------------------------------
    def _spectrum(self, columns):
        fn = power_spectrum_functions[self.parameters.power_spectra_algorithm][0]
        return fn(columns, self.dynamic, self.parameters)

    def get_power_spectrum_phonon(self):                                     # :721-746
        if self._power_spectrum_phonon is None:
            print("Calculating phonon projection power spectra")
            if self.parameters.use_symmetry and self._star_provider is not None and self._harmonic_provider is not None:
                # star of q (:725-739).  The reference loops over the equivalent points (projection, contraction and
                # spectra one q at a time); here the whole star is projected in one pass over the trajectory and all
                # n_star * M spectra come from one launch.  np.average over the members in the reference's order.
                initial = np.array(self.get_reduced_q_vector(), dtype=float)
                star = [np.array(q, dtype=float) for q in self._star_provider(initial)]
                to_cart = 2.0 * np.pi * np.linalg.inv(self.dynamic.structure.get_primitive_cell()).T
                vc = projection.project_onto_wave_vectors(self.dynamic, np.array([np.dot(q, to_cart) for q in star]),
                                                          project_on_atom=self.parameters.project_on_atom)
                eig = np.array([np.asarray(self._harmonic_provider(q)[0], dtype=complex) for q in star])
                ops = self.dynamic.ops
                vq = ops.project_phonon(vc, eig)                                         # (T, n_star, M) on the device
                nt, ns, m = ops.be.shape(vq)
                ps = self._spectrum(vq.reshape(nt, ns * m))                              # (F, n_star * M)
                self._power_spectrum_phonon = np.average([ps[:, i * m:(i + 1) * m] for i in range(ns)], axis=0)
            else:
                self._power_spectrum_phonon = self._spectrum(self.get_vq())
        return self._power_spectrum_phonon

    def get_power_spectrum_wave_vector(self):                                # :748-779
        if self._power_spectrum_wave_vector is None:
            print('Calculating wave vector projection power spectrum')
            vc = self.get_vc()
            size = vc.shape[1] * vc.shape[2]
            self._power_spectrum_wave_vector = self._spectrum(vc.swapaxes(1, 2).reshape(-1, size))
        return np.nansum(self._power_spectrum_wave_vector, axis=1)

    def _velocity_mass_average(self):
        """(T, N, 3) mass-weighted velocities: the device tensor when the Dynamics mirror has one (no host round
        trip), else the reference-typed numpy array."""
        if hasattr(self.dynamic, "velocity_mass_average_device"):
            return self.dynamic.velocity_mass_average_device()
        return self.dynamic.get_velocity_mass_average()

    def get_power_spectrum_full(self, projection_on_coordinate=-1):          # :781-831
        """All-atom power spectrum, branch for branch like the reference: ``parameters.project_on_atom >= 0`` keeps
        the atoms of that primitive type ('Atom type ... does not exist' -> exit), and only then is
        ``projection_on_coordinate`` honoured (>= number of dimensions -> exit); ``parameters.silent`` accumulates
        the columns one by one in (atom, coordinate) order, otherwise the (F, columns) array of the
        coordinate-major layout ``vm.swapaxes(1, 2).reshape(-1, size)`` is summed over its columns.  Either way the
        spectra of ALL columns come from one launch.  Cached until ``power_spectra_clear`` whatever the arguments,
        as in the reference."""
        number_of_dimensions = self.dynamic.structure.get_number_of_dimensions()
        projected_atom_type = self.parameters.project_on_atom
        if self._power_spectrum_direct is None:
            print("Calculation full power spectrum")
            vm = self._velocity_mass_average()
            if projected_atom_type >= 0:
                print('Power spectrum projected onto atom type {0}'.format(projected_atom_type))
                supercell = self.dynamic.get_supercell_matrix()
                atom_types = np.array(self.dynamic.structure.get_atom_type_index(supercell=supercell))
                atom_indices = np.argwhere(atom_types == projected_atom_type).flatten()
                if len(atom_indices) == 0:
                    print('Atom type {0} does not exist'.format(projected_atom_type))
                    raise SystemExit
                # Only works if project on atom is requested! (reference comment, :803)
                if projection_on_coordinate >= number_of_dimensions:
                    print('Projected coordinate should be smaller than {}'.format(number_of_dimensions))
                    raise SystemExit
                idx = atom_indices if isinstance(vm, np.ndarray) else \
                    self.dynamic.ops.be.asarray(atom_indices, np.int32).long()
                if projection_on_coordinate > -1:
                    print('Power spectrum projected onto coordinate {}'.format(projection_on_coordinate))
                    vm = vm[:, idx, projection_on_coordinate][:, :, None]
                else:
                    vm = vm[:, idx]
            n_atoms, n_coord = vm.shape[1], vm.shape[2]
            size = n_atoms * n_coord
            cols = vm.swapaxes(1, 2).reshape(-1, size) if isinstance(vm, np.ndarray) \
                else vm.transpose(1, 2).reshape(-1, size)
            ps = self._spectrum(cols)                                        # (F, size), column j * n_atoms + i
            if self.parameters.silent:
                # "Memory efficient algorithm" of the reference (:812-821): same columns, accumulated one at a time
                acc = np.zeros_like(np.asarray(self.parameters.frequency_range, dtype=float)[None].T)
                for i in range(n_atoms):
                    for j in range(n_coord):
                        acc += ps[:, j * n_atoms + i][None].T
                ps = acc
            self._power_spectrum_direct = np.sum(ps, axis=1)
        return self._power_spectrum_direct

    def get_power_spectrum_partials(self, save_to_file=None):                # :833-854
        if self._power_spectrum_partials is None:
            print("Calculation power spectrum partials")
            vm = self._velocity_mass_average()
            self._power_spectrum_partials = self._spectrum(vm[:, :, 0])
            for i in [1, 2]:
                self._power_spectrum_partials += self._spectrum(vm[:, :, i])
        if save_to_file is not None:
            np.savetxt(save_to_file, np.hstack([self.get_frequency_range()[None].T, self._power_spectrum_partials]))
        return self._power_spectrum_partials

    # -- whole commensurate mesh in one pass (the q loop of get_commensurate_points_data, :1130-1174) --------
    def get_commensurate_points_spectra(self, eigenvectors, frequencies=None, com_points=None, use_degeneracy=None,
                                        cutoff=1e-4):
        """Power spectra of EVERY commensurate q point at once, ready for the fitting stage.

        The reference loops over the commensurate points one q at a time (set_reduced_q_vector -> get_vc -> get_vq ->
        get_power_spectrum_phonon -> phonon_fitting_analysis).  Here all q are projected from one pass over the
        trajectory (``MeshPipeline``), the spectra of all Q*M modes are computed by the selected algorithm, and the
        degenerate-set average plus the np.max / np.argmax initial guess of ``phonon_fitting_analysis``
        (analysis/fitting/__init__.py:36-61) are taken on the device.  Lorentzian fitting stays in scipy.

        eigenvectors: (Q, M, n_p, 3) in the arrangement of ``get_eigenvectors`` for every q of ``com_points``
        frequencies:  (Q, M) harmonic frequencies (needed for the degenerate sets), or None
        com_points:   (Q, 3) reduced q; default: all commensurate points of the MD supercell
        Returns a dict: q_points (Q,3), power_spectra (Q,F,M), averaged (Q,F,M) or None, peak_bin (Q,M),
        guess_position (Q,M), guess_height (Q,M), degenerate_sets (list per q)."""
        from . import pipeline
        from .fitting import degenerate_sets
        from .ops import to_host
        dyn = self.dynamic
        st = dyn.structure
        sc = dyn.get_supercell_matrix()
        if com_points is None:
            com_points = st.get_commensurate_points(sc)
        com_points = np.atleast_2d(np.asarray(com_points, dtype=float))
        q_cart = np.array([np.dot(q, 2.0 * np.pi * np.linalg.inv(st.get_primitive_cell()).T) for q in com_points])   # :167-169
        n_prim = st.get_number_of_primitive_atoms()
        M = 3 * n_prim
        Q = len(com_points)
        eig = np.asarray(eigenvectors, dtype=complex).reshape(Q, M, M)
        plan = pipeline.CommensuratePlan(st.get_cell(), sc, dyn.positions_supercell(), st.get_masses(supercell=sc),
                                         st.get_atom_type_index(supercell=sc), n_prim, q_cart)
        if not (plan.usable() and plan.commensurate.all()):
            raise ValueError("the batched path needs q points commensurate with a diagonal MD supercell")
        if use_degeneracy is None:
            use_degeneracy = bool(self.parameters.use_symmetry) and frequencies is not None
        freq = np.asarray(self.get_frequency_range(), dtype=float)
        pipe = pipeline.MeshPipeline(plan, eig, freq, dyn.get_time_step_average(), ops=dyn.ops,
                                     algorithm=self.parameters.power_spectra_algorithm,
                                     mem_coefficients=self.parameters.number_of_coefficients_mem,
                                     correlation_step=self.parameters.correlation_function_step,
                                     integration_method=self.parameters.integration_method,
                                     correlation_weight=1 if getattr(self.parameters, "correlation_weight",
                                                                     "reference") == "intended" else 0,
                                     project_on_atom=self.parameters.project_on_atom)
        ps = pipe.run(dyn.velocity_device)                                  # (F, Q*M) on the device
        sets_q, sets_flat = [], None
        if use_degeneracy:
            fr = np.asarray(frequencies, dtype=float).reshape(Q, M)
            sets_q = [degenerate_sets(fr[i], cutoff=cutoff) for i in range(Q)]
            sets_flat = [[i * M + j for j in g] for i, sq in enumerate(sets_q) for g in sq]
        avg, bins, heights = dyn.ops.spectra_peaks(ps, sets_flat)
        F = freq.size
        bins_h = to_host(bins).reshape(Q, M)
        out = {
            "q_points": com_points,
            "power_spectra": to_host(ps).reshape(F, Q, M).transpose(1, 0, 2),
            "averaged": None if avg is None else to_host(avg).reshape(F, Q, M).transpose(1, 0, 2),
            "peak_bin": bins_h,
            "guess_position": freq[bins_h],
            "guess_height": to_host(heights).reshape(Q, M),
            "degenerate_sets": sets_q,
        }
        return out
