"""``Dynamics`` with the reference's interface (dynaphopy/dynamics.py) on device-resident arrays.

Same constructor arguments, same method / property names, same return shapes and dtypes
((T, N, 3) complex128 numpy arrays from the reference-named accessors); the work behind
``get_relative_trajectory`` / ``velocity`` / ``get_velocity_mass_average`` is done by the CUDA
kernels of csrc/dp_dynamics.cu in one launch instead of a Python loop over atoms.  ``*_device``
accessors return the fp64 device tensors (real when the input has no imaginary part -- the
reference stores zeros there) for the projection kernels, without a host round trip.
"""
import numpy as np

from . import ops as _ops


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class Dynamics:
    def __init__(self, structure=None, trajectory=None, scaled_trajectory=None, velocity=None, energy=None,
                 time=None, supercell=None, memmap=False, ops=None):
        self._time = time
        self._trajectory = trajectory
        self._scaled_trajectory = scaled_trajectory
        self._energy = energy
        self._velocity = velocity
        self._supercell = None if supercell is None else np.array(supercell, dtype=float)
        self._memmap = memmap                     # accepted for compatibility; HBM replaces the disk spill
        self._structure = structure
        if structure is None:
            print('Warning: Initialization without structure')
        self._ops = ops
        self._time_step_average = None
        self._supercell_matrix = None
        self._number_of_atoms = None
        self._d_velocity = None                   # device caches
        self._d_vm = None
        self._d_relative = None
        self._sqrt_mass = None
        self._moments = None                      # per-atom displacement sums (host, N x 3 and N x 3 x 3)
        self._mean_displacement_matrix = None

    # -- plumbing ---------------------------------------------------------------------------
    @property
    def ops(self):
        if self._ops is None:
            self._ops = _ops.default_ops()
        return self._ops

    def _to_device(self, x):
        be = self.ops.be
        if be.is_complex(x):
            imag_zero = (not bool((x.imag != 0).any()))
            if imag_zero:
                x = x.real
        return be.asarray(x, np.complex128 if be.is_complex(x) else np.float64)

    def _to_host_complex(self, x):
        x = x.detach().cpu().numpy() if _is_torch(x) else np.asarray(x)
        return x.astype(complex)

    # -- reference interface ----------------------------------------------------------------
    @property
    def structure(self):
        return self._structure

    def set_structure(self, structure):
        self._structure = structure

    def get_time(self):
        return self._time

    def set_time(self, time):
        self._time = time
        self._time_step_average = None

    def get_energy(self):
        return self._energy

    def get_supercell(self):
        return self._supercell

    def get_number_of_atoms(self):
        if self._number_of_atoms is None:
            self._number_of_atoms = int(self.structure.get_number_of_atoms() * np.prod(self.get_supercell_matrix()))
        return self._number_of_atoms

    def get_time_step_average(self):
        """Mean of consecutive differences, accumulated in order, rounded to 8 decimals
        (dynamics.py:129-138)."""
        if not self._time_step_average:
            t = np.asarray(self.get_time(), dtype=float)
            n = len(t)
            acc = np.cumsum((t[1:] - t[:-1]) / (n - 1))[-1] if n > 1 else 0
            self._time_step_average = acc
        self._time_step_average = np.round(self._time_step_average, decimals=8)
        return self._time_step_average

    def get_supercell_matrix(self, tolerance=0.01):
        """Three integers from the ratio of cell-vector norms (dynamics.py:181-205)."""
        if self._supercell_matrix is None:
            real = np.divide(np.linalg.norm(self.get_supercell(), axis=1),
                             np.linalg.norm(self.structure.get_cell(), axis=1))
            self._supercell_matrix = np.around(real).astype("int")
            if abs(sum(self._supercell_matrix - real) / np.linalg.norm(real)) > tolerance:
                raise ValueError('Warning! Defined unit cell and MD supercell do not match! '
                                 'Cell size relation is not integer: {0}'.format(real))
            print('MD cell size relation: {0}'.format(self._supercell_matrix))
        return self._supercell_matrix

    def crop_trajectory(self, last_steps):
        """Keep the last ``last_steps`` frames of every array (dynamics.py:65-107)."""
        if last_steps is None or last_steps <= 0:
            return
        for name in ("_trajectory", "_scaled_trajectory", "_energy", "_time", "_velocity", "_d_velocity"):
            arr = getattr(self, name)
            if arr is not None:
                setattr(self, name, arr[-last_steps:])
        if self._d_velocity is None and self._velocity is None:
            _ = self.velocity_device                       # derive before the positions are dropped
            self._d_velocity = self._d_velocity[-last_steps:]
        self._d_vm = None
        self._d_relative = None
        self._time_step_average = None
        self._moments = None
        self._mean_displacement_matrix = None

    @property
    def trajectory(self):
        if self._trajectory is None:
            if self._scaled_trajectory is not None:
                self._trajectory = np.dot(self._scaled_trajectory, self.get_supercell())
            else:
                raise ValueError('No trajectory loaded')
        return self._trajectory

    # -- device accessors (what the kernels consume) -------------------------------------------
    def sqrt_mass_device(self):
        if self._sqrt_mass is None:
            m = np.asarray(self.structure.get_masses(supercell=self.get_supercell_matrix()), dtype=float)
            self._sqrt_mass = self.ops.be.asarray(np.sqrt(m), np.float64)
        return self._sqrt_mass

    def positions_supercell(self):
        return np.asarray(self.structure.get_positions(supercell=self.get_supercell_matrix()), dtype=float)

    def relative_trajectory_device(self):
        if self._d_relative is None:
            self._d_relative = self.ops.atomic_displacements(self._to_device(self.trajectory),
                                                             self.positions_supercell(), self.get_supercell())
        return self._d_relative

    @property
    def velocity_device(self):
        """(T,N,3) device tensor; derived from positions by the fused kernel when not supplied."""
        if self._d_velocity is None:
            if self._velocity is not None:
                self._d_velocity = self._to_device(self._velocity)
            else:
                print('No velocity provided! calculating it from coordinates...')
                self._d_velocity = self.ops.velocity_from_positions(
                    self._to_device(self.trajectory), self.positions_supercell(), self.get_supercell(),
                    self.get_time_step_average())
        return self._d_velocity

    def velocity_mass_average_device(self):
        if self._d_vm is None:
            if self._d_velocity is None and self._velocity is None:
                # positions only: one fused pass positions -> displacements -> gradient -> sqrt(m)
                print('No velocity provided! calculating it from coordinates...')
                self._d_vm = self.ops.velocity_from_positions(
                    self._to_device(self.trajectory), self.positions_supercell(), self.get_supercell(),
                    self.get_time_step_average(), sqrt_mass=self.sqrt_mass_device())
            else:
                self._d_vm = self.ops.mass_weight(self.velocity_device, self.sqrt_mass_device())
        return self._d_vm

    # -- reference-named accessors (numpy complex128, the reference's dtype) ------------------
    @property
    def velocity(self):
        return self._to_host_complex(self.velocity_device)

    @velocity.setter
    def velocity(self, velocity):
        self._velocity = velocity
        self._d_velocity = None
        self._d_vm = None

    def get_velocity_mass_average(self):
        return self._to_host_complex(self.velocity_mass_average_device())

    def get_relative_trajectory(self):
        return self._to_host_complex(self.relative_trajectory_device())

    # -- displacement statistics (dynamics.py:207-295) from ONE fused pass over the positions ---------
    def _displacement_moments(self):
        if self._moments is None:
            from .ops import to_host
            sums = to_host(self.ops.displacement_moments(self._to_device(self.trajectory), self.positions_supercell(),
                                                         self.get_supercell()))
            s1 = sums[:, :3]
            i, j = np.triu_indices(3)
            s2 = np.empty((sums.shape[0], 3, 3))
            s2[:, i, j] = sums[:, 3:]
            s2[:, j, i] = sums[:, 3:]
            self._moments = (s1, s2)
        return self._moments

    def get_mean_displacement_matrix(self, use_average_positions=False):
        """dynamics.py:207-239: per atom type, sum_i sum_t u_t^T (u_t - shift_i) / (n_equiv(type) * N_cells * T)."""
        if self._mean_displacement_matrix is None:
            supercell = self.get_supercell_matrix()
            atom_type = np.asarray(self.structure.get_atom_type_index())
            per_type = np.unique(atom_type, return_counts=True)[1]
            type_index = np.asarray(self.structure.get_atom_type_index(supercell=supercell))
            s1, s2 = self._displacement_moments()
            nt = self.trajectory.shape[0]
            shift = s1 / nt if use_average_positions else np.zeros_like(s1)
            per_atom = s2 - s1[:, :, None] * shift[:, None, :]          # sum_t u_a (u_b - shift_b)
            nd = self.structure.get_number_of_dimensions()
            mats = np.zeros((self.structure.get_number_of_atom_types(), nd, nd))
            for i in range(per_atom.shape[0]):                            # reference accumulation order (:228-232)
                mats[type_index[i]] += per_atom[i] / per_type[type_index[i]]
            self._mean_displacement_matrix = mats / (np.prod(supercell) * nt)
        return self._mean_displacement_matrix

    def average_positions(self, number_of_samples=None, to_unit_cell=False):
        """dynamics.py:241-295 (complex result like the reference; ``number_of_samples`` draws frames at random
        there, which has no deterministic counterpart -- the full average is used)."""
        supercell = self.get_supercell_matrix()
        nd = self.structure.get_number_of_dimensions()
        cell = self.get_supercell()
        positions = self.structure.get_positions(supercell=supercell)
        s1, _ = self._displacement_moments()
        averaged = (s1 / self.trajectory.shape[0]).astype(complex)
        natoms = averaged.shape[0]
        if to_unit_cell:
            cell = self.structure.get_cell()
            natoms = self.structure.get_number_of_atoms()
            positions = self.structure.get_positions()
            type_unit = self.structure.get_atom_type_index()
            type_index = self.structure.get_atom_type_index(supercell=supercell)
            unit = np.zeros((self.structure.get_number_of_atom_types(), nd), dtype=complex)
            norm = np.prod(supercell)
            for i, coordinates in enumerate(averaged):
                unit[type_index[i], :] += coordinates / norm
            averaged = np.array([unit[type_unit[i]] for i in range(positions.shape[0])])
        averaged = averaged + positions
        for j in range(natoms):
            diff = np.around(np.dot(np.linalg.inv(cell.T), averaged[j, :] - 0.5 * np.dot(np.ones(nd), cell)), decimals=0)
            averaged[j, :] -= np.dot(diff, cell)
        return averaged
