"""``projection`` with the reference's operator interface (dynaphopy/projection.py), served by the CUDA
kernels through the C ABI.  Same names, argument meaning and return shapes / dtypes:

    project_onto_wave_vector(trajectory, q_vector, project_on_atom=-1) -> (T, n_p, 3) complex128
    project_onto_phonon(vc, eigenvectors)                               -> (T, M)      complex128

``trajectory`` is a ``Dynamics`` (ours, or any object with the reference's accessors).  A batched form
for many wave vectors (what the reference runs as a Python loop over q) is
``project_onto_wave_vectors``.
"""
import numpy as np

from . import ops as _ops
from . import pipeline as _pipeline


def _ops_of(trajectory):
    return getattr(trajectory, "ops", None) or _ops.default_ops()


def _device_vm(trajectory):
    if hasattr(trajectory, "velocity_mass_average_device"):
        return trajectory.velocity_mass_average_device()
    return trajectory.get_velocity_mass_average()          # reference Dynamics: host complex128 array


def project_onto_wave_vectors(trajectory, q_vectors, project_on_atom=-1, method="auto"):
    """(T, Q, n_p, 3) device array for a list of Cartesian wave vectors: the 3-D lattice-DFT kernel for
    the commensurate ones (every frame read once for all of them), the dense kernel for the rest."""
    ops = _ops_of(trajectory)
    st = trajectory.structure
    supercell = trajectory.get_supercell_matrix()
    coordinates = np.asarray(st.get_positions(supercell), dtype=float)
    atom_type = np.asarray(st.get_atom_type_index(supercell=supercell), dtype=np.int32)
    n_prim = st.get_number_of_primitive_atoms()
    q_vectors = np.atleast_2d(np.asarray(q_vectors, dtype=float))
    if q_vectors.shape[1] != coordinates.shape[1]:
        print("Warning!! Q-vector and coordinates dimension do not match")     # projection.py:19-21
        raise SystemExit
    vm = _device_vm(trajectory)
    masses = np.asarray(st.get_masses(supercell=supercell), dtype=float)
    plan = _pipeline.CommensuratePlan(st.get_cell(), supercell, coordinates, masses, atom_type, n_prim, q_vectors)
    return _pipeline.project_q_list(ops, vm, None, plan, coordinates, atom_type, project_on_atom, method=method,
                                    mass_weighted=True)


def project_onto_wave_vector(trajectory, q_vector, project_on_atom=-1):
    """dynaphopy/projection.py:4-40."""
    vc = project_onto_wave_vectors(trajectory, np.asarray(q_vector, dtype=float)[None, :], project_on_atom)
    return _ops.to_host(vc)[:, 0]


def project_onto_phonon(vc, eigenvectors, ops=None):
    """dynaphopy/projection.py:43-54: vc (T, n_p, 3), eigenvectors (M, n_p, 3) -> (T, M)."""
    ops = ops or _ops.default_ops()
    vc = np.asarray(vc) if not _ops.is_device_array(vc) else vc
    e = np.asarray(eigenvectors, dtype=complex)
    vq = ops.project_phonon(vc[:, None], e[None])
    return _ops.to_host(vq)[:, 0]
