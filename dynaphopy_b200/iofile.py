"""Trajectory ingest with the reference's parser interface (dynaphopy/interface/iofile/trajectory_parsers.py), the text
conversion done natively by all host threads (``dp_lammps_scan`` / ``dp_lammps_read``) into ONE (T, N, ndim) float64
buffer -- pinned when a CUDA device is present, so that ``MeshPipeline.run_from_host`` / the ``Dynamics`` device
accessors can stream it without another host copy.

    read_lammps_trajectory(file_name, structure, time_step, limit_number_steps, last_steps, initial_cut, end_cut,
                           memmap, template) -> Dynamics                      (trajectory_parsers.py:135-352)
    read_VASP_XDATCAR(same arguments)            -> Dynamics                  (trajectory_parsers.py:355-483)

Differences from the reference, all of representation only: the array is real float64 (the reference stores
complex128 with zero imaginary part) and ``memmap`` is accepted but unnecessary (no Python list of frames is built).
"""
import ctypes
import os

import numpy as np

from . import _lib
from .dynamics import Dynamics
from .structure import Structure


def _host_buffer(shape, pinned):
    if pinned:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
        except Exception:
            pass
    return np.empty(shape, dtype=np.float64)


def lammps_supercell(bounds, bound_cols, structure, number_of_dimensions=3):
    """Cell of the dump and the rotation onto the unit-cell orientation (trajectory_parsers.py:217-277)."""
    bounds = np.array(bounds, dtype=float).reshape(3, 3)[:, :bound_cols]
    if bounds.shape[1] == 2:
        bounds = np.append(bounds, np.array([0, 0, 0])[None].T, axis=1)
    xy, xz, yz = bounds[0, 2], bounds[1, 2], bounds[2, 2]
    xlo = bounds[0, 0] - np.min([0.0, xy, xz, xy + xz])
    xhi = bounds[0, 1] - np.max([0.0, xy, xz, xy + xz])
    ylo = bounds[1, 0] - np.min([0.0, yz])
    yhi = bounds[1, 1] - np.max([0.0, yz])
    zlo, zhi = bounds[2, 0], bounds[2, 1]
    supercell = np.array([[xhi - xlo, xy, xz], [0, yhi - ylo, yz], [0, 0, zhi - zlo]]).T
    supercell = supercell[:number_of_dimensions, :number_of_dimensions]

    def unit_matrix(matrix):
        return np.array([np.array(row) / np.linalg.norm(row) for row in matrix])

    transformation_mat = np.dot(np.linalg.inv(unit_matrix(structure.get_cell())), unit_matrix(supercell)).T
    return np.dot(supercell, transformation_mat), transformation_mat


def read_lammps_trajectory(file_name, structure=None, time_step=None, limit_number_steps=10000000, last_steps=None,
                           initial_cut=1, end_cut=None, memmap=False, template=None, pinned=True, threads=0, ops=None):
    lib = _lib.load()
    if not os.path.isfile(file_name):                                   # :162-164
        print('Trajectory file does not exist!')
        raise SystemExit
    if time_step is None:                                               # :167-170
        print('Warning! LAMMPS trajectory file does not contain time step information')
        print('Using default: 0.001 ps')
        time_step = 0.001
    print("Reading LAMMPS trajectory")
    number_of_dimensions = 3 if structure is None else structure.get_number_of_dimensions()
    n_frames, n_atoms, cols = ctypes.c_int64(0), ctypes.c_int(0), ctypes.c_int(0)
    bounds = np.zeros(9)
    labels = ctypes.create_string_buffer(512)
    lib.call("dp_lammps_scan", os.fsencode(file_name), ctypes.byref(n_frames), ctypes.byref(n_atoms),
             bounds.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cols), labels, 512)
    n_total, n_atoms = int(n_frames.value), int(n_atoms.value)
    if structure is not None and n_atoms % structure.get_number_of_cell_atoms() != 0:       # :214-216
        print('Warning: Number of atoms not matching, check LAMMPS output file')
    supercell, transformation_mat = lammps_supercell(bounds, int(cols.value), structure, number_of_dimensions)
    # frames are counted from 1; [initial_cut, end_cut] inclusive, and the "security routine" that stops after
    # limit_number_steps + initial_cut + 1 frames (:318-324)
    first = max(int(initial_cut), 1)
    last = n_total
    if end_cut is not None:
        last = min(last, int(end_cut))
    last = min(last, int(limit_number_steps) + int(initial_cut) + 1)
    if int(limit_number_steps) + int(initial_cut) + 1 < n_total and (end_cut is None or end_cut > last):
        print("Warning! maximum number of steps reached! No more steps will be read")
    count = max(last - first + 1, 0)
    data = _host_buffer((count, n_atoms, number_of_dimensions), pinned)
    steps = np.empty(count)
    if count:
        lib.call("dp_lammps_read", os.fsencode(file_name), first - 1, count, n_atoms, number_of_dimensions,
                 data.ctypes.data_as(ctypes.c_void_p), steps.ctypes.data_as(ctypes.c_void_p), int(threads))
    if template is not None:                                            # :304-306
        data[...] = data[:, np.argsort(template), :]
    if not np.array_equal(transformation_mat, np.identity(number_of_dimensions)):
        for i in range(count):                                          # frame by frame like the reference (:309)
            data[i] = np.dot(data[i], transformation_mat)
    time = steps * time_step
    if last_steps is not None and not memmap:                           # :326-332
        data = data[-last_steps:, :, :]
        time = time[-last_steps:]
    lammps_labels = labels.value
    if b'vx vy' in lammps_labels:                                       # :335-349
        return Dynamics(structure=structure, velocity=data, time=time, supercell=supercell, memmap=memmap, ops=ops)
    if b'x y' in lammps_labels:
        return Dynamics(structure=structure, trajectory=data, time=time, supercell=supercell, memmap=memmap, ops=ops)
    print('LAMMPS parsing error. Data not recognized: {}'.format(lammps_labels))
    raise SystemExit


def read_VASP_XDATCAR(file_name, structure=None, time_step=None, limit_number_steps=10000000, last_steps=None,
                      initial_cut=1, end_cut=None, memmap=False, template=None, pinned=True, threads=0, ops=None):
    lib = _lib.load()
    if not os.path.isfile(file_name):                                   # :381-383
        print('Trajectory file does not exist!')
        raise SystemExit
    if time_step is None:                                               # :386-389
        print('Warning! XDATCAR file does not contain time step information')
        print('Using default: 0.001 ps')
        time_step = 0.001
    print("Reading XDATCAR file")
    n_frames, n_atoms = ctypes.c_int64(0), ctypes.c_int(0)
    super_cell = np.zeros((3, 3))
    lib.call("dp_xdatcar_scan", os.fsencode(file_name), ctypes.byref(n_frames), ctypes.byref(n_atoms),
             super_cell.ctypes.data_as(ctypes.c_void_p))
    n_total, n_atoms = int(n_frames.value), int(n_atoms.value)
    first = max(int(initial_cut), 1)
    last = n_total
    if end_cut is not None:
        last = min(last, int(end_cut))
    last = min(last, int(limit_number_steps) + int(initial_cut) + 1)
    if int(limit_number_steps) + int(initial_cut) + 1 < n_total and (end_cut is None or end_cut > last):
        print("Warning! maximum number of steps reached! No more steps will be read")
    count = max(last - first + 1, 0)
    data = _host_buffer((count, n_atoms, 3), pinned)
    # the reference appends the time of EVERY frame it passes, skipped ones included (:418 is before the cut at :429)
    steps = np.empty(max(last, 0))
    if count or steps.size:
        lib.call("dp_xdatcar_read", os.fsencode(file_name), first - 1, count, n_atoms,
                 data.ctypes.data_as(ctypes.c_void_p), steps.ctypes.data_as(ctypes.c_void_p), steps.size, int(threads))
    if template is not None:                                            # :440-442
        data[...] = data[:, np.argsort(template), :]
    time = steps * time_step
    if last_steps is not None and not memmap:                           # :472-474
        data = data[-last_steps:, :, :]
        time = time[-last_steps:]
    return Dynamics(structure=structure, scaled_trajectory=data, time=time, supercell=super_cell, memmap=memmap, ops=ops)


# -- atom order of a trajectory file vs the order the kernels assume (interface/iofile/__init__.py:23-128) ------------
def get_correct_arrangement(reference, structure):
    """Template such that ``frame[np.argsort(template)]`` is in DynaPhoPy's supercell order
    (i = u*N_cells + x + Nx*(y + Ny*z), atoms.py:139-156) -- same result as the reference's O(N^2) search
    (nearest unit-cell atom with its folded metric, then nearest lattice point), computed per atom in O(N)."""
    cell = np.asarray(structure.get_cell(), dtype=float)
    scaled = np.dot(np.asarray(reference).real, np.linalg.inv(cell))
    n_cell_atoms = structure.get_number_of_atoms()
    n_super = scaled.shape[0]
    dims = np.array(np.round(np.max(scaled, axis=0)), dtype=int)
    unit_scaled = scaled - np.array(scaled, dtype=int)                        # truncation, like the reference (:56)
    sp = np.dot(np.asarray(structure.get_positions(), dtype=float), np.linalg.inv(cell))
    diff = np.abs(unit_scaled[:, None, :] - sp[None, :, :])                   # :62-66
    diff[diff >= 0.5] -= 1.0
    diff[diff < -0.5] += 1.0
    unit_index = np.argmin(np.linalg.norm(diff, axis=2), axis=1)
    lp = scaled - sp[unit_index]                                              # :91-98
    for k in range(3):
        hi = lp[:, k] > dims[k] - 0.5
        lp[hi, k] -= dims[k]
        lo = lp[:, k] < -0.5
        lp[lo, k] += dims[k]
    xyz = np.clip(np.round(lp).astype(int), 0, dims - 1)                      # nearest point of dynaphopy_order (:80,:100-101)
    lattice = xyz[:, 0] + dims[0] * (xyz[:, 1] + dims[1] * xyz[:, 2])
    template = lattice + unit_index * n_super / n_cell_atoms
    if len(np.unique(template)) < len(template):                              # :122-125
        print('template failed, auto-order will not be applied')
        print('unique: {} / {}'.format(len(np.unique(template)), len(template)))
        return range(len(template))
    return template


def check_atoms_order(filename, trajectory_reading_function, structure):
    """interface/iofile/__init__.py:23-40: read the first frame and derive the template from it."""
    trajectory = trajectory_reading_function(filename, structure=structure, initial_cut=0, end_cut=1)
    return get_correct_arrangement(np.asarray(trajectory.trajectory)[0], structure)


def get_trajectory_parser(file_name, bytes_to_check=1000000):
    """interface/iofile/__init__.py:139-164: pick the reader from the keywords found at the head of the file (the
    deprecated OUTCAR reader of the reference is not provided: such files return None here)."""
    parsers_keywords = {'lammps_dump': {'function': read_lammps_trajectory,
                                        'keywords': ['ITEM: TIMESTEP', 'ITEM: NUMBER OF ATOMS', 'ITEM: BOX BOUNDS']},
                        'vasp_xdatcar': {'function': read_VASP_XDATCAR,
                                         'keywords': ['Direct configuration', 'Direct configuration', '=']}}
    if not os.path.isfile(file_name):
        print(file_name + ' file does not exist')
        raise SystemExit
    with open(file_name, "rb") as f:
        head = f.read(int(bytes_to_check))
    for parser in parsers_keywords.values():
        if all(head.find(keyword.encode()) >= 0 for keyword in parser['keywords']):
            return parser['function']
    return None


def read_from_file_structure_poscar(file_name, number_of_dimensions=3, masses=None):
    """interface/iofile/__init__.py:268-328 (new and old style POSCAR, Direct or Cartesian coordinates).  The
    reference looks the atomic masses up in its own periodic table (atoms.py:383-); this package carries no such table:
    pass ``masses`` as ``{element: mass}`` or as one value per atom."""
    if not os.path.isfile(file_name):
        print('Structure file does not exist!')
        raise SystemExit
    print("Reading VASP POSCAR structure")
    with open(file_name, 'r') as poscar_file:
        data_lines = poscar_file.read().split('\n')
    multiply = float(data_lines[1])
    direct_cell = np.array([data_lines[i].split() for i in range(2, 2 + number_of_dimensions)], dtype=float)
    direct_cell *= multiply
    scaled_positions = None
    positions = None
    try:
        number_of_types = np.array(data_lines[3 + number_of_dimensions].split(), dtype=int)
        first, names = 8, data_lines[5].split()
        coordinates_type = data_lines[4 + number_of_dimensions][0]
    except ValueError:                                                  # old style: no element line (:306-321)
        print("Reading old style POSCAR")
        number_of_types = np.array(data_lines[5].split(), dtype=int)
        first, names = 7, data_lines[0].split()
        coordinates_type = data_lines[6][0]
    coords = np.array([data_lines[first + k].split()[0:3] for k in range(np.sum(number_of_types))], dtype=float)
    if coordinates_type in ('D', 'd'):
        scaled_positions = coords
    else:
        positions = coords
    atomic_types = [name for i, name in enumerate(names[:len(number_of_types)]) for _ in range(number_of_types[i])]
    if masses is None:
        raise ValueError("atomic masses are not stored in a POSCAR: pass masses={element: mass} (or one mass per atom)")
    if isinstance(masses, dict):
        masses = [masses[e] for e in atomic_types]
    return Structure(cell=direct_cell, scaled_positions=scaled_positions, positions=positions, masses=masses,
                     atomic_elements=atomic_types)
