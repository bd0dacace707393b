"""TEST INFRASTRUCTURE -- CPU restatement ("port") of the reference hot path.

numpy for the parts the reference itself writes in numpy, ctypes over the plain-C
restatement (``oracle/c_port/oracle_port.c`` -> ``oracle/_build/liboracle_port.so``) for
the parts the reference writes in C.  Citations are into ``/root/reference``.
Never imported by ``dynaphopy_b200``.
"""
import ctypes
import os
import subprocess

import numpy as np

UNIT_CONVERSION = 0.00010585723   # dynaphopy/power_spectrum/__init__.py:9

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_port.so")
_lib = None


def build_c_port():
    """(Re)build the C restatement; building the checker is not using it."""
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def _c():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build_c_port()
        lib = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        lib.op_burg_part.restype = ctypes.c_double
        lib.op_burg_part.argtypes = [dp, ctypes.c_int, ctypes.c_int, dp]
        lib.op_mem_spectrum.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, ctypes.c_double,
                                        ctypes.c_int, dp]
        lib.op_correlation.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, ctypes.c_double,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
        lib.op_displacements.argtypes = [dp, ctypes.c_int, dp, dp, dp]
        lib.op_cell_inverse.argtypes = [dp, dp, dp]
        lib.op_gradient.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp]
        lib.op_project_q.argtypes = [dp, ctypes.c_int, ctypes.c_int, dp, ip, ctypes.c_int, dp,
                                     ctypes.c_int, dp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ----------------------------------------------------------------------------------
# structure helpers (dynaphopy/atoms.py:139-156, 201-208, 285-317)
# ----------------------------------------------------------------------------------
def supercell_positions(unit_positions, unit_cell, dims):
    """Atom-major replication, x fastest: i = k_unit*Nc + x + Nx*(y + Ny*z) (atoms.py:150-153)."""
    unit_positions = np.asarray(unit_positions, float)
    unit_cell = np.asarray(unit_cell, float)
    out = []
    for k in range(unit_positions.shape[0]):
        for z in range(dims[2]):
            for y in range(dims[1]):
                for x in range(dims[0]):
                    out.append(unit_positions[k] + np.dot(np.array([x, y, z]), unit_cell))
    return np.array(out)


def supercell_replicate(per_unit_atom, dims):
    """masses / type index replicated per unit atom (atoms.py:201-208, 314-317)."""
    return np.repeat(np.asarray(per_unit_atom), int(np.prod(dims)))


def q_cartesian(q_reduced, primitive_cell):
    """Quasiparticle.get_q_vector (dynaphopy/__init__.py:167-169)."""
    return np.dot(np.asarray(q_reduced, float), 2.0 * np.pi * np.linalg.inv(primitive_cell).T)


# ----------------------------------------------------------------------------------
# A1  mass weighting (dynamics.py:143-156)
# ----------------------------------------------------------------------------------
def velocity_mass_average(velocity, masses):
    return velocity * np.sqrt(np.asarray(masses, float))[None, :, None]


# ----------------------------------------------------------------------------------
# A2/A3  projections (projection.py:4-54)
# ----------------------------------------------------------------------------------
def project_onto_wave_vector(vm, positions, atom_type, n_prim, q_vector, project_on_atom=-1):
    """vc[t,p,k] = sum_{i:type(i)=p} vm[t,i,k] exp(-i q.r_i) / sqrt(N/n_p)  (projection.py:24-39).

    Accumulates atom by atom in reference order so rounding follows the reference."""
    vm = np.asarray(vm)
    nt, natoms, ndim = vm.shape
    out = np.zeros((nt, n_prim, ndim), dtype=complex)
    for i in range(natoms):
        if project_on_atom > -1 and atom_type[i] != project_on_atom:
            continue
        phase = np.exp(-1j * np.dot(q_vector, positions[i, :]))
        for k in range(ndim):
            out[:, atom_type[i], k] += vm[:, i, k] * phase
    out /= np.sqrt(natoms / n_prim)
    return out


def project_onto_phonon(vc, eigenvectors):
    """vq[t,s] = sum_{p} dot(vc[t,p,:], conj(e[s,p,:]))  (projection.py:43-54)."""
    out = np.zeros((vc.shape[0], eigenvectors.shape[0]), dtype=complex)
    for s in range(eigenvectors.shape[0]):
        for p in range(vc.shape[1]):
            out[:, s] += np.dot(vc[:, p, :], eigenvectors[s, p, :].conj())
    return out


def project_onto_wave_vector_c(vm_real, positions, atom_type, n_prim, q_vector, project_on_atom=-1):
    """Same sum through the C restatement (real input only); used for CPU-baseline timing."""
    vm_real = np.ascontiguousarray(vm_real, dtype=np.float64)
    nt, natoms, _ = vm_real.shape
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    typ = np.ascontiguousarray(atom_type, dtype=np.int32)
    q = np.ascontiguousarray(q_vector, dtype=np.float64)
    out = np.empty((nt, n_prim, 3), dtype=np.complex128)
    _c().op_project_q(_p(vm_real), nt, natoms, _p(pos), typ.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                      n_prim, _p(q), int(project_on_atom), _p(out.view(np.float64)))
    return out


# ----------------------------------------------------------------------------------
# A7  FFT power spectrum (power_spectrum/__init__.py:33-51, 226-270)
# ----------------------------------------------------------------------------------
def division_of_data(resolution, number_of_data, time_step):
    """power_spectrum/__init__.py:33-51 -- note npieces+1 windows and the duplicated
    single-window case."""
    piece_size = round(1. / (time_step * resolution))
    number_of_pieces = int((number_of_data - 1) / piece_size)
    if number_of_pieces > 0:
        interval = (number_of_data - piece_size) / number_of_pieces
    else:
        interval = 0
        number_of_pieces = 1
        piece_size = number_of_data
    pieces = []
    for i in range(number_of_pieces + 1):
        ini = int((piece_size / 2 + i * interval) - piece_size / 2)
        fin = int((piece_size / 2 + i * interval) + piece_size / 2)
        pieces.append([ini, fin])
    return pieces


def autocorrelation_same(x):
    """np.correlate(x, x, 'same') by zero-padded FFT: a[m] = sum_n x[n+k] conj(x[n]), k = m - L//2.
    Restatement used for long pieces (the direct np.correlate is O(L^2))."""
    L = x.size
    P = 1
    while P < L + L // 2 + 1:
        P *= 2
    X = np.fft.fft(x, P)
    r = np.fft.ifft(X * np.conj(X))
    k = np.arange(L) - L // 2
    return r[k % P]


def fft_power(frequency_range, data, time_step, use_fft_correlate=False):
    """_numpy_power (power_spectrum/__init__.py:226-245)."""
    pieces = division_of_data(frequency_range[1] - frequency_range[0], data.size, time_step)
    ps = []
    for ini, fin in pieces:
        piece = data[ini:fin]
        if use_fft_correlate:
            acf = autocorrelation_same(piece) / piece.size
        else:
            acf = np.correlate(piece, piece, mode='same') / piece.size
        ps.append(np.abs(np.fft.fft(acf)) * time_step)
    ps = np.average(ps, axis=0)
    freqs = np.fft.fftfreq(piece.size, time_step)
    idx = np.argsort(freqs)
    return np.interp(frequency_range, freqs[idx], ps[idx])


def fft_spectra(vq, time_step, frequency_range, use_fft_correlate=False):
    """get_fft_numpy_spectra (power_spectrum/__init__.py:248-270): (T,M) -> (F,M), eV*ps."""
    frequency_range = np.array(frequency_range)
    cols = [fft_power(frequency_range, vq[:, i], time_step, use_fft_correlate)
            for i in range(vq.shape[1])]
    return np.array(cols).T * UNIT_CONVERSION


# ----------------------------------------------------------------------------------
# A6  MEM (power_spectrum/__init__.py:81-103, c/mem.c) -- canonical Data[N] := 0
# ----------------------------------------------------------------------------------
def burg_part(x, m):
    x = np.ascontiguousarray(x, dtype=np.float64)
    d = np.zeros(m, dtype=np.float64)
    xms = _c().op_burg_part(_p(x), x.size, int(m), _p(d))
    return xms, d


def mem_power(frequency_range, data, time_step, coefficients):
    f = np.ascontiguousarray(frequency_range, dtype=np.float64)
    v = np.ascontiguousarray(data, dtype=np.complex128)
    out = np.empty(f.size, dtype=np.float64)
    _c().op_mem_spectrum(_p(f), f.size, _p(v.view(np.float64)), v.size, float(time_step),
                         int(coefficients), _p(out))
    return out


def mem_spectra(vq, time_step, frequency_range, coefficients):
    """get_mem_power_spectra (power_spectrum/__init__.py:81-103)."""
    cols = [mem_power(frequency_range, vq[:, i], time_step, coefficients) for i in range(vq.shape[1])]
    return np.nan_to_num(np.array(cols).T) * UNIT_CONVERSION


# ----------------------------------------------------------------------------------
# A5  direct correlation (power_spectrum/__init__.py:57-75, c/correlation.c)
# ----------------------------------------------------------------------------------
def correlation_power(frequency_range, data, time_step, step=10, integration_method=1, weight="reference"):
    f = np.ascontiguousarray(frequency_range, dtype=np.float64)
    v = np.ascontiguousarray(data, dtype=np.complex128)
    out = np.empty(f.size, dtype=np.float64)
    _c().op_correlation(_p(f), f.size, _p(v.view(np.float64)), v.size, float(time_step), int(step),
                        int(integration_method), 0 if weight == "reference" else 1, _p(out))
    return out


def correlation_spectra(vq, time_step, frequency_range, step=10, integration_method=1, weight="reference"):
    cols = [correlation_power(frequency_range, vq[:, i], time_step, step, integration_method, weight)
            for i in range(vq.shape[1])]
    with np.errstate(all="ignore"):
        return np.array(cols).T * UNIT_CONVERSION


def correlation_lag_sums(data, step):
    """C(i) = sum_{j<N-i-step} conj(v_j) v_{j+i}, i = 0, step, ... (frequency independent part)."""
    v = np.asarray(data, dtype=np.complex128)
    n = v.size
    lags = np.arange(0, n, step)
    out = np.zeros(lags.size, dtype=np.complex128)
    for a, i in enumerate(lags):
        cnt = n - i - step
        if cnt > 0:
            out[a] = np.sum(np.conj(v[:cnt]) * v[i:i + cnt])
    return out


# ----------------------------------------------------------------------------------
# A8/A9  displacements + velocity (c/displacements.c, dynamics.py:158-179, 313-329)
# ----------------------------------------------------------------------------------
def atomic_displacements(trajectory_one_atom, position, cell):
    tr = np.ascontiguousarray(np.real(trajectory_one_atom), dtype=np.float64)
    pos = np.ascontiguousarray(position, dtype=np.float64)
    c = np.ascontiguousarray(cell, dtype=np.float64)
    out = np.empty_like(tr)
    _c().op_displacements(_p(tr), tr.shape[0], _p(pos), _p(c), _p(out))
    return out


def cell_inverse(cell):
    c = np.ascontiguousarray(cell, dtype=np.float64)
    cm = np.empty((3, 3)); ci = np.empty((3, 3))
    _c().op_cell_inverse(_p(c), _p(cm), _p(ci))
    return cm, ci


def relative_trajectory(trajectory, positions, cell):
    """Dynamics.get_relative_trajectory (dynamics.py:158-179) on a real (T,N,3) array."""
    out = np.empty(trajectory.shape, dtype=np.float64)
    for i in range(trajectory.shape[1]):
        out[:, i, :] = atomic_displacements(trajectory[:, i, :], positions[i], cell)
    return out


def mean_displacement_matrix(rel, atom_type_supercell, atom_type_unit, n_cells, use_average_positions=False):
    """Dynamics.get_mean_displacement_matrix (dynamics.py:207-239) on a (T,N,3) relative trajectory."""
    rel = np.asarray(rel)
    per_type = np.unique(atom_type_unit, return_counts=True)[1]
    ntypes = len(per_type)
    diff = np.average(rel, axis=0) if use_average_positions else np.zeros(rel.shape[1:])
    mats = np.zeros((ntypes, 3, 3))
    for i in range(rel.shape[1]):
        mats[atom_type_supercell[i]] += np.dot(np.conj(rel[:, i, :]).T, rel[:, i, :] - diff[i]).real \
            / per_type[atom_type_supercell[i]]
    return mats / (n_cells * rel.shape[0])


def velocity_from_relative(rel, time_step):
    """Dynamics.velocity (dynamics.py:313-329): np.gradient along time for every (atom, dim)."""
    return np.gradient(rel, time_step, axis=0)


def time_step_average(time):
    """Dynamics.get_time_step_average (dynamics.py:129-138), same accumulation order."""
    acc = 0
    n = len(time)
    for i in range(n - 1):
        acc += (time[i + 1] - time[i]) / (n - 1)
    return np.round(acc, decimals=8)


# ----------------------------------------------------------------------------------
# (f)4  LAMMPS dump parsing (interface/iofile/trajectory_parsers.py:135-352), line by line like the reference
# ----------------------------------------------------------------------------------
def read_lammps_frames(file_name, number_of_dimensions=3, initial_cut=1, end_cut=None):
    """(frames (T,N,ndim), TIMESTEP values (T,), bounds rows, labels) with the reference's marker logic."""
    with open(file_name, "rb") as f:
        lines = f.read().split(b"\n")
    frames, steps, bounds, labels, natoms = [], [], None, None, None
    counter, i = 0, 0
    while True:
        while i < len(lines) and b"TIMESTEP" not in lines[i]:
            i += 1
        if i >= len(lines):
            break
        counter += 1
        step = float(lines[i + 1])
        if natoms is None:
            j = next(k for k, l in enumerate(lines) if b"NUMBER OF ATOMS" in l)
            natoms = int(lines[j + 1])
        if bounds is None:
            j = next(k for k, l in enumerate(lines) if b"BOX BOUNDS" in l)
            bounds = [[float(x) for x in lines[j + 1 + r].split()] for r in range(3)]
        while i < len(lines) and b"ITEM: ATOMS" not in lines[i]:
            i += 1
        if i >= len(lines):
            break
        labels = lines[i]
        i += 1
        if initial_cut > counter:
            continue
        frames.append([[float(x) for x in lines[i + a].split()[0:number_of_dimensions]] for a in range(natoms)])
        steps.append(step)
        i += natoms
        if end_cut is not None and end_cut <= counter:
            break
    return np.array(frames, dtype=float), np.array(steps), bounds, labels


def get_correct_arrangement(reference, cell, scaled_positions):
    """interface/iofile/__init__.py:43-128, the reference's O(N^2) search restated step by step."""
    scaled_coordinates = [np.array(np.dot(c, np.linalg.inv(cell)).real, dtype=float) for c in reference]
    number_of_cell_atoms = len(scaled_positions)
    number_of_supercell_atoms = len(scaled_coordinates)
    supercell_dim = np.array(np.round(np.max(scaled_coordinates, axis=0)), dtype=int)
    unit = scaled_coordinates - np.array(scaled_coordinates, dtype=int)
    atom_unit_cell_index = []
    for coordinate in unit:
        diff = np.abs(np.array([coordinate] * number_of_cell_atoms) - scaled_positions)
        diff[diff >= 0.5] -= 1.0
        diff[diff < -0.5] += 1.0
        atom_unit_cell_index.append(np.argmin(np.linalg.norm(diff, axis=1)))
    atom_unit_cell_index = np.array(atom_unit_cell_index)

    def order(i, size):
        return np.array([np.mod(i, size[0]), np.mod(i, size[0] * size[1]) // size[0],
                         np.mod(i, size[0] * size[1] * size[2]) // (size[1] * size[0])])
    original_conf = np.array([order(j, supercell_dim) for j in range(number_of_supercell_atoms)])
    template = []
    for i, coordinate in enumerate(scaled_coordinates):
        lp = coordinate - scaled_positions[atom_unit_cell_index[i]]
        for k in range(3):
            if lp[k] > supercell_dim[k] - 0.5:
                lp[k] = lp[k] - supercell_dim[k]
            if lp[k] < -0.5:
                lp[k] = lp[k] + supercell_dim[k]
        comparison = np.array([lp] * number_of_supercell_atoms)
        dm = comparison / supercell_dim[None, :].astype(float) - original_conf / supercell_dim[None, :].astype(float)
        template.append(np.argmin(np.linalg.norm(dm, axis=1)) +
                        atom_unit_cell_index[i] * number_of_supercell_atoms / number_of_cell_atoms)
    return np.array(template)
