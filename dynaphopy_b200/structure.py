"""Minimal crystal-structure holder with the part of the reference ``Structure`` interface that
the hot path reads (dynaphopy/atoms.py).  The reference's own ``dynaphopy.atoms.Structure`` can be
passed to ``Dynamics`` instead (duck typing); this class exists so the path runs without the
reference package (e.g. on the GPU box).  Pure set-up code, O(N) once -- not a kernel.
"""
import numpy as np


class Structure:
    def __init__(self, cell, positions=None, scaled_positions=None, masses=None, atomic_numbers=None,
                 atomic_elements=None, primitive_matrix=None, atom_type_index=None):
        self._cell = np.array(cell, dtype=float)                       # lattice vectors in rows
        if positions is None:
            if scaled_positions is None:
                raise ValueError("No positions provided")
            positions = np.dot(np.array(scaled_positions, dtype=float), self._cell)
        self._positions = np.array(positions, dtype=float)
        self._masses = np.array(masses, dtype=float)
        self._atomic_numbers = None if atomic_numbers is None else np.array(atomic_numbers)
        self._atomic_elements = atomic_elements
        self._primitive_matrix = None if primitive_matrix is None else np.array(primitive_matrix, dtype=float)
        self._atom_type_index = None if atom_type_index is None else np.array(atom_type_index, dtype=int)

    # -- cell -------------------------------------------------------------------------------
    def get_cell(self):
        return self._cell

    def get_number_of_dimensions(self):
        return self._cell.shape[1]

    def get_primitive_matrix(self):
        if self._primitive_matrix is None:                              # atoms.py:127-134
            self._primitive_matrix = np.identity(self.get_number_of_dimensions())
        return self._primitive_matrix

    def set_primitive_matrix(self, primitive_matrix):
        self._primitive_matrix = np.array(primitive_matrix, dtype=float)
        self._atom_type_index = None

    def get_primitive_cell(self):
        return np.dot(self.get_cell().T, self.get_primitive_matrix()).T   # atoms.py:113-116

    # -- atoms ------------------------------------------------------------------------------
    def get_number_of_cell_atoms(self):
        return self._positions.shape[0]

    get_number_of_atoms = get_number_of_cell_atoms

    def get_positions(self, supercell=None):
        """Atom-major replication, x fastest: i = u*Nc + x + Nx*(y + Ny*z) (atoms.py:139-156)."""
        if supercell is None:
            return self._positions
        sc = [int(s) for s in supercell]
        grid = np.array([[x, y, z] for z in range(sc[2]) for y in range(sc[1]) for x in range(sc[0])], dtype=float)
        shifts = grid @ self._cell
        return (self._positions[:, None, :] + shifts[None, :, :]).reshape(-1, self._cell.shape[1])

    def get_masses(self, supercell=None):
        reps = 1 if supercell is None else int(np.prod(supercell))
        return np.repeat(self._masses, reps)                            # atoms.py:201-208

    def get_atom_type_index(self, supercell=None):
        """Unit-cell atoms that differ by a primitive lattice vector share a type (atoms.py:285-317)."""
        if self._atom_type_index is None:
            n = self.get_number_of_cell_atoms()
            inv = np.linalg.inv(self.get_primitive_cell())
            types = [None] * n
            counter = 0
            for i in range(n):
                if types[i] is None:
                    types[i] = counter
                    counter += 1
                for j in range(i + 1, n):
                    d = self._positions[j] - self._positions[i]
                    shift = np.around(np.dot(inv.T, d))
                    proj = self._positions[j] - np.dot(self.get_primitive_cell().T, shift)
                    if np.linalg.norm(proj - self._positions[i]) ** 2 < 0.001 and self._masses[i] == self._masses[j]:
                        types[j] = types[i]
            self._atom_type_index = np.array(types, dtype=int)
        reps = 1 if supercell is None else int(np.prod(supercell))
        return np.repeat(self._atom_type_index, reps)

    def get_number_of_atom_types(self):
        return len(set(self.get_atom_type_index().tolist()))

    get_number_of_primitive_atoms = get_number_of_atom_types

    def get_atomic_elements(self, supercell=None, unique=False):
        el = self._atomic_elements if self._atomic_elements is not None else ["X%d" % t for t in self.get_atom_type_index()]
        if unique:
            seen, out = set(), []
            for t, e in zip(self.get_atom_type_index(), el):
                if t not in seen:
                    seen.add(t)
                    out.append(e)
            return out
        reps = 1 if supercell is None else int(np.prod(supercell))
        return [e for e in el for _ in range(reps)]

    def get_commensurate_points(self, supercell):
        """Reduced q (primitive reciprocal basis) commensurate with a diagonal supercell:
        q = (m / dims) . P, m in the supercell grid, kept distinct modulo the primitive reciprocal
        lattice (inverse of Quasiparticle.check_commensurate, dynaphopy/__init__.py:610-623).
        Plain lattice arithmetic -- phonopy is only needed by the reference for symmetry reduction."""
        sc = [int(s) for s in supercell]
        pm = self.get_primitive_matrix()
        seen, out = set(), []
        for mz in range(sc[2]):
            for my in range(sc[1]):
                for mx in range(sc[0]):
                    q = np.dot(np.array([mx / sc[0], my / sc[1], mz / sc[2]]), pm)
                    q = q - np.floor(q + 1e-9)
                    key = tuple(np.round(q, 6) % 1.0)
                    if key not in seen:
                        seen.add(key)
                        out.append(q)
        return np.array(out)
