"""Molecular structure container, AMBER prmtop/inpcrd/rst7 + PDB readers and Amber-mask selection.

Stands in for the ``parmed.Structure`` surface that the reference touches:
``load_file(prmtop, xyz=)`` (``blues/settings.py:60-90``), ``.topology.atoms()/.residues()`` with
``.index/.name/.residue.name/.element._mass`` (``blues/moves.py:206-251``), ``.positions``,
``.box_vectors``, ``.velocities``, ``structure[idx_list]``, ``.createSystem(**kw)``
(``blues/simulation.py:219``), ``.save(...)``, ``AmberMask(struct, sel).Selected()``
(``blues/simulation.py:108-137``), ``Rst7`` (``blues/settings.py:79``) and
``geometry.center_of_mass`` (``blues/moves.py:269``).  Everything is backed by flat numpy arrays so
the native engine can consume it without per-atom Python objects.
"""
import math
import re
import numpy as np

from . import unit as u

# standard atomic weights (Da) and symbols, Z = 1..54 plus a few heavier ions
_SYMBOLS = ['X', 'H', 'He', 'Li', 'Be', 'B', 'C', 'N', 'O', 'F', 'Ne', 'Na', 'Mg', 'Al', 'Si', 'P', 'S', 'Cl', 'Ar',
            'K', 'Ca', 'Sc', 'Ti', 'V', 'Cr', 'Mn', 'Fe', 'Co', 'Ni', 'Cu', 'Zn', 'Ga', 'Ge', 'As', 'Se', 'Br', 'Kr',
            'Rb', 'Sr', 'Y', 'Zr', 'Nb', 'Mo', 'Tc', 'Ru', 'Rh', 'Pd', 'Ag', 'Cd', 'In', 'Sn', 'Sb', 'Te', 'I', 'Xe',
            'Cs', 'Ba']
_WEIGHTS = [0.0, 1.007947, 4.003, 6.9412, 9.0121823, 10.8117, 12.01078, 14.00672, 15.99943, 18.99840325, 20.17976,
            22.989769282, 24.30506, 26.98153868, 28.08553, 30.9737622, 32.0655, 35.4532, 39.9481, 39.09831, 40.0784,
            44.9559126, 47.8671, 50.94151, 51.99616, 54.9380455, 55.8452, 58.9331955, 58.69342, 63.5463, 65.4094,
            69.7231, 72.641, 74.921602, 78.963, 79.9041, 83.7982, 85.46783, 87.621, 88.905852, 91.2242, 92.906382,
            95.942, 98.0, 101.072, 102.905502, 106.421, 107.86822, 112.4118, 114.8183, 118.7107, 121.7601, 127.603,
            126.904473, 131.2936, 132.90545192, 137.3277]


class Element(object):
    """Chemical element with ``_mass`` (Quantity, dalton) like ``simtk.openmm.app.Element``."""
    _cache = {}

    def __init__(self, z):
        self.atomic_number = int(z)
        self.symbol = _SYMBOLS[z] if z < len(_SYMBOLS) else 'X'
        self._mass = u.Quantity(_WEIGHTS[z] if z < len(_WEIGHTS) else 0.0, u.dalton)

    mass = property(lambda self: self._mass)

    @classmethod
    def get(cls, z):
        z = int(z)
        if z not in cls._cache:
            cls._cache[z] = Element(z)
        return cls._cache[z]

    @classmethod
    def from_symbol(cls, sym):
        s = sym.strip().capitalize()
        return cls.get(_SYMBOLS.index(s)) if s in _SYMBOLS else cls.get(0)

    def __repr__(self):
        return '<Element %s>' % self.symbol


def element_from_mass(mass):
    if mass <= 0.5:
        return 0
    return int(np.argmin(np.abs(np.asarray(_WEIGHTS[1:]) - mass))) + 1


class Residue(object):
    def __init__(self, struct, index):
        self._s = struct
        self.index = self.idx = index

    name = property(lambda self: self._s.residue_names[self.index])
    number = property(lambda self: self.index + 1)
    id = property(lambda self: str(self.index + 1))

    def atoms(self):
        s = self._s
        return iter(s.atoms[s.residue_pointers[self.index]:s.residue_pointers[self.index + 1]])

    def __len__(self):
        return int(self._s.residue_pointers[self.index + 1] - self._s.residue_pointers[self.index])

    def __iter__(self):
        return self.atoms()

    def __repr__(self):
        return '<Residue %s[%d]>' % (self.name, self.index)


class Atom(object):
    def __init__(self, struct, index):
        self._s = struct
        self.index = self.idx = index

    name = property(lambda self: self._s.atom_names[self.index])
    type = property(lambda self: self._s.atom_types[self.index])
    atomic_number = property(lambda self: int(self._s.atomic_numbers[self.index]))
    mass = property(lambda self: float(self._s.masses[self.index]))
    charge = property(lambda self: float(self._s.charges[self.index]))
    element = property(lambda self: Element.get(self._s.atomic_numbers[self.index]))
    residue = property(lambda self: self._s.residues[self._s.atom_residue[self.index]])
    xx = property(lambda self: float(self._s.coordinates[self.index, 0]))
    xy = property(lambda self: float(self._s.coordinates[self.index, 1]))
    xz = property(lambda self: float(self._s.coordinates[self.index, 2]))
    sigma = property(lambda self: float(self._s.lj_sigma[self.index]))
    epsilon = property(lambda self: float(self._s.lj_epsilon[self.index]))

    @property
    def bond_partners(self):
        return [self._s.atoms[j] for j in self._s.bond_graph()[self.index]]

    def __repr__(self):
        return '<Atom %s [%d]; in %s %d>' % (self.name, self.index, self.residue.name, self.residue.index)


class Topology(object):
    """The ``structure.topology`` view (OpenMM ``Topology``-like)."""

    def __init__(self, struct):
        self._s = struct

    def atoms(self):
        return iter(self._s.atoms)

    def residues(self):
        return iter(self._s.residues)

    def getNumAtoms(self):
        return self._s.n_atoms

    def getNumResidues(self):
        return len(self._s.residues)

    def bonds(self):
        for i, j in self._s.bonds[:, :2]:
            yield self._s.atoms[i], self._s.atoms[j]

    def getPeriodicBoxVectors(self):
        return self._s.box_vectors

    def getUnitCellDimensions(self):
        if self._s.box is None:
            return None
        return u.Quantity(np.asarray(self._s.box[:3]) * 0.1, u.nanometers)


class Structure(object):
    """Flat-array molecular structure.  Lengths are Å, energies kcal/mol, charges e (AMBER conventions);
    conversion to the nm / kJ/mol engine units happens in :meth:`createSystem`."""

    def __init__(self):
        self.n_atoms = 0
        self.atom_names = []
        self.atom_types = []
        self.atomic_numbers = np.zeros(0, int)
        self.masses = np.zeros(0)
        self.charges = np.zeros(0)
        self.lj_sigma = np.zeros(0)      # Å
        self.lj_epsilon = np.zeros(0)    # kcal/mol
        self.residue_names = []
        self.residue_pointers = np.zeros(1, int)
        self.atom_residue = np.zeros(0, int)
        self.bonds = np.zeros((0, 2), int)          # i, j
        self.bond_k = np.zeros(0)                   # kcal/mol/Å² (AMBER K, E = K (r-r0)²)
        self.bond_r0 = np.zeros(0)
        self.angles = np.zeros((0, 3), int)
        self.angle_k = np.zeros(0)                  # kcal/mol/rad²
        self.angle_t0 = np.zeros(0)
        self.dihedrals = np.zeros((0, 4), int)
        self.dihedral_k = np.zeros(0)
        self.dihedral_per = np.zeros(0)
        self.dihedral_phase = np.zeros(0)
        self.dihedral_scee = np.zeros(0)
        self.dihedral_scnb = np.zeros(0)
        self.dihedral_ignore_end = np.zeros(0, bool)
        self.dihedral_improper = np.zeros(0, bool)
        self.excluded_atoms = None                  # list of sets straight from the prmtop (for checking)
        self.coordinates = None                     # (N,3) Å
        self._velocities = None                     # (N,3) Å/ps
        self.box = None                             # [a,b,c,alpha,beta,gamma] Å / degrees
        self.title = ''
        self._atoms = None
        self._residues = None
        self._graph = None

    # -- views --------------------------------------------------------------------------------
    @property
    def atoms(self):
        if self._atoms is None:
            self._atoms = [Atom(self, i) for i in range(self.n_atoms)]
        return self._atoms

    @property
    def residues(self):
        if self._residues is None:
            self._residues = [Residue(self, i) for i in range(len(self.residue_names))]
        return self._residues

    @property
    def topology(self):
        return Topology(self)

    @property
    def positions(self):
        return u.Quantity(np.array(self.coordinates, dtype=float), u.angstroms)

    @positions.setter
    def positions(self, value):
        if u.is_quantity(value):
            value = value.value_in_unit(u.angstroms)
        self.coordinates = np.array(value, dtype=float).reshape(-1, 3)

    @property
    def velocities(self):
        return self._velocities

    @velocities.setter
    def velocities(self, value):
        if u.is_quantity(value):
            value = value.value_in_unit(u.angstroms / u.picoseconds)
        self._velocities = None if value is None else np.array(value, dtype=float).reshape(-1, 3)

    @property
    def box_vectors(self):
        if self.box is None:
            return None
        a, b, c = self.box[:3]
        if any(abs(x - 90.0) > 1e-6 for x in self.box[3:6]):
            raise NotImplementedError('only orthorhombic periodic boxes are supported')
        vecs = np.diag([a, b, c]).astype(float)
        return u.Quantity([tuple(vecs[0]), tuple(vecs[1]), tuple(vecs[2])], u.angstroms)

    @box_vectors.setter
    def box_vectors(self, value):
        if value is None:
            self.box = None
            return
        if u.is_quantity(value):
            value = value.value_in_unit(u.angstroms)
        v = np.asarray([[float(c._value) * c.unit.conversion_factor_to(u.angstroms) if u.is_quantity(c) else float(c)
                         for c in row] if not u.is_quantity(row) else row.value_in_unit(u.angstroms)
                        for row in value], dtype=float)
        self.box = [v[0, 0], v[1, 1], v[2, 2], 90.0, 90.0, 90.0]

    def bond_graph(self):
        if self._graph is None:
            g = [[] for _ in range(self.n_atoms)]
            for i, j in self.bonds[:, :2]:
                g[i].append(int(j))
                g[j].append(int(i))
            self._graph = g
        return self._graph

    # -- subsetting -----------------------------------------------------------------------------
    def __getitem__(self, sel):
        """``structure[list_of_atom_indices]`` → new Structure with those atoms (bonded terms fully inside kept)."""
        if isinstance(sel, (int, np.integer)):
            return self.atoms[sel]
        if isinstance(sel, str):
            sel = AmberMask(self, sel).Selected()
        idx = np.asarray(list(sel), dtype=int)
        remap = -np.ones(self.n_atoms, int)
        remap[idx] = np.arange(len(idx))
        s = Structure()
        s.n_atoms = len(idx)
        s.atom_names = [self.atom_names[i] for i in idx]
        s.atom_types = [self.atom_types[i] for i in idx]
        for name in ('atomic_numbers', 'masses', 'charges', 'lj_sigma', 'lj_epsilon'):
            setattr(s, name, getattr(self, name)[idx].copy())
        # residues
        rid = self.atom_residue[idx]
        s.residue_names, ptr, s_ar = [], [0], np.zeros(len(idx), int)
        last = None
        for k, r in enumerate(rid):
            if r != last:
                if last is not None:
                    ptr.append(k)
                s.residue_names.append(self.residue_names[r])
                last = r
            s_ar[k] = len(s.residue_names) - 1
        ptr.append(len(idx))
        s.residue_pointers = np.asarray(ptr, int)
        s.atom_residue = s_ar

        def keep(ix):
            m = remap[ix]
            return np.all(m >= 0, axis=1) if len(ix) else np.zeros(0, bool), m

        k, m = keep(self.bonds)
        s.bonds, s.bond_k, s.bond_r0 = m[k], self.bond_k[k], self.bond_r0[k]
        k, m = keep(self.angles)
        s.angles, s.angle_k, s.angle_t0 = m[k], self.angle_k[k], self.angle_t0[k]
        k, m = keep(self.dihedrals)
        s.dihedrals = m[k]
        for name in ('dihedral_k', 'dihedral_per', 'dihedral_phase', 'dihedral_scee', 'dihedral_scnb',
                     'dihedral_ignore_end', 'dihedral_improper'):
            setattr(s, name, getattr(self, name)[k])
        if self.coordinates is not None:
            s.coordinates = self.coordinates[idx].copy()
        if self._velocities is not None:
            s._velocities = self._velocities[idx].copy()
        s.box = None if self.box is None else list(self.box)
        return s

    # -- system construction --------------------------------------------------------------------
    def createSystem(self, **kwargs):
        from .system import create_system
        return create_system(self, **kwargs)

    # -- array (de)serialisation: lets fixtures travel without the original prmtop/pdb --------------
    _ARRAYS = ('atomic_numbers', 'masses', 'charges', 'lj_sigma', 'lj_epsilon', 'residue_pointers', 'atom_residue',
               'bonds', 'bond_k', 'bond_r0', 'angles', 'angle_k', 'angle_t0', 'dihedrals', 'dihedral_k',
               'dihedral_per', 'dihedral_phase', 'dihedral_scee', 'dihedral_scnb', 'dihedral_ignore_end',
               'dihedral_improper')

    def to_arrays(self):
        d = {k: np.asarray(getattr(self, k)) for k in self._ARRAYS}
        d['atom_names'] = np.asarray(self.atom_names, dtype='U8')
        d['atom_types'] = np.asarray(self.atom_types, dtype='U8')
        d['residue_names'] = np.asarray(self.residue_names, dtype='U8')
        d['coordinates'] = np.asarray(self.coordinates, float)
        if self._velocities is not None:
            d['velocities'] = np.asarray(self._velocities, float)
        if self.box is not None:
            d['box'] = np.asarray(self.box, float)
        return d

    @classmethod
    def from_arrays(cls, d):
        s = cls()
        for k in cls._ARRAYS:
            setattr(s, k, np.asarray(d[k]))
        s.atom_names = [str(x) for x in d['atom_names']]
        s.atom_types = [str(x) for x in d['atom_types']]
        s.residue_names = [str(x) for x in d['residue_names']]
        s.n_atoms = len(s.atom_names)
        s.coordinates = np.array(d['coordinates'], float)
        s._velocities = np.array(d['velocities'], float) if 'velocities' in d else None
        s.box = [float(x) for x in d['box']] if 'box' in d else None
        return s

    def save_npz(self, path):
        np.savez_compressed(path, **self.to_arrays())

    @classmethod
    def load_npz(cls, path):
        with np.load(path, allow_pickle=False) as z:
            return cls.from_arrays({k: z[k] for k in z.files})

    # -- output -----------------------------------------------------------------------------------
    def save(self, fname, format=None, overwrite=False, **kwargs):
        import os
        if isinstance(fname, str):
            fmt = (format or os.path.splitext(fname)[1].lstrip('.')).lower()
            if os.path.exists(fname) and not overwrite:
                raise IOError('%s exists; not overwriting' % fname)
            with open(fname, 'w') as f:
                self._write(f, fmt)
        else:
            self._write(fname, (format or 'pdb').lower())

    def _write(self, f, fmt):
        if fmt == 'pdb':
            write_pdb(self, f)
        elif fmt in ('rst7', 'inpcrd', 'restrt'):
            write_inpcrd(self, f)
        else:
            raise ValueError('unsupported output format %r' % fmt)


# =========================================================================================================
# AMBER prmtop / inpcrd
# =========================================================================================================
_FMT = re.compile(r'\(?(\d*)([aAiIeEfF])(\d+)(?:\.(\d+))?\)?')


def _read_prmtop_sections(path):
    sections, order = {}, []
    flag, fmt, lines = None, None, []

    def flush():
        if flag is None:
            return
        m = _FMT.search(fmt)
        kind, width = m.group(2).lower(), int(m.group(3))
        vals = []
        for ln in lines:
            ln = ln.rstrip('\n')
            for k in range(0, len(ln), width):
                tok = ln[k:k + width]
                if kind == 'a':
                    vals.append(tok.strip() if tok.strip() else tok)
                elif tok.strip():
                    vals.append(int(tok) if kind == 'i' else float(tok))
        sections[flag] = vals
        order.append(flag)

    with open(path) as fh:
        for ln in fh:
            if ln.startswith('%VERSION') or ln.startswith('%COMMENT'):
                continue
            if ln.startswith('%FLAG'):
                flush()
                flag, fmt, lines = ln.split()[1], None, []
            elif ln.startswith('%FORMAT'):
                fmt = ln[7:].strip()
            else:
                lines.append(ln)
    flush()
    return sections


def load_prmtop(path):
    sec = _read_prmtop_sections(path)
    ptr = sec['POINTERS']
    natom, ntypes = ptr[0], ptr[1]
    nres = ptr[11]
    s = Structure()
    s.n_atoms = natom
    s.title = ' '.join(str(x) for x in sec.get('TITLE', []))
    s.atom_names = [str(x).strip() for x in sec['ATOM_NAME']][:natom]
    s.atom_types = [str(x).strip() for x in sec['AMBER_ATOM_TYPE']][:natom]
    s.charges = np.asarray(sec['CHARGE'], float) / 18.2223
    s.masses = np.asarray(sec['MASS'], float)
    if 'ATOMIC_NUMBER' in sec:
        s.atomic_numbers = np.asarray(sec['ATOMIC_NUMBER'], int)
        s.atomic_numbers[s.atomic_numbers < 0] = 0
    else:
        s.atomic_numbers = np.asarray([element_from_mass(m) for m in s.masses], int)
    s.residue_names = [str(x).strip() for x in sec['RESIDUE_LABEL']][:nres]
    rp = np.asarray(sec['RESIDUE_POINTER'], int) - 1
    s.residue_pointers = np.concatenate([rp, [natom]])
    s.atom_residue = np.repeat(np.arange(nres), np.diff(s.residue_pointers))

    # Lennard-Jones: per-atom sigma/epsilon from the diagonal of the A/B tables (Lorentz-Berthelot files)
    tix = np.asarray(sec['ATOM_TYPE_INDEX'], int) - 1
    nbidx = np.asarray(sec['NONBONDED_PARM_INDEX'], int)
    acoef = np.asarray(sec['LENNARD_JONES_ACOEF'], float)
    bcoef = np.asarray(sec['LENNARD_JONES_BCOEF'], float)
    sig_t, eps_t = np.zeros(ntypes), np.zeros(ntypes)
    for t in range(ntypes):
        k = nbidx[ntypes * t + t] - 1
        if k < 0 or acoef[k] < 1e-10 or bcoef[k] < 1e-10:
            continue
        sig_t[t] = (acoef[k] / bcoef[k]) ** (1.0 / 6.0)
        eps_t[t] = bcoef[k] * bcoef[k] / (4.0 * acoef[k])
    s.lj_sigma, s.lj_epsilon = sig_t[tix], eps_t[tix]
    s._lj_tables = (tix, nbidx, acoef, bcoef, ntypes)

    def terms(inc, without, width):
        a = np.asarray(sec.get(inc, []) + sec.get(without, []), int).reshape(-1, width)
        return a

    b = terms('BONDS_INC_HYDROGEN', 'BONDS_WITHOUT_HYDROGEN', 3)
    s.bonds = b[:, :2] // 3
    s.bond_k = np.asarray(sec['BOND_FORCE_CONSTANT'], float)[b[:, 2] - 1]
    s.bond_r0 = np.asarray(sec['BOND_EQUIL_VALUE'], float)[b[:, 2] - 1]
    a = terms('ANGLES_INC_HYDROGEN', 'ANGLES_WITHOUT_HYDROGEN', 4)
    s.angles = a[:, :3] // 3
    s.angle_k = np.asarray(sec['ANGLE_FORCE_CONSTANT'], float)[a[:, 3] - 1] if len(a) else np.zeros(0)
    s.angle_t0 = np.asarray(sec['ANGLE_EQUIL_VALUE'], float)[a[:, 3] - 1] if len(a) else np.zeros(0)
    d = terms('DIHEDRALS_INC_HYDROGEN', 'DIHEDRALS_WITHOUT_HYDROGEN', 5)
    s.dihedrals = np.abs(d[:, :4]) // 3
    s.dihedral_ignore_end = d[:, 2] < 0
    s.dihedral_improper = d[:, 3] < 0
    ti = d[:, 4] - 1
    s.dihedral_k = np.asarray(sec['DIHEDRAL_FORCE_CONSTANT'], float)[ti] if len(d) else np.zeros(0)
    s.dihedral_per = np.asarray(sec['DIHEDRAL_PERIODICITY'], float)[ti] if len(d) else np.zeros(0)
    s.dihedral_phase = np.asarray(sec['DIHEDRAL_PHASE'], float)[ti] if len(d) else np.zeros(0)
    nd = len(sec['DIHEDRAL_FORCE_CONSTANT'])
    scee = np.asarray(sec.get('SCEE_SCALE_FACTOR', [1.2] * nd), float)
    scnb = np.asarray(sec.get('SCNB_SCALE_FACTOR', [2.0] * nd), float)
    s.dihedral_scee = scee[ti] if len(d) else np.zeros(0)
    s.dihedral_scnb = scnb[ti] if len(d) else np.zeros(0)

    # raw exclusion list (kept only for validation of the derived one)
    nex = np.asarray(sec['NUMBER_EXCLUDED_ATOMS'], int)
    exl = np.asarray(sec['EXCLUDED_ATOMS_LIST'], int)
    excl, pos = [], 0
    for i in range(natom):
        excl.append(set(int(x) - 1 for x in exl[pos:pos + nex[i]] if x > 0))
        pos += nex[i]
    s.excluded_atoms = excl
    if 'BOX_DIMENSIONS' in sec:
        bd = sec['BOX_DIMENSIONS']
        s.box = [bd[1], bd[2], bd[3], bd[0], bd[0], bd[0]]
    return s


def read_inpcrd(path):
    """ASCII AMBER inpcrd/restrt: returns (coords Å, velocities Å/ps or None, box or None)."""
    with open(path) as fh:
        fh.readline()
        head = fh.readline().split()
        natom = int(head[0])
        vals = []
        for ln in fh:
            ln = ln.rstrip('\n')
            vals.extend(float(ln[k:k + 12]) for k in range(0, len(ln), 12) if ln[k:k + 12].strip())
    vals = np.asarray(vals, float)
    n3 = 3 * natom
    coords = vals[:n3].reshape(natom, 3)
    rest = vals[n3:]
    vel, box = None, None
    if len(rest) >= n3:
        vel = rest[:n3].reshape(natom, 3) * 20.455  # AMBER velocity unit → Å/ps
        rest = rest[n3:]
    if len(rest) >= 6:
        box = list(rest[:6])
    elif len(rest) >= 3:
        box = list(rest[:3]) + [90.0, 90.0, 90.0]
    return coords, vel, box


def write_inpcrd(s, f):
    f.write('%s\n' % (s.title or 'restart created by blues_b200'))
    f.write('%5d  0.0000000e+00\n' % s.n_atoms)

    def block(a):
        flat = np.asarray(a).reshape(-1)
        for k in range(0, len(flat), 6):
            f.write(''.join('%12.7f' % x for x in flat[k:k + 6]) + '\n')

    block(s.coordinates)
    if s._velocities is not None:
        block(s._velocities / 20.455)
    if s.box is not None:
        f.write(''.join('%12.7f' % x for x in s.box) + '\n')


AMBER_VEL_SCALE = 20.455         # AMBER internal velocity unit -> Angstrom / ps (scale_factor of the NetCDF conventions)


def is_netcdf(path):
    with open(path, 'rb') as fh:
        return fh.read(3) == b'CDF'


def write_netcdf_restart(path, coordinates, velocities=None, box=None, time=0.0, title=''):
    """AMBER NetCDF restart (Conventions ``AMBERRESTART`` 1.0) — what the reference writes through parmed's
    ``RestartReporter(netcdf=True)`` (``blues/reporters.py:224``).  Coordinates / box in Angstrom, velocities in
    Angstrom/ps (stored in AMBER units with ``scale_factor`` 20.455), time in ps.  NetCDF-3 64-bit offset file."""
    from scipy.io import netcdf_file
    xyz = np.asarray(coordinates, float).reshape(-1, 3)
    nc = netcdf_file(path, 'w', version=2)
    try:
        nc.Conventions = 'AMBERRESTART'
        nc.ConventionVersion = '1.0'
        nc.program = 'blues_b200'
        nc.programVersion = '0.1.0'
        nc.title = title or 'restart created by blues_b200'
        nc.createDimension('spatial', 3)
        nc.createDimension('atom', len(xyz))
        v = nc.createVariable('spatial', 'c', ('spatial',))
        v[:] = np.frombuffer(b'xyz', dtype='S1')
        v = nc.createVariable('time', 'd', ())
        v.units = 'picosecond'
        v.data[...] = float(time)                  # (scipy's assignValue indexes a 0-d array with [:])
        v = nc.createVariable('coordinates', 'd', ('atom', 'spatial'))
        v.units = 'angstrom'
        v[:] = xyz
        if velocities is not None:
            v = nc.createVariable('velocities', 'd', ('atom', 'spatial'))
            v.units = 'angstrom/picosecond'
            v.scale_factor = np.float64(AMBER_VEL_SCALE)
            v[:] = np.asarray(velocities, float).reshape(-1, 3) / AMBER_VEL_SCALE
        if box is not None:
            nc.createDimension('cell_spatial', 3)
            nc.createDimension('cell_angular', 3)
            nc.createDimension('label', 5)
            v = nc.createVariable('cell_spatial', 'c', ('cell_spatial',))
            v[:] = np.frombuffer(b'abc', dtype='S1')
            v = nc.createVariable('cell_angular', 'c', ('cell_angular', 'label'))
            v[:] = np.frombuffer(b'alphabeta gamma', dtype='S1').reshape(3, 5)
            v = nc.createVariable('cell_lengths', 'd', ('cell_spatial',))
            v.units = 'angstrom'
            v[:] = np.asarray(box[:3], float)
            v = nc.createVariable('cell_angles', 'd', ('cell_angular',))
            v.units = 'degree'
            v[:] = np.asarray(list(box[3:6]) if len(box) >= 6 else [90.0, 90.0, 90.0], float)
    finally:
        nc.close()


def read_netcdf_restart(path):
    """(coords Angstrom, velocities Angstrom/ps or None, box or None, time ps) of an AMBER NetCDF restart."""
    from scipy.io import netcdf_file
    nc = netcdf_file(path, 'r', mmap=False)
    try:
        conv = getattr(nc, 'Conventions', b'')
        conv = conv.decode() if isinstance(conv, bytes) else str(conv)
        if 'AMBERRESTART' not in conv:
            raise ValueError('%s is a NetCDF file but not an AMBER restart (Conventions = %r)' % (path, conv))
        coords = np.array(nc.variables['coordinates'][:], float).reshape(-1, 3)
        vel = None
        if 'velocities' in nc.variables:
            var = nc.variables['velocities']
            vel = np.array(var[:], float).reshape(-1, 3) * float(getattr(var, 'scale_factor', 1.0))
        box = None
        if 'cell_lengths' in nc.variables:
            ang = np.array(nc.variables['cell_angles'][:], float) if 'cell_angles' in nc.variables else [90.0] * 3
            box = [float(x) for x in nc.variables['cell_lengths'][:]] + [float(x) for x in ang]
        time = float(nc.variables['time'].getValue()) if 'time' in nc.variables else 0.0
    finally:
        nc.close()
    return coords, vel, box, time


class Rst7(object):
    """Restart-file reader, ASCII or NetCDF by content like ``parmed.amber.Rst7`` (``blues/settings.py:79-85``)."""

    def __init__(self, filename):
        self.time = 0.0
        if is_netcdf(filename):
            c, v, b, self.time = read_netcdf_restart(filename)
        else:
            c, v, b = read_inpcrd(filename)
        self.coordinates = c
        self.vels = v
        self.box = b
        self.hasvels = v is not None
        self.hasbox = b is not None

    @property
    def positions(self):
        return u.Quantity(np.array(self.coordinates), u.angstroms)

    @property
    def velocities(self):
        return None if self.vels is None else u.Quantity(np.array(self.vels), u.angstroms / u.picoseconds)

    @property
    def box_vectors(self):
        s = Structure()
        s.box = self.box
        return s.box_vectors


# =========================================================================================================
# PDB
# =========================================================================================================
def load_pdb(path):
    names, resn, resid, chain, xyz, elem = [], [], [], [], [], []
    box, conect = None, []
    model_done = False
    with open(path) as fh:
        for ln in fh:
            rec = ln[:6]
            if rec == 'CRYST1':
                box = [float(ln[6:15]), float(ln[15:24]), float(ln[24:33]),
                       float(ln[33:40]), float(ln[40:47]), float(ln[47:54])]
            elif rec in ('ATOM  ', 'HETATM') and not model_done:
                names.append(ln[12:16].strip())
                resn.append(ln[17:21].strip())
                chain.append(ln[21])
                resid.append(ln[22:27])
                xyz.append((float(ln[30:38]), float(ln[38:46]), float(ln[46:54])))
                e = ln[76:78].strip() if len(ln) >= 78 else ''
                elem.append(e)
            elif rec == 'CONECT':
                f = [int(ln[k:k + 5]) for k in range(6, len(ln.rstrip()), 5) if ln[k:k + 5].strip()]
                conect.extend((f[0] - 1, j - 1) for j in f[1:])
            elif rec == 'ENDMDL':
                model_done = True
    s = Structure()
    n = len(names)
    s.n_atoms = n
    s.atom_names = names
    s.atom_types = list(names)
    z = []
    for nm, e in zip(names, elem):
        el = Element.from_symbol(e) if e else Element.from_symbol(re.sub(r'[^A-Za-z]', '', nm)[:1])
        z.append(el.atomic_number)
    s.atomic_numbers = np.asarray(z, int)
    s.masses = np.asarray([_WEIGHTS[k] for k in z], float)
    s.charges = np.zeros(n)
    s.lj_sigma = np.zeros(n)
    s.lj_epsilon = np.zeros(n)
    ptr, rnames = [], []
    last = None
    for k in range(n):
        key = (chain[k], resid[k], resn[k])
        if key != last:
            ptr.append(k)
            rnames.append(resn[k])
            last = key
    s.residue_names = rnames
    s.residue_pointers = np.asarray(ptr + [n], int)
    s.atom_residue = np.repeat(np.arange(len(rnames)), np.diff(s.residue_pointers))
    s.coordinates = np.asarray(xyz, float).reshape(-1, 3)
    s.box = box
    if conect:
        b = set((min(i, j), max(i, j)) for i, j in conect if i != j)
        s.bonds = np.asarray(sorted(b), int).reshape(-1, 2)
        s.bond_k = np.zeros(len(s.bonds))
        s.bond_r0 = np.zeros(len(s.bonds))
    return s


def write_pdb(s, f):
    if s.box is not None:
        f.write('CRYST1%9.3f%9.3f%9.3f%7.2f%7.2f%7.2f P 1           1\n' % tuple(s.box[:6]))
    xyz = s.coordinates
    for i in range(s.n_atoms):
        nm = s.atom_names[i]
        nm4 = (' %-3s' % nm) if len(nm) < 4 else nm[:4]
        r = int(s.atom_residue[i])
        sym = _SYMBOLS[s.atomic_numbers[i]] if s.atomic_numbers[i] < len(_SYMBOLS) else 'X'
        f.write('ATOM  %5d %4s %-4s%1s%4d    %8.3f%8.3f%8.3f  1.00  0.00          %2s\n' %
                ((i + 1) % 100000, nm4, s.residue_names[r][:4], 'A', (r + 1) % 10000,
                 xyz[i, 0], xyz[i, 1], xyz[i, 2], sym))
    f.write('END\n')


def load_file(filename, xyz=None, **kwargs):
    """``parmed.load_file`` stand-in: prmtop (+ inpcrd/rst7 coordinates) or PDB."""
    low = filename.lower()
    if low.endswith('.npz'):
        s = Structure.load_npz(filename)
    elif low.endswith('.pdb'):
        s = load_pdb(filename)
    elif low.endswith(('.inpcrd', '.rst7', '.restrt', '.crd')):
        return Rst7(filename)
    else:
        s = load_prmtop(filename)
    if xyz is not None:
        if isinstance(xyz, str):
            r = Rst7(xyz)                                   # ASCII or NetCDF, by content
            s.coordinates = r.coordinates
            s._velocities = r.vels
            if r.box is not None:
                s.box = r.box
        else:
            s.positions = xyz
    return s


# =========================================================================================================
# geometry
# =========================================================================================================
class geometry(object):
    @staticmethod
    def center_of_mass(coordinates, masses):
        """Mass-weighted mean of ``coordinates`` (N,3); dtype follows the inputs, as the reference relies on
        float32 here (``blues/moves.py:253-270``)."""
        c = np.asarray(coordinates)
        m = np.asarray(masses._value if u.is_quantity(masses) else masses).reshape(-1)
        m = m.astype(c.dtype, copy=False)
        return (c * m[:, None]).sum(axis=0) / m.sum()


# =========================================================================================================
# Amber mask
# =========================================================================================================
class AmberMask(object):
    """Subset of the AMBER mask grammar: ``:res`` / ``@atom`` lists (names with ``*``/``?`` wildcards, numbers,
    ranges), ``@%type``, ``@/element``, ``!``, ``&``, ``|``, parentheses and distance operators ``<:d``,
    ``>:d``, ``<@d``, ``>@d`` (Å, whole residues for ``:``)."""

    def __init__(self, structure, mask):
        self.s = structure
        self.mask = mask.strip()

    def Selected(self, invert=False):
        sel = self.Selection(invert)
        return iter(np.nonzero(sel)[0].tolist())

    def Selection(self, invert=False):
        toks = self._tokenize(self.mask)
        self._toks, self._p = toks, 0
        if not toks:
            res = np.zeros(self.s.n_atoms, bool)
        else:
            res = self._parse_or()
            if self._p != len(toks):
                raise ValueError('could not parse mask %r' % self.mask)
        if invert:
            res = ~res
        return res.astype(int)

    # tokens: '(', ')', '!', '&', '|', ('SEL', kind, text), ('DIST', op, kind, d)
    def _tokenize(self, m):
        toks, i, n = [], 0, len(m)
        while i < n:
            c = m[i]
            if c.isspace():
                i += 1
            elif c in '()!&|':
                toks.append(c)
                i += 1
            elif c in '<>':
                mm = re.match(r'([<>])\s*([:@])\s*([0-9.]+)', m[i:])
                if not mm:
                    raise ValueError('bad distance operator in mask %r' % m)
                toks.append(('DIST', mm.group(1), mm.group(2), float(mm.group(3))))
                i += mm.end()
            elif c in ':@':
                j = i + 1
                while j < n and m[j] not in '()!&|<>:@':
                    j += 1
                toks.append(('SEL', c, m[i + 1:j].strip()))
                i = j
            elif c == '*':
                toks.append(('SEL', '@', '*'))
                i += 1
            else:
                raise ValueError('unexpected %r in mask %r' % (c, m))
        # implicit '&' between adjacent ':res' '@atom' selectors (":LIG@C1") and before distance ops
        out = []
        for t in toks:
            if out and isinstance(t, tuple) and t[0] == 'SEL' and isinstance(out[-1], tuple) and out[-1][0] == 'SEL':
                out.append('&')
            out.append(t)
        return out

    def _peek(self):
        return self._toks[self._p] if self._p < len(self._toks) else None

    def _parse_or(self):
        v = self._parse_and()
        while self._peek() == '|':
            self._p += 1
            v = v | self._parse_and()
        return v

    def _parse_and(self):
        v = self._parse_unary()
        while True:
            t = self._peek()
            if t == '&':
                self._p += 1
                v = v & self._parse_unary()
            elif isinstance(t, tuple) and t[0] == 'DIST':
                self._p += 1
                v = self._distance(v, t)
            else:
                return v

    def _parse_unary(self):
        t = self._peek()
        if t == '!':
            self._p += 1
            return ~self._parse_unary()
        if t == '(':
            self._p += 1
            v = self._parse_or()
            if self._peek() != ')':
                raise ValueError('unbalanced parentheses in mask %r' % self.mask)
            self._p += 1
            return v
        if isinstance(t, tuple) and t[0] == 'SEL':
            self._p += 1
            return self._select(t[1], t[2])
        raise ValueError('could not parse mask %r' % self.mask)

    def _select(self, kind, text):
        s = self.s
        out = np.zeros(s.n_atoms, bool)
        for item in [x.strip() for x in text.split(',') if x.strip()]:
            if kind == '@' and item[0] == '%':
                out |= _match(s.atom_types, item[1:])
            elif kind == '@' and item[0] == '/':
                sym = [_SYMBOLS[z] if z < len(_SYMBOLS) else 'X' for z in s.atomic_numbers]
                out |= _match(sym, item[1:])
            elif re.fullmatch(r'\d+(-\d+)?', item):
                lo, _, hi = item.partition('-')
                lo, hi = int(lo), int(hi or lo)
                if kind == '@':
                    out[max(lo - 1, 0):hi] = True
                else:
                    out |= (s.atom_residue >= lo - 1) & (s.atom_residue <= hi - 1)
            elif kind == '@':
                out |= _match(s.atom_names, item)
            else:
                rm = _match(s.residue_names, item)
                out |= rm[s.atom_residue]
        return out

    def _distance(self, center, tok):
        _, op, kind, d = tok
        s = self.s
        xyz = np.asarray(s.coordinates, float)
        c = xyz[center]
        if len(c) == 0:
            within = np.zeros(s.n_atoms, bool)
        else:
            from scipy.spatial import cKDTree
            dist, _ = cKDTree(c).query(xyz, k=1)
            within = dist < d
        if kind == ':':
            res_any = np.zeros(len(s.residue_names), bool)
            np.logical_or.at(res_any, s.atom_residue, within)
            within = res_any[s.atom_residue]
        return within if op == '<' else ~within


def _match(names, pattern):
    if pattern == '*':
        return np.ones(len(names), bool)
    if '*' in pattern or '?' in pattern:
        rx = re.compile('^' + re.escape(pattern).replace(r'\*', '.*').replace(r'\?', '.') + '$')
        return np.asarray([bool(rx.match(n)) for n in names], bool)
    return np.asarray([n == pattern for n in names], bool)
