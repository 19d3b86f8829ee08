"""Reading back what ``NetCDF4Reporter`` writes: the few ``mdtraj`` calls BLUES users (and the reference's own
``blues/tests/test_ethylene.py:113-163``, ``blues/example.py:54-59``) run on the trajectories of a simulation.

``load`` / ``load_netcdf`` return a ``Trajectory`` with ``xyz`` in nanometres (frames, atoms, 3) like mdtraj;
``compute_distances`` and ``compute_dihedrals`` follow mdtraj's argument order and result shapes.  The topology argument
is accepted for call compatibility and only used for an atom-count check when it names a readable structure file.
"""
import numpy as np


class Trajectory(object):
    def __init__(self, xyz, time=None, unitcell_lengths=None, unitcell_angles=None, extras=None):
        self.xyz = np.asarray(xyz, np.float32)
        self.time = np.arange(len(self.xyz), dtype=np.float32) if time is None else np.asarray(time, np.float32)
        self.unitcell_lengths = unitcell_lengths
        self.unitcell_angles = unitcell_angles
        self.extras = extras or {}          # protocolWork / alchemicalLambda per frame when the file has them

    @property
    def n_frames(self):
        return self.xyz.shape[0]

    @property
    def n_atoms(self):
        return self.xyz.shape[1]

    def __len__(self):
        return self.n_frames

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            key = slice(key, key + 1)
        cut = lambda a: None if a is None else a[key]
        return Trajectory(self.xyz[key], self.time[key], cut(self.unitcell_lengths), cut(self.unitcell_angles),
                          {k: v[key] for k, v in self.extras.items()})


def load_netcdf(filename, top=None, stride=None, atom_indices=None, frame=None):
    """AMBER NetCDF trajectory (ångström on disk) → ``Trajectory`` (nanometres)."""
    from scipy.io import netcdf_file
    nc = netcdf_file(str(filename), 'r', mmap=False)
    try:
        if 'coordinates' not in nc.variables:
            raise ValueError('%s holds no coordinates' % filename)
        xyz = np.array(nc.variables['coordinates'][:], np.float64) * 0.1
        time = np.array(nc.variables['time'][:], np.float64) if 'time' in nc.variables else None
        lengths = angles = None
        if 'cell_lengths' in nc.variables:
            lengths = np.array(nc.variables['cell_lengths'][:], np.float64) * 0.1
            angles = np.array(nc.variables['cell_angles'][:], np.float64)
        extras = {k: np.array(nc.variables[k][:], np.float64) for k in ('protocolWork', 'alchemicalLambda')
                  if k in nc.variables}
    finally:
        nc.close()
    if top is not None and isinstance(top, str):
        try:
            from .structure import load_file
            n = len(load_file(top).atoms)
        except Exception:
            n = None
        if n is not None and n != xyz.shape[1]:
            raise ValueError('topology %s has %d atoms, the trajectory %d' % (top, n, xyz.shape[1]))
    traj = Trajectory(xyz, time, lengths, angles, extras)
    if atom_indices is not None:
        traj.xyz = traj.xyz[:, np.asarray(atom_indices, int)]
    if frame is not None:
        return traj[int(frame)]
    if stride:
        return traj[::int(stride)]
    return traj


def load(filename, top=None, **kwargs):
    """``mdtraj.load`` for the formats this package writes: ``.nc`` / ``.ncdf`` trajectories, or one structure frame."""
    name = str(filename)
    if name.endswith(('.nc', '.ncdf', '.netcdf')):
        return load_netcdf(name, top=top, **kwargs)
    from .structure import load_file
    s = load_file(name)
    xyz = np.asarray(s.coordinates, np.float64)[None] * 0.1
    box = getattr(s, 'box', None)
    lengths = None if box is None else np.asarray(box[:3], np.float64)[None] * 0.1
    angles = None if box is None else np.asarray(box[3:6], np.float64)[None]
    return Trajectory(xyz, None, lengths, angles)


def _displacements(traj, a, b, periodic):
    d = traj.xyz[:, b].astype(np.float64) - traj.xyz[:, a].astype(np.float64)
    if periodic and traj.unitcell_lengths is not None:
        L = np.asarray(traj.unitcell_lengths, np.float64)[:, None, :]
        d -= L * np.round(d / L)
    return d


def compute_distances(traj, atom_pairs, periodic=True, opt=True):
    """Distances (nm) of every atom pair in every frame, shape (frames, pairs); minimum image in rectangular cells."""
    pairs = np.asarray(atom_pairs, int).reshape(-1, 2)
    d = _displacements(traj, pairs[:, 0], pairs[:, 1], periodic)
    return np.sqrt((d * d).sum(axis=2)).astype(np.float32)


def compute_dihedrals(traj, indices, periodic=True, opt=True):
    """Dihedral angles (radians, IUPAC sign) of every atom quadruple in every frame, shape (frames, quadruples)."""
    q = np.asarray(indices, int).reshape(-1, 4)
    b1 = _displacements(traj, q[:, 0], q[:, 1], periodic)
    b2 = _displacements(traj, q[:, 1], q[:, 2], periodic)
    b3 = _displacements(traj, q[:, 2], q[:, 3], periodic)
    c1, c2 = np.cross(b2, b3), np.cross(b1, b2)
    p1 = (b1 * c1).sum(axis=2) * np.sqrt((b2 * b2).sum(axis=2))
    p2 = (c1 * c2).sum(axis=2)
    return np.arctan2(p1, p2).astype(np.float32)
