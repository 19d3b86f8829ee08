"""Monte Carlo barostat for the MD leg (``SimulationFactory.addBarostat``, ``blues/simulation.py:602-626``).

The reference attaches ``openmm.MonteCarloBarostat(pressure, temperature, frequency)`` to the MD system only
(``blues/simulation.py:781-785``: "NCMC simulation will NOT have pressure control").  OpenMM's algorithm, restated
here on top of the C ABI (``bl_get_positions`` / ``bl_set_box`` / ``bl_set_positions`` / ``bl_get_energy``): every
``frequency`` steps the box volume is changed by ``dV = volumeScale * U(-1, 1)``, every molecule is translated so
that its (unweighted) centre, wrapped into the primary cell, scales with the box, and the move is accepted with
probability ``min(1, exp(-w / kT))``, ``w = dE + P dV - N_mol kT ln(V'/V)``.  ``volumeScale`` starts at 1 % of the
volume and is tuned every 10 attempts towards 25–75 % acceptance.  This is a host-driven move on the MD leg (one
state round-trip every ``frequency`` steps), not part of the NCMC hot path.
"""
import numpy as np

AVOGADRO = 6.02214076e23
BOLTZ = 0.0083144626181532          # kJ/mol/K
BAR_NM3_TO_KJ_MOL = AVOGADRO * 1e-25   # 1 bar * 1 nm^3 in kJ/mol


def molecule_ids(topo):
    """Connected components of the bond + constraint graph: atom -> molecule index (numpy int array)."""
    n = int(topo['n_atoms'])
    parent = list(range(n))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for arr in (topo['bonds'], topo['constraints']):
        for i, j in np.asarray(arr, np.int64).reshape(-1, 2):
            ri, rj = find(int(i)), find(int(j))
            if ri != rj:
                parent[ri] = rj
    roots = np.asarray([find(a) for a in range(n)])
    return np.unique(roots, return_inverse=True)[1]


def scale_molecules(x, box, new_box, mol):
    """Translate every molecule so that its wrapped centre scales with the box (rigid molecules, no internal strain)."""
    nm = int(mol.max()) + 1
    cnt = np.bincount(mol, minlength=nm).astype(float)
    cen = np.stack([np.bincount(mol, weights=x[:, k], minlength=nm) / cnt for k in range(3)], axis=1)
    wrapped = cen - np.floor(cen / box) * box
    offset = wrapped * (np.asarray(new_box) / np.asarray(box)) - cen
    return x + offset[mol]


class MonteCarloBarostatDriver(object):
    """State of one barostat (volume step size, counters, RNG) acting on an engine with one walker."""

    def __init__(self, topo, pressure_bar, temperature, frequency=25, seed=None):
        self.pressure = float(pressure_bar) * BAR_NM3_TO_KJ_MOL
        self.kT = BOLTZ * float(temperature)
        self.frequency = int(frequency)
        self.mol = molecule_ids(topo)
        self.n_molecules = int(self.mol.max()) + 1
        self.volume_scale = None
        self.attempted = 0
        self.accepted = 0
        self.total_attempted = 0
        self.total_accepted = 0
        self.rng = np.random.RandomState(seed)

    def attempt(self, engine, uniforms=None):
        """One volume move; returns True if it was accepted.  ``uniforms`` = (u_volume, u_accept) overrides the RNG."""
        if engine.n_replicas != 1:
            raise NotImplementedError('MonteCarloBarostat needs a context with one walker (the box is shared)')
        u_vol, u_acc = uniforms if uniforms is not None else self.rng.random_sample(2)
        box = engine.get_box()
        volume = float(np.prod(box))
        if self.volume_scale is None:
            self.volume_scale = 0.01 * volume
        e0 = float(engine.get_energy(True, False)[0][0])
        x = engine.get_positions(0)
        d_volume = self.volume_scale * 2.0 * (u_vol - 0.5)
        new_volume = volume + d_volume
        new_box = box * (new_volume / volume) ** (1.0 / 3.0)
        engine.set_box(new_box)
        engine.set_positions(scale_molecules(x, box, new_box, self.mol), 0)
        e1 = float(engine.get_energy(True, False)[0][0])
        w = e1 - e0 + self.pressure * d_volume - self.n_molecules * self.kT * np.log(new_volume / volume)
        ok = not (w > 0.0 and u_acc > np.exp(-w / self.kT)) and np.isfinite(w)
        if ok:
            self.accepted += 1
            self.total_accepted += 1
        else:
            engine.set_box(box)
            engine.set_positions(x, 0)
        self.attempted += 1
        self.total_attempted += 1
        if self.attempted >= 10:
            if self.accepted < 0.25 * self.attempted:
                self.volume_scale /= 1.1
                self.attempted = self.accepted = 0
            elif self.accepted > 0.75 * self.attempted:
                self.volume_scale = min(self.volume_scale * 1.1, volume * 0.3)
                self.attempted = self.accepted = 0
        return ok
