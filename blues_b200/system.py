"""Force-field parameter containers (``System`` + ``*Force`` objects) and the AMBER → System builder.

Replaces the OpenMM ``System``/``Force`` objects that ``parmed.Structure.createSystem`` hands to the
reference (``blues/simulation.py:139-219``) with plain numpy-backed parameter tables, exposing the
members BLUES and its tests use (SURVEY.md §8b ``System`` row: ``getNumParticles``,
``get/setParticleMass``, ``getForces``/``addForce``/``getNumForces`` with ``…Force`` class names).
``System.flatten()`` lowers everything to the flat array dictionary consumed by the C-ABI
(``include/blues_b200.h``) — all values in nm, ps, dalton, kJ/mol, e, radian.
"""
import copy
import math
import numpy as np

from . import unit as u

ONE_4PI_EPS0 = 138.935456  # kJ nm / (mol e^2)
KCAL = 4.184

# nonbonded method / constraint enums (``simtk.openmm.app`` names, ``blues/settings.py:206-230``)
NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME = 'NoCutoff', 'CutoffNonPeriodic', 'CutoffPeriodic', 'Ewald', 'PME'
HBonds, AllBonds, HAngles = 'HBonds', 'AllBonds', 'HAngles'
_METHOD_CODE = {NoCutoff: 0, CutoffNonPeriodic: 1, CutoffPeriodic: 2, Ewald: 4, PME: 4}


def _val(x, unit_):
    return x.value_in_unit(unit_) if u.is_quantity(x) else x


class Force(object):
    def __init__(self):
        self.force_group = 0

    def getForceGroup(self):
        return self.force_group

    def setForceGroup(self, g):
        self.force_group = int(g)

    def usesPeriodicBoundaryConditions(self):
        return False


class HarmonicBondForce(Force):
    """E = ½ k (r − r0)²   (k kJ/mol/nm², r0 nm)"""

    def __init__(self, idx=None, r0=None, k=None):
        Force.__init__(self)
        self.idx = np.zeros((0, 2), np.int32) if idx is None else np.asarray(idx, np.int32).reshape(-1, 2)
        self.r0 = np.zeros(0) if r0 is None else np.asarray(r0, float)
        self.k = np.zeros(0) if k is None else np.asarray(k, float)

    def getNumBonds(self):
        return len(self.idx)

    def addBond(self, i, j, r0, k):
        self.idx = np.vstack([self.idx, [[i, j]]]).astype(np.int32)
        self.r0 = np.append(self.r0, _val(r0, u.nanometers))
        self.k = np.append(self.k, _val(k, u.kilojoules_per_mole / u.nanometers ** 2))
        return len(self.idx) - 1

    def getBondParameters(self, n):
        return [int(self.idx[n, 0]), int(self.idx[n, 1]), self.r0[n] * u.nanometers,
                self.k[n] * u.kilojoules_per_mole / u.nanometers ** 2]


class HarmonicAngleForce(Force):
    """E = ½ k (θ − θ0)²"""

    def __init__(self, idx=None, t0=None, k=None):
        Force.__init__(self)
        self.idx = np.zeros((0, 3), np.int32) if idx is None else np.asarray(idx, np.int32).reshape(-1, 3)
        self.t0 = np.zeros(0) if t0 is None else np.asarray(t0, float)
        self.k = np.zeros(0) if k is None else np.asarray(k, float)

    def getNumAngles(self):
        return len(self.idx)

    def addAngle(self, i, j, k_, t0, k):
        self.idx = np.vstack([self.idx, [[i, j, k_]]]).astype(np.int32)
        self.t0 = np.append(self.t0, _val(t0, u.radians))
        self.k = np.append(self.k, _val(k, u.kilojoules_per_mole / u.radians ** 2))
        return len(self.idx) - 1


class PeriodicTorsionForce(Force):
    """E = k (1 + cos(n φ − φ0))"""

    def __init__(self, idx=None, n=None, phase=None, k=None):
        Force.__init__(self)
        self.idx = np.zeros((0, 4), np.int32) if idx is None else np.asarray(idx, np.int32).reshape(-1, 4)
        self.n = np.zeros(0, np.int32) if n is None else np.asarray(n, np.int32)
        self.phase = np.zeros(0) if phase is None else np.asarray(phase, float)
        self.k = np.zeros(0) if k is None else np.asarray(k, float)

    def getNumTorsions(self):
        return len(self.idx)

    def addTorsion(self, a, b, c, d, n, phase, k):
        self.idx = np.vstack([self.idx, [[a, b, c, d]]]).astype(np.int32)
        self.n = np.append(self.n, int(n)).astype(np.int32)
        self.phase = np.append(self.phase, _val(phase, u.radians))
        self.k = np.append(self.k, _val(k, u.kilojoules_per_mole))
        return len(self.idx) - 1


class NonbondedForce(Force):
    """Lennard-Jones (Lorentz–Berthelot) + Coulomb with exceptions; PME / cutoff / no-cutoff."""
    NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME = 0, 1, 2, 3, 4

    def __init__(self, n=0):
        Force.__init__(self)
        self.charge = np.zeros(n)
        self.sigma = np.zeros(n)
        self.epsilon = np.zeros(n)
        self.exc_idx = np.zeros((0, 2), np.int32)
        self.exc_qq = np.zeros(0)
        self.exc_sigma = np.zeros(0)
        self.exc_eps = np.zeros(0)
        self.method = NoCutoff
        self.cutoff = 1.0
        self.ewald_tol = 5e-4
        self.use_dispersion_correction = True
        self.switch_distance = 0.0
        self.pme_params = None  # (alpha, nx, ny, nz) override

    def getNumParticles(self):
        return len(self.charge)

    def getNumExceptions(self):
        return len(self.exc_idx)

    def getParticleParameters(self, i):
        return [self.charge[i] * u.elementary_charge, self.sigma[i] * u.nanometers,
                self.epsilon[i] * u.kilojoules_per_mole]

    def setParticleParameters(self, i, q, sigma, eps):
        self.charge[i] = _val(q, u.elementary_charge)
        self.sigma[i] = _val(sigma, u.nanometers)
        self.epsilon[i] = _val(eps, u.kilojoules_per_mole)

    def getExceptionParameters(self, k):
        return [int(self.exc_idx[k, 0]), int(self.exc_idx[k, 1]), self.exc_qq[k] * u.elementary_charge ** 2,
                self.exc_sigma[k] * u.nanometers, self.exc_eps[k] * u.kilojoules_per_mole]

    def getNonbondedMethod(self):
        return _METHOD_CODE[self.method]

    def getCutoffDistance(self):
        return self.cutoff * u.nanometers

    def getEwaldErrorTolerance(self):
        return self.ewald_tol

    def getUseDispersionCorrection(self):
        return self.use_dispersion_correction

    def setUseDispersionCorrection(self, flag):
        self.use_dispersion_correction = bool(flag)

    def usesPeriodicBoundaryConditions(self):
        return self.method in (CutoffPeriodic, Ewald, PME)

    def getPMEParameters(self, box):
        """(alpha [1/nm], nx, ny, nz) by OpenMM's rule."""
        if self.pme_params is not None:
            return self.pme_params
        return pme_parameters(self.cutoff, self.ewald_tol, box)


class CMMotionRemover(Force):
    def __init__(self, frequency=1):
        Force.__init__(self)
        self.frequency = int(frequency)

    def getFrequency(self):
        return self.frequency


class MonteCarloBarostat(Force):
    def __init__(self, pressure, temperature, frequency=25):
        Force.__init__(self)
        self.pressure = _val(pressure, u.bar)
        self.temperature = _val(temperature, u.kelvin)
        self.frequency = int(frequency)

    def getFrequency(self):
        return self.frequency

    def getDefaultPressure(self):
        return self.pressure * u.bar

    def getDefaultTemperature(self):
        return self.temperature * u.kelvin

    def usesPeriodicBoundaryConditions(self):
        return True


class CustomExternalForce(Force):
    """Positional restraint ``k*periodicdistance(x,y,z,x0,y0,z0)^2`` (``blues/simulation.py:347-362``)."""

    def __init__(self, energy='k*periodicdistance(x,y,z,x0,y0,z0)^2'):
        Force.__init__(self)
        self.energy = energy
        self.global_params = {}
        self.per_particle_names = []
        self.atoms = []
        self.params = []

    def getEnergyFunction(self):
        return self.energy

    def addGlobalParameter(self, name, value):
        self.global_params[name] = float(_val(value, u.kilojoules_per_mole / u.nanometers ** 2))
        return len(self.global_params) - 1

    def addPerParticleParameter(self, name):
        self.per_particle_names.append(name)
        return len(self.per_particle_names) - 1

    def addParticle(self, index, params=()):
        self.atoms.append(int(index))
        self.params.append([float(_val(p, u.nanometers)) for p in params])
        return len(self.atoms) - 1

    def getNumParticles(self):
        return len(self.atoms)


class CustomNonbondedForce(Force):
    """Parameter record for one softcore interaction class of the alchemical system (see alchemy.py)."""

    def __init__(self, energy, role):
        Force.__init__(self)
        self.energy = energy
        self.role = role

    def getEnergyFunction(self):
        return self.energy


class CustomBondForce(CustomNonbondedForce):
    """Parameter record for alchemically-modified exceptions (see alchemy.py)."""


class System(object):
    def __init__(self, n=0):
        self.masses = np.zeros(n)
        self.forces = []
        self.constraints = np.zeros((0, 2), np.int32)
        self.constraint_d = np.zeros(0)
        self.box = None   # (3,) nm orthorhombic
        self.alchemical = None  # dict written by alchemy.AbsoluteAlchemicalFactory

    # -- OpenMM-like surface -----------------------------------------------------------------
    def getNumParticles(self):
        return len(self.masses)

    def addParticle(self, mass):
        self.masses = np.append(self.masses, _val(mass, u.dalton))
        return len(self.masses) - 1

    def getParticleMass(self, i):
        return self.masses[i] * u.dalton

    def setParticleMass(self, i, mass):
        self.masses[i] = _val(mass, u.dalton)

    def getNumForces(self):
        return len(self.forces)

    def getForces(self):
        return list(self.forces)

    def getForce(self, i):
        return self.forces[i]

    def addForce(self, force):
        self.forces.append(force)
        return len(self.forces) - 1

    def removeForce(self, i):
        del self.forces[i]

    def getNumConstraints(self):
        return len(self.constraints)

    def addConstraint(self, i, j, d):
        self.constraints = np.vstack([self.constraints, [[i, j]]]).astype(np.int32)
        self.constraint_d = np.append(self.constraint_d, _val(d, u.nanometers))
        return len(self.constraints) - 1

    def getConstraintParameters(self, k):
        return [int(self.constraints[k, 0]), int(self.constraints[k, 1]), self.constraint_d[k] * u.nanometers]

    def getDefaultPeriodicBoxVectors(self):
        b = self.box if self.box is not None else np.array([2.0, 2.0, 2.0])
        return u.Quantity(np.diag(b), u.nanometers)

    def setDefaultPeriodicBoxVectors(self, a, b, c):
        self.box = np.array([_val(a, u.nanometers)[0], _val(b, u.nanometers)[1], _val(c, u.nanometers)[2]], float)

    def usesPeriodicBoundaryConditions(self):
        return any(f.usesPeriodicBoundaryConditions() for f in self.forces)

    def __deepcopy__(self, memo):
        s = System()
        s.masses = self.masses.copy()
        s.forces = [copy.deepcopy(f, memo) for f in self.forces]
        s.constraints = self.constraints.copy()
        s.constraint_d = self.constraint_d.copy()
        s.box = None if self.box is None else np.array(self.box)
        s.alchemical = copy.deepcopy(self.alchemical, memo)
        return s

    def _force(self, cls):
        for f in self.forces:
            if type(f) is cls:
                return f
        return None

    # -- lowering to the C-ABI tables -----------------------------------------------------------
    def flatten(self, box=None):
        """Flat dictionary of numpy arrays/scalars = the ``bl_topology`` of ``include/blues_b200.h``."""
        n = self.getNumParticles()
        t = {'n_atoms': n, 'mass': self.masses.astype(np.float64).copy()}
        box = np.asarray(self.box if box is None else box, float) if (box is not None or self.box is not None) \
            else np.array([0.0, 0.0, 0.0])
        t['box'] = box.copy()
        hb, ha, pt, nb = (self._force(c) for c in (HarmonicBondForce, HarmonicAngleForce, PeriodicTorsionForce,
                                                   NonbondedForce))
        t['bonds'] = (hb.idx if hb else np.zeros((0, 2))).astype(np.int32)
        t['bond_k'] = (hb.k if hb else np.zeros(0)).astype(np.float64)
        t['bond_r0'] = (hb.r0 if hb else np.zeros(0)).astype(np.float64)
        t['angles'] = (ha.idx if ha else np.zeros((0, 3))).astype(np.int32)
        t['angle_k'] = (ha.k if ha else np.zeros(0)).astype(np.float64)
        t['angle_t0'] = (ha.t0 if ha else np.zeros(0)).astype(np.float64)
        t['torsions'] = (pt.idx if pt else np.zeros((0, 4))).astype(np.int32)
        t['torsion_k'] = (pt.k if pt else np.zeros(0)).astype(np.float64)
        t['torsion_n'] = (pt.n if pt else np.zeros(0)).astype(np.int32)
        t['torsion_phase'] = (pt.phase if pt else np.zeros(0)).astype(np.float64)
        if nb is None:
            nb = NonbondedForce(n)
        t['charge'] = nb.charge.astype(np.float64).copy()
        t['sigma'] = nb.sigma.astype(np.float64).copy()
        t['epsilon'] = nb.epsilon.astype(np.float64).copy()
        ex = nb.exc_idx.astype(np.int32).copy()
        swap = ex[:, 0] > ex[:, 1]
        ex[swap] = ex[swap][:, ::-1]
        t['excl_pairs'] = ex
        t['excl_qq'] = nb.exc_qq.astype(np.float64).copy()
        t['excl_sigma'] = nb.exc_sigma.astype(np.float64).copy()
        t['excl_eps'] = nb.exc_eps.astype(np.float64).copy()
        t['nb_method'] = _METHOD_CODE[nb.method]
        t['cutoff'] = float(nb.cutoff)
        t['use_dispersion_correction'] = int(nb.use_dispersion_correction and nb.method in (CutoffPeriodic, Ewald, PME))
        if nb.method in (Ewald, PME):
            alpha, nx, ny, nz = nb.getPMEParameters(box)
            t['ewald_alpha'], t['pme_grid'] = float(alpha), np.array([nx, ny, nz], np.int32)
        else:
            t['ewald_alpha'], t['pme_grid'] = 0.0, np.zeros(3, np.int32)
        t['pme_order'] = 5
        t['dispersion_coeff'] = dispersion_coefficient(nb) if t['use_dispersion_correction'] else 0.0
        # constraints: drop those between two frozen atoms, refuse mixed ones
        cons, cd = self.constraints, self.constraint_d
        if len(cons):
            m0 = self.masses[cons[:, 0]] == 0
            m1 = self.masses[cons[:, 1]] == 0
            if np.any(m0 != m1):
                raise ValueError('A constraint cannot involve a massless particle and a massive one')
            keep = ~(m0 & m1)
            cons, cd = cons[keep], cd[keep]
        t['constraints'] = cons.astype(np.int32).reshape(-1, 2)
        t['constraint_d'] = cd.astype(np.float64)
        cm = self._force(CMMotionRemover)
        t['remove_cm'] = int(cm is not None)
        # positional restraints
        ra, rk, rx = [], [], []
        for f in self.forces:
            if isinstance(f, CustomExternalForce) and 'periodicdistance' in f.energy:
                # the coefficient is the global parameter multiplying periodicdistance(...)^2, whatever its name
                # ('k_restr' in blues/simulation.py:346-348)
                name = f.energy.split('*')[0].strip()
                k = f.global_params.get(name, next(iter(f.global_params.values()), 0.0))
                for a, p in zip(f.atoms, f.params):
                    ra.append(a)
                    rk.append(k)
                    rx.append(p[:3])
        t['restraint_atoms'] = np.asarray(ra, np.int32)
        t['restraint_k'] = np.asarray(rk, np.float64)
        t['restraint_x0'] = np.asarray(rx, np.float64).reshape(-1, 3)
        # alchemical region
        al = self.alchemical or {}
        t['alch_atoms'] = np.asarray(al.get('atoms', []), np.int32)
        t['alch_charge'] = np.asarray(al.get('charge', []), np.float64)
        t['alch_sigma'] = np.asarray(al.get('sigma', []), np.float64)
        t['alch_eps'] = np.asarray(al.get('epsilon', []), np.float64)
        t['alch_exc_pairs'] = np.asarray(al.get('exc_pairs', []), np.int32).reshape(-1, 2)
        t['alch_exc_qq'] = np.asarray(al.get('exc_qq', []), np.float64)
        t['alch_exc_sigma'] = np.asarray(al.get('exc_sigma', []), np.float64)
        t['alch_exc_eps'] = np.asarray(al.get('exc_eps', []), np.float64)
        for k_, d in (('softcore_alpha', 0.5), ('softcore_a', 1.0), ('softcore_b', 1.0), ('softcore_c', 6.0),
                      ('softcore_beta', 0.0), ('softcore_d', 1.0), ('softcore_e', 1.0), ('softcore_f', 2.0)):
            t[k_] = float(al.get(k_, d))
        t['annihilate_sterics'] = int(al.get('annihilate_sterics', False))
        t['annihilate_electrostatics'] = int(al.get('annihilate_electrostatics', True))
        return t


# =========================================================================================================
# parameter rules
# =========================================================================================================
def _legal_fft_size(n):
    n = max(int(n), 6)
    while True:
        m = n
        for p in (2, 3, 5, 7):
            while m % p == 0:
                m //= p
        if m == 1:
            return n
        n += 1


def pme_parameters(cutoff, tol, box):
    """OpenMM's PME parameter rule: α = sqrt(−ln 2tol)/rc, n_d = ceil(2 α L_d / (3 tol^{1/5})) → next 2·3·5·7-smooth."""
    alpha = math.sqrt(-math.log(2.0 * tol)) / cutoff
    dims = [_legal_fft_size(math.ceil(2.0 * alpha * L / (3.0 * tol ** 0.2))) for L in box]
    return alpha, dims[0], dims[1], dims[2]


def dispersion_coefficient(nb):
    """Long-range LJ correction coefficient C such that E_disp = C / V (hard cutoff, no switch)."""
    sig = np.round(nb.sigma, 12)
    eps = np.round(nb.epsilon, 12)
    classes, counts = np.unique(np.stack([sig, eps], axis=1), axis=0, return_counts=True)
    n = len(sig)
    if n == 0:
        return 0.0
    s, e = classes[:, 0], classes[:, 1]
    sij = 0.5 * (s[:, None] + s[None, :])
    eij = np.sqrt(e[:, None] * e[None, :])
    cnt = counts[:, None].astype(float) * counts[None, :]
    iu = np.triu_indices(len(s), 1)
    diag = counts * (counts + 1) / 2.0
    s6 = sij ** 6
    sum1 = np.sum(diag * e * s ** 12) + np.sum(cnt[iu] * eij[iu] * s6[iu] ** 2)
    sum2 = np.sum(diag * e * s ** 6) + np.sum(cnt[iu] * eij[iu] * s6[iu])
    ninter = n * (n + 1) / 2.0
    sum1 /= ninter
    sum2 /= ninter
    rc = nb.cutoff
    return 8.0 * n * n * math.pi * (sum1 / (9.0 * rc ** 9) - sum2 / (3.0 * rc ** 3))


def find_waters(struct):
    """Residues made of exactly one O and two H with two O–H bonds → list of (O, H1, H2)."""
    g = struct.bond_graph()
    out = []
    rp = struct.residue_pointers
    z = struct.atomic_numbers
    for r in range(len(struct.residue_names)):
        a0, a1 = rp[r], rp[r + 1]
        if a1 - a0 != 3:
            continue
        zs = z[a0:a1]
        if sorted(zs) != [1, 1, 8]:
            continue
        o = a0 + int(np.argmax(zs == 8))
        hs = [a for a in range(a0, a1) if a != o]
        if all(h in g[o] for h in hs):
            out.append((o, hs[0], hs[1]))
    return out


def create_system(struct, nonbondedMethod=None, nonbondedCutoff=8.0 * u.angstroms, switchDistance=0.0 * u.angstroms,
                  constraints=None, rigidWater=True, removeCMMotion=True, hydrogenMass=None,
                  ewaldErrorTolerance=0.0005, flexibleConstraints=True, verbose=False, splitDihedrals=False,
                  implicitSolvent=None, **kwargs):
    """Build the MD ``System`` from an AMBER-parameterised :class:`Structure` (restates what
    ``parmed.Structure.createSystem`` produces for the keyword set BLUES passes,
    ``blues/simulation.py:139-219``, ``examples/rotmove_cuda.yml:19-28``; formulas SURVEY.md Appendix A.1)."""
    if implicitSolvent is not None:
        raise NotImplementedError('implicit solvent is outside the NCMC hot path')
    n = struct.n_atoms
    method = nonbondedMethod or NoCutoff
    if hasattr(method, 'name'):
        method = method.name
    cons = constraints.name if hasattr(constraints, 'name') else constraints
    system = System(n)
    system.masses = struct.masses.astype(float).copy()
    if struct.box is not None and method in (CutoffPeriodic, Ewald, PME):
        if any(abs(x - 90.0) > 1e-6 for x in struct.box[3:6]):
            raise NotImplementedError('only orthorhombic periodic boxes are supported')
        system.box = np.asarray(struct.box[:3], float) * 0.1
    z = struct.atomic_numbers
    bonds = struct.bonds
    isH = z == 1

    # hydrogen mass repartitioning
    if hydrogenMass is not None:
        hm = _val(hydrogenMass, u.dalton)
        g = struct.bond_graph()
        for a in np.nonzero(isH)[0]:
            heavy = next((b for b in g[a] if not isH[b]), None)
            if heavy is None:
                continue
            transfer = hm - system.masses[a]
            system.masses[a] = hm
            system.masses[heavy] -= transfer

    # constraints
    waters = find_waters(struct) if rigidWater else []
    water_atoms = set(a for w in waters for a in w)
    constrained = np.zeros(len(bonds), bool)
    bond_r0_nm = struct.bond_r0 * 0.1
    if cons in (HBonds, AllBonds, HAngles):
        for k, (i, j) in enumerate(bonds):
            if cons != HBonds or isH[i] or isH[j]:
                constrained[k] = True
    if waters:
        for k, (i, j) in enumerate(bonds):
            if i in water_atoms and j in water_atoms:
                constrained[k] = True
    if cons == HAngles:
        raise NotImplementedError('HAngles constraints are not supported')
    pairs, dists = [], []
    seen = set()
    for k in np.nonzero(constrained)[0]:
        i, j = int(bonds[k, 0]), int(bonds[k, 1])
        key = (min(i, j), max(i, j))
        if key not in seen:
            seen.add(key)
            pairs.append(key)
            dists.append(bond_r0_nm[k])
    if waters:
        ang = {}
        for k, (a, b, c) in enumerate(struct.angles):
            ang[(int(a), int(b), int(c))] = k
            ang[(int(c), int(b), int(a))] = k
        blen = {}
        for k, (i, j) in enumerate(bonds):
            blen[(int(i), int(j))] = blen[(int(j), int(i))] = bond_r0_nm[k]
        for (o, h1, h2) in waters:
            key = (min(h1, h2), max(h1, h2))
            if key in seen:
                continue
            if (h1, o, h2) in ang:
                th = struct.angle_t0[ang[(h1, o, h2)]]
            else:
                th = math.radians(104.52)
            d1, d2 = blen[(o, h1)], blen[(o, h2)]
            seen.add(key)
            pairs.append(key)
            dists.append(math.sqrt(d1 * d1 + d2 * d2 - 2.0 * d1 * d2 * math.cos(th)))
    system.constraints = np.asarray(pairs, np.int32).reshape(-1, 2)
    system.constraint_d = np.asarray(dists, float)

    # bonded terms (AMBER K → OpenMM k = 2K)
    keepb = np.ones(len(bonds), bool) if flexibleConstraints else ~constrained
    system.addForce(HarmonicBondForce(bonds[keepb], bond_r0_nm[keepb], 2.0 * struct.bond_k[keepb] * KCAL * 100.0))
    keepa = np.ones(len(struct.angles), bool)
    if not flexibleConstraints and waters:
        for k, (a, b, c) in enumerate(struct.angles):
            if a in water_atoms and b in water_atoms and c in water_atoms:
                keepa[k] = False
    system.addForce(HarmonicAngleForce(struct.angles[keepa], struct.angle_t0[keepa],
                                       2.0 * struct.angle_k[keepa] * KCAL))
    nz = struct.dihedral_k != 0 if len(struct.dihedral_k) else np.zeros(0, bool)
    system.addForce(PeriodicTorsionForce(struct.dihedrals[nz], struct.dihedral_per[nz].astype(int),
                                         struct.dihedral_phase[nz], struct.dihedral_k[nz] * KCAL))

    # nonbonded
    nb = NonbondedForce(n)
    nb.charge = struct.charges.astype(float).copy()
    nb.sigma = struct.lj_sigma * 0.1
    nb.epsilon = struct.lj_epsilon * KCAL
    nb.method = method
    nb.cutoff = _val(nonbondedCutoff, u.nanometers)
    nb.ewald_tol = float(ewaldErrorTolerance)
    nb.switch_distance = _val(switchDistance, u.nanometers)
    if nb.switch_distance > 0:
        raise NotImplementedError('switching functions are not supported (BLUES leaves switchDistance unset)')
    exc = {}  # (i,j) → (qq, sigma, eps)
    for k in range(len(struct.dihedrals)):
        if struct.dihedral_ignore_end[k] or struct.dihedral_improper[k]:
            continue
        i, j = int(struct.dihedrals[k, 0]), int(struct.dihedrals[k, 3])
        key = (min(i, j), max(i, j))
        if key in exc:
            continue
        scee = struct.dihedral_scee[k] or 1.2
        scnb = struct.dihedral_scnb[k] or 2.0
        exc[key] = (nb.charge[i] * nb.charge[j] / scee, 0.5 * (nb.sigma[i] + nb.sigma[j]),
                    math.sqrt(nb.epsilon[i] * nb.epsilon[j]) / scnb)
    g = struct.bond_graph()
    zero = {}
    for i in range(n):
        for j in g[i]:
            zero[(min(i, j), max(i, j))] = True
            for k_ in g[j]:
                if k_ != i:
                    zero[(min(i, k_), max(i, k_))] = True
    for k in range(len(struct.dihedrals)):
        i, j = int(struct.dihedrals[k, 0]), int(struct.dihedrals[k, 3])
        if i != j:
            zero.setdefault((min(i, j), max(i, j)), True)
    bonded13 = set()   # 1-2 / 1-3 always win over a 1-4 through another path
    for i in range(n):
        for j in g[i]:
            bonded13.add((min(i, j), max(i, j)))
            for k_ in g[j]:
                if k_ != i:
                    bonded13.add((min(i, k_), max(i, k_)))
    keys = sorted(set(zero) | set(exc))
    idx, qq, sg, ep = [], [], [], []
    for key in keys:
        idx.append(key)
        if key in exc and key not in bonded13:
            a, b, c = exc[key]
        else:
            a, b, c = 0.0, 0.5 * (nb.sigma[key[0]] + nb.sigma[key[1]]), 0.0
        qq.append(a)
        sg.append(b)
        ep.append(c)
    nb.exc_idx = np.asarray(idx, np.int32).reshape(-1, 2)
    nb.exc_qq, nb.exc_sigma, nb.exc_eps = np.asarray(qq, float), np.asarray(sg, float), np.asarray(ep, float)
    system.addForce(nb)
    if removeCMMotion:
        system.addForce(CMMotionRemover(1))
    return system
