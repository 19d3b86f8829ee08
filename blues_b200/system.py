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

    def addParticle(self, charge, sigma, epsilon):
        self.charge = np.append(self.charge, _val(charge, u.elementary_charge))
        self.sigma = np.append(self.sigma, _val(sigma, u.nanometers))
        self.epsilon = np.append(self.epsilon, _val(epsilon, u.kilojoules_per_mole))
        return len(self.charge) - 1

    def getNumExceptions(self):
        return len(self.exc_idx)

    def addException(self, i, j, chargeProd, sigma, epsilon, replace=False):
        """Exception (chargeProd, sigma, epsilon) for the pair; all zero = plain exclusion."""
        self.exc_idx = np.vstack([self.exc_idx, [[int(i), int(j)]]]).astype(np.int32)
        self.exc_qq = np.append(self.exc_qq, _val(chargeProd, u.elementary_charge ** 2))
        self.exc_sigma = np.append(self.exc_sigma, _val(sigma, u.nanometers))
        self.exc_eps = np.append(self.exc_eps, _val(epsilon, u.kilojoules_per_mole))
        return len(self.exc_idx) - 1

    def setNonbondedMethod(self, method):
        names = {0: NoCutoff, 1: CutoffNonPeriodic, 2: CutoffPeriodic, 3: Ewald, 4: PME}
        self.method = names.get(method, method)

    def setCutoffDistance(self, cutoff):
        self.cutoff = float(_val(cutoff, u.nanometers))

    def setEwaldErrorTolerance(self, tol):
        self.ewald_tol = float(tol)

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


class _CustomForce(Force):
    """Shared part of the generic ``Custom*Force`` classes: energy expression, global parameters, role.

    ``role`` is set on the parameter RECORDS that ``alchemy.AbsoluteAlchemicalFactory`` adds for API parity with
    openmmtools (their arithmetic is the engine's softcore kernel); forces built by a user or read from a serialized
    System (``XmlSerializer.deserialize``) have no role and are lowered by ``System.flatten`` to stack programs for the
    engine's custom-force evaluator."""

    def __init__(self, energy, role=None):
        Force.__init__(self)
        self.energy = str(energy)
        self.role = role
        self.global_params = {}
        self.periodic = False

    def getEnergyFunction(self):
        return self.energy

    def setEnergyFunction(self, energy):
        self.energy = str(energy)

    def addGlobalParameter(self, name, default):
        self.global_params[str(name)] = float(_val(default, None) if u.is_quantity(default) else default)
        return len(self.global_params) - 1

    def getNumGlobalParameters(self):
        return len(self.global_params)

    def getGlobalParameterName(self, k):
        return list(self.global_params)[k]

    def getGlobalParameterDefaultValue(self, k):
        return list(self.global_params.values())[k]

    def usesPeriodicBoundaryConditions(self):
        return bool(self.periodic)

    def _constants(self):
        return {k: v for k, v in self.global_params.items() if k not in ('lambda_sterics', 'lambda_electrostatics')}


class CustomNonbondedForce(_CustomForce):
    """OpenMM ``CustomNonbondedForce``: pair energy ``f(r; p1, p2, globals)`` over interaction groups or all pairs."""
    NoCutoff, CutoffNonPeriodic, CutoffPeriodic = 0, 1, 2

    def __init__(self, energy, role=None):
        _CustomForce.__init__(self, energy, role)
        self.per_particle_names = []
        self.particles = []
        self.exclusions = []
        self.groups = []
        self.method = 0
        self.cutoff = 1.0

    def addPerParticleParameter(self, name):
        self.per_particle_names.append(str(name))
        return len(self.per_particle_names) - 1

    def getNumPerParticleParameters(self):
        return len(self.per_particle_names)

    def addParticle(self, params=()):
        self.particles.append([float(v) for v in params])
        return len(self.particles) - 1

    def getNumParticles(self):
        return len(self.particles)

    def getParticleParameters(self, i):
        return tuple(self.particles[i])

    def setParticleParameters(self, i, params):
        self.particles[i] = [float(v) for v in params]

    def addExclusion(self, i, j):
        self.exclusions.append((int(i), int(j)))
        return len(self.exclusions) - 1

    def getNumExclusions(self):
        return len(self.exclusions)

    def addInteractionGroup(self, set1, set2):
        self.groups.append((sorted(int(a) for a in set1), sorted(int(a) for a in set2)))
        return len(self.groups) - 1

    def getNumInteractionGroups(self):
        return len(self.groups)

    def getInteractionGroupParameters(self, k):
        return self.groups[k]

    def setNonbondedMethod(self, m):
        self.method = int(m)
        self.periodic = self.method == self.CutoffPeriodic

    def getNonbondedMethod(self):
        return self.method

    def setCutoffDistance(self, c):
        self.cutoff = float(_val(c, u.nanometers))

    def getCutoffDistance(self):
        return self.cutoff * u.nanometers

    def _terms(self, n_atoms):
        """[(atoms A, weights A, atoms B, weights B, params)], parameter-name map of the expression."""
        if self.role is not None or not self.particles:
            return [], {}
        if len(self.particles) != n_atoms:
            raise ValueError('CustomNonbondedForce must have one particle per System particle')
        excl = set((min(a, b), max(a, b)) for a, b in self.exclusions)
        pairs, seen = [], set()
        if self.groups:
            for s1, s2 in self.groups:
                for i in s1:
                    for j in s2:
                        key = (min(i, j), max(i, j))
                        if i != j and key not in seen and key not in excl:
                            seen.add(key)
                            pairs.append((i, j))
        else:
            if n_atoms > 2000:
                raise NotImplementedError('a CustomNonbondedForce over all pairs of %d particles needs interaction groups '
                                          '(the custom-force evaluator enumerates its pairs explicitly)' % n_atoms)
            pairs = [(i, j) for i in range(n_atoms) for j in range(i + 1, n_atoms) if (i, j) not in excl]
        P = len(self.per_particle_names)
        names = {}
        for k, nm in enumerate(self.per_particle_names):
            names[nm + '1'] = k
            names[nm + '2'] = P + k
        return [([i], [1.0], [j], [1.0], self.particles[i] + self.particles[j]) for i, j in pairs], names


class CustomBondForce(_CustomForce):
    """OpenMM ``CustomBondForce``: ``f(r; per-bond parameters, globals)`` for listed particle pairs."""

    def __init__(self, energy, role=None):
        _CustomForce.__init__(self, energy, role)
        self.per_bond_names = []
        self.bonds = []

    def addPerBondParameter(self, name):
        self.per_bond_names.append(str(name))
        return len(self.per_bond_names) - 1

    def addBond(self, i, j, params=()):
        self.bonds.append((int(i), int(j), [float(v) for v in params]))
        return len(self.bonds) - 1

    def getNumBonds(self):
        return len(self.bonds)

    def getBondParameters(self, k):
        return self.bonds[k]

    def setUsesPeriodicBoundaryConditions(self, flag):
        self.periodic = bool(flag)

    def _terms(self, n_atoms):
        if self.role is not None:
            return [], {}
        return ([([i], [1.0], [j], [1.0], list(p)) for i, j, p in self.bonds],
                {nm: k for k, nm in enumerate(self.per_bond_names)})


class CustomCentroidBondForce(_CustomForce):
    """OpenMM ``CustomCentroidBondForce`` restricted to two groups per bond and energies in ``distance(g1,g2)``
    (``blues/tests/data/ethylene_system.xml:96-113``)."""

    def __init__(self, num_groups, energy):
        _CustomForce.__init__(self, energy, None)
        self.num_groups = int(num_groups)
        self.per_bond_names = []
        self.group_defs = []
        self.bonds = []

    def getNumGroupsPerBond(self):
        return self.num_groups

    def addPerBondParameter(self, name):
        self.per_bond_names.append(str(name))
        return len(self.per_bond_names) - 1

    def addGroup(self, particles, weights=None):
        self.group_defs.append(([int(a) for a in particles], None if weights is None or len(weights) == 0
                                else [float(w) for w in weights]))
        return len(self.group_defs) - 1

    def getNumGroups(self):
        return len(self.group_defs)

    def getGroupParameters(self, k):
        return self.group_defs[k]

    def addBond(self, groups, params=()):
        self.bonds.append(([int(g) for g in groups], [float(v) for v in params]))
        return len(self.bonds) - 1

    def getNumBonds(self):
        return len(self.bonds)

    def getBondParameters(self, k):
        return self.bonds[k]

    def setUsesPeriodicBoundaryConditions(self, flag):
        self.periodic = bool(flag)

    def _terms(self, n_atoms, masses=None):
        if self.num_groups != 2:
            raise NotImplementedError('CustomCentroidBondForce with %d groups per bond (only distance(g1,g2) between two '
                                      'groups is supported)' % self.num_groups)
        out = []
        for groups, params in self.bonds:
            sides = []
            for gi in groups:
                atoms, w = self.group_defs[gi]
                w = np.asarray([masses[a] for a in atoms] if w is None else w, float)     # OpenMM: default weights = masses
                if w.sum() <= 0:
                    raise ValueError('CustomCentroidBondForce group %d has zero total weight' % gi)
                sides.append((list(atoms), list(w / w.sum())))
            out.append((sides[0][0], sides[0][1], sides[1][0], sides[1][1], list(params)))
        return out, {nm: k for k, nm in enumerate(self.per_bond_names)}


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
        t.update(self._flatten_custom())
        return t

    def _flatten_custom(self):
        """Generic Custom*Force objects → the ``custom_*`` tables of ``bl_topology`` (stack programs + term lists)."""
        from . import lepton
        n = self.getNumParticles()
        terms, cutoffs, params, progs_op, progs_arg, pstart = [], [], [], [], [], [0]
        gstart, gatoms, gweights, gindex = [0], [], [], {}

        def group_id(atoms, weights):
            key = (tuple(atoms), tuple(round(w, 15) for w in weights))
            if key not in gindex:
                gindex[key] = len(gstart) - 1
                gatoms.extend(atoms)
                gweights.extend(weights)
                gstart.append(len(gatoms))
            return gindex[key]

        for f in self.forces:
            if not isinstance(f, _CustomForce) or f.role is not None:
                continue
            if isinstance(f, CustomCentroidBondForce):
                tl, names = f._terms(n, self.masses)
                dist = ['distance(g1,g2)', 'distance(g2,g1)']
            else:
                tl, names = f._terms(n)
                dist = []
            if not tl:
                continue
            ops, args = lepton.compile_program(f.energy, names, f._constants(), dist)
            prog = len(pstart) - 1
            progs_op.extend(ops)
            progs_arg.extend(args)
            pstart.append(len(progs_op))
            cut = -1.0
            if isinstance(f, CustomNonbondedForce) and f.method != CustomNonbondedForce.NoCutoff:
                cut = float(f.cutoff)
            for a, wa, b, wb, par in tl:
                terms.append([group_id(a, wa), group_id(b, wb), prog, 1 if f.periodic else 0])
                cutoffs.append(cut)
                params.append(list(par))
        npar = max([len(p) for p in params] + [0])
        ptab = np.zeros((len(terms), max(npar, 1)))
        for k, p in enumerate(params):
            ptab[k, :len(p)] = p
        return {'custom_term': np.asarray(terms, np.int32).reshape(-1, 4), 'custom_cutoff': np.asarray(cutoffs, np.float64),
                'custom_n_params': int(max(npar, 1)) if terms else 0, 'custom_params': ptab,
                'custom_group_start': np.asarray(gstart, np.int32), 'custom_group_atoms': np.asarray(gatoms, np.int32),
                'custom_group_weights': np.asarray(gweights, np.float64),
                'custom_prog_start': np.asarray(pstart, np.int32), 'custom_code_op': np.asarray(progs_op, np.int32),
                'custom_code_arg': np.asarray(progs_arg, np.float64)}


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


# =========================================================================================================
# serialized systems
# =========================================================================================================
class XmlSerializer(object):
    """``openmm.XmlSerializer.deserialize`` for serialized ``System`` objects (``blues/tests/test_ethylene.py:64-67``
    reads ``tests/data/ethylene_system.xml``).  Orthorhombic boxes; the force classes of this module."""

    @staticmethod
    def deserialize(xml):
        import xml.etree.ElementTree as ET
        root = ET.fromstring(xml)
        if root.tag != 'System':
            raise ValueError('XmlSerializer.deserialize: only <System> documents are supported (got <%s>)' % root.tag)
        masses = [float(p.get('mass')) for p in root.find('Particles')]
        system = System(len(masses))
        system.masses = np.asarray(masses, float)
        box = root.find('PeriodicBoxVectors')
        if box is not None:
            a, b, c = box.find('A'), box.find('B'), box.find('C')
            off = [a.get('y'), a.get('z'), b.get('x'), b.get('z'), c.get('x'), c.get('y')]
            if any(abs(float(v)) > 1e-12 for v in off):
                raise NotImplementedError('triclinic boxes are not supported')
            system.box = np.array([float(a.get('x')), float(b.get('y')), float(c.get('z'))])
        cons = root.find('Constraints')
        if cons is not None and len(cons):
            system.constraints = np.asarray([[int(c.get('p1')), int(c.get('p2'))] for c in cons], np.int32)
            system.constraint_d = np.asarray([float(c.get('d')) for c in cons], float)

        def params_of(el):
            out, k = [], 1
            while el.get('param%d' % k) is not None:
                out.append(float(el.get('param%d' % k)))
                k += 1
            return out

        def globals_into(force, el):
            gp = el.find('GlobalParameters')
            for p in (gp if gp is not None else []):
                force.addGlobalParameter(p.get('name'), float(p.get('default')))

        for el in (root.find('Forces') if root.find('Forces') is not None else []):
            kind = el.get('type')
            if kind == 'HarmonicBondForce':
                b = list(el.find('Bonds'))
                f = HarmonicBondForce([[int(x.get('p1')), int(x.get('p2'))] for x in b], [float(x.get('d')) for x in b],
                                      [float(x.get('k')) for x in b])
            elif kind == 'HarmonicAngleForce':
                a = list(el.find('Angles'))
                f = HarmonicAngleForce([[int(x.get('p1')), int(x.get('p2')), int(x.get('p3'))] for x in a],
                                       [float(x.get('a')) for x in a], [float(x.get('k')) for x in a])
            elif kind == 'PeriodicTorsionForce':
                tt = list(el.find('Torsions'))
                f = PeriodicTorsionForce([[int(x.get('p%d' % q)) for q in (1, 2, 3, 4)] for x in tt],
                                         [int(x.get('periodicity')) for x in tt], [float(x.get('phase')) for x in tt],
                                         [float(x.get('k')) for x in tt])
            elif kind == 'CustomNonbondedForce':
                f = CustomNonbondedForce(el.get('energy'))
                for p in el.find('PerParticleParameters'):
                    f.addPerParticleParameter(p.get('name'))
                globals_into(f, el)
                for p in el.find('Particles'):
                    f.addParticle(params_of(p))
                for x in (el.find('Exclusions') if el.find('Exclusions') is not None else []):
                    f.addExclusion(int(x.get('p1')), int(x.get('p2')))
                for g in (el.find('InteractionGroups') if el.find('InteractionGroups') is not None else []):
                    f.addInteractionGroup([int(x.get('index')) for x in g.find('Set1')], [int(x.get('index')) for x in g.find('Set2')])
                f.setNonbondedMethod(int(el.get('method', 0)))
                f.setCutoffDistance(float(el.get('cutoff', 1.0)))
            elif kind == 'CustomBondForce':
                f = CustomBondForce(el.get('energy'))
                for p in el.find('PerBondParameters'):
                    f.addPerBondParameter(p.get('name'))
                globals_into(f, el)
                for x in el.find('Bonds'):
                    f.addBond(int(x.get('p1')), int(x.get('p2')), params_of(x))
                f.setUsesPeriodicBoundaryConditions(el.get('usesPeriodic', '0') == '1')
            elif kind == 'CustomCentroidBondForce':
                f = CustomCentroidBondForce(int(el.get('groups')), el.get('energy'))
                for p in el.find('PerBondParameters'):
                    f.addPerBondParameter(p.get('name'))
                globals_into(f, el)
                for g in el.find('Groups'):
                    parts = list(g)
                    w = [x.get('weight') for x in parts]
                    if any(v is not None for v in w) and not all(v is not None for v in w):
                        raise ValueError('CustomCentroidBondForce group mixes explicit and default weights')
                    f.addGroup([int(x.get('p')) for x in parts], None if w[0] is None else [float(v) for v in w])
                for x in el.find('Bonds'):
                    gs, k = [], 1
                    while x.get('g%d' % k) is not None:
                        gs.append(int(x.get('g%d' % k)))
                        k += 1
                    f.addBond(gs, params_of(x))
                f.setUsesPeriodicBoundaryConditions(el.get('usesPeriodic', '0') == '1')
            elif kind == 'CMMotionRemover':
                f = CMMotionRemover(int(el.get('frequency', 1)))
            else:
                raise NotImplementedError('XmlSerializer.deserialize: force type %s is not supported' % kind)
            f.setForceGroup(int(el.get('forceGroup', 0)))
            system.addForce(f)
        return system
