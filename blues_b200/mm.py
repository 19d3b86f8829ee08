"""The OpenMM object surface BLUES touches, backed by the native engine.

``Context`` / ``State`` / ``Simulation`` / ``LangevinIntegrator`` / ``Platform`` are duck types of the
``simtk.openmm`` classes used at ``blues/simulation.py:707-745`` (``generateSimFromStruct``),
``:883-963`` (state in/out), ``blues/moves.py:292-309`` and by the reporters
(``blues/reporters.py:345-371``).  ``getState`` returns host *copies* wrapped in ``Quantity`` exactly like
OpenMM does; ``setPositions``/``setVelocities`` copy in.  Each ``Context`` owns one native handle
(``blues_b200._native.Engine``) holding ``n_replicas`` independent walkers; the classic single-walker calls
address walker 0 unless ``replica=`` is given.
"""
import logging
import os
import time as _time
import numpy as np

from . import unit as u
from . import _native
from .system import System, MonteCarloBarostat

logger = logging.getLogger(__name__)
OpenMMException = _native.EngineError


class Vec3(tuple):
    def __new__(cls, x, y, z):
        return tuple.__new__(cls, (x, y, z))

    x = property(lambda self: self[0])
    y = property(lambda self: self[1])
    z = property(lambda self: self[2])


class Platform(object):
    """The only platform: hand-written sm_100a kernels.  Any OpenMM platform name is accepted and mapped here
    (``blues/simulation.py:730-737`` passes 'CUDA' / 'OpenCL' / 'CPU')."""

    def __init__(self, requested='B200'):
        self.requested = requested

    @classmethod
    def getPlatformByName(cls, name):
        if str(name) not in ('B200', 'CUDA'):
            logger.info("platform '%s' requested; blues_b200 runs on its native CUDA (sm_100a) platform" % name)
        return cls(str(name))

    def getName(self):
        return 'B200'

    def getSpeed(self):
        return 1000.0

    def getPropertyNames(self):
        return ['DeviceIndex', 'Precision', 'Replicas']

    def getPropertyValue(self, context, prop):
        return str(context._properties.get(prop, ''))

    @staticmethod
    def getOpenMMVersion():
        return 'blues_b200'


def _strip(x, unit_):
    return x.value_in_unit(unit_) if u.is_quantity(x) else x


class LangevinIntegrator(object):
    """``openmm.LangevinIntegrator(temperature, friction, dt)`` for the MD leg (``blues/simulation.py:628-648``)."""

    def __init__(self, temperature, frictionCoeff, stepSize):
        self._temperature = float(_strip(temperature, u.kelvin))
        self._friction = float(_strip(frictionCoeff, u.picoseconds ** -1))
        self._dt = float(_strip(stepSize, u.picoseconds))
        self._tol = 1e-5
        self._seed = 0
        self._context = None

    def getTemperature(self):
        return self._temperature * u.kelvin

    def setTemperature(self, t):
        self._temperature = float(_strip(t, u.kelvin))
        self._rebind()

    def getFriction(self):
        return self._friction / u.picoseconds

    def getStepSize(self):
        return self._dt * u.picoseconds

    def setStepSize(self, dt):
        self._dt = float(_strip(dt, u.picoseconds))
        self._rebind()

    def getConstraintTolerance(self):
        return self._tol

    def setConstraintTolerance(self, tol):
        self._tol = float(tol)
        self._rebind()

    def getRandomNumberSeed(self):
        return self._seed

    def setRandomNumberSeed(self, seed):
        self._seed = int(seed)
        if self._context is not None:
            self._context._engine.set_seed(self._seed)

    @property
    def kT(self):
        return u.MOLAR_GAS_CONSTANT_R * (self._temperature * u.kelvin)

    def _rebind(self):
        if self._context is not None:
            self._bind(self._context)

    def _bind(self, context):
        self._context = context
        context._engine.set_langevin_integrator(self._temperature, self._friction, self._dt, self._tol)

    def step(self, n):
        ctx = self._context
        baro = ctx._barostat
        n = int(n)
        if baro is None:
            ctx._engine.md_run(n)
        else:
            # MonteCarloBarostat: a volume move whenever the step count reaches a multiple of its frequency
            done = 0
            while done < n:
                chunk = min(n - done, baro.frequency - ctx._md_steps % baro.frequency)
                ctx._engine.md_run(chunk)
                done += chunk
                ctx._md_steps += chunk
                if ctx._md_steps % baro.frequency == 0:
                    baro.attempt(ctx._engine)
        ctx._time += n * self._dt


class State(object):
    def __init__(self, positions=None, velocities=None, forces=None, potential=None, kinetic=None, box=None,
                 time=0.0, parameters=None):
        self._pos, self._vel, self._frc = positions, velocities, forces
        self._pe, self._ke, self._box, self._time, self._params = potential, kinetic, box, time, parameters

    def _need(self, v, what):
        if v is None:
            raise OpenMMException('Invoked %s on a State which does not contain that information.' % what)
        return v

    def getPositions(self, asNumpy=False):
        return u.Quantity(self._need(self._pos, 'getPositions()').copy(), u.nanometers)

    def getVelocities(self, asNumpy=False):
        return u.Quantity(self._need(self._vel, 'getVelocities()').copy(), u.nanometers / u.picoseconds)

    def getForces(self, asNumpy=False):
        return u.Quantity(self._need(self._frc, 'getForces()').copy(), u.kilojoules_per_mole / u.nanometers)

    def getPotentialEnergy(self):
        return u.Quantity(float(self._need(self._pe, 'getPotentialEnergy()')), u.kilojoules_per_mole)

    def getKineticEnergy(self):
        return u.Quantity(float(self._need(self._ke, 'getKineticEnergy()')), u.kilojoules_per_mole)

    def getPeriodicBoxVectors(self, asNumpy=False):
        b = np.diag(self._box)
        if asNumpy:
            return u.Quantity(b, u.nanometers)
        return u.Quantity([Vec3(*b[0]), Vec3(*b[1]), Vec3(*b[2])], u.nanometers)

    def getPeriodicBoxVolume(self):
        return u.Quantity(float(np.prod(self._box)), u.nanometers ** 3)

    def getTime(self):
        return u.Quantity(self._time, u.picoseconds)

    def getParameters(self):
        return dict(self._need(self._params, 'getParameters()'))


class Context(object):
    def __init__(self, system, integrator, platform=None, properties=None, n_replicas=None):
        if not isinstance(system, System):
            raise TypeError('Context needs a blues_b200.system.System')
        self._system = system
        self._integrator = integrator
        self._platform = platform or Platform()
        self._properties = dict(properties or {})
        dev = 0
        for key in ('DeviceIndex', 'CudaDeviceIndex', 'OpenCLDeviceIndex'):
            if key in self._properties:
                dev = int(str(self._properties[key]).split(',')[0])
        if n_replicas is None:
            n_replicas = int(self._properties.get('Replicas', 1))
        self._n_replicas = int(n_replicas)
        self._topo = system.flatten()
        # OpenMM semantics: seed 0 asks for a fresh random seed per Context (independent runs, and the md / alch / ncmc
        # contexts of one job, must not share their thermostat noise); the seed in use is reported by
        # integrator.getRandomNumberSeed()
        seed = int(getattr(integrator, '_seed', 0) or 0)
        if seed == 0:
            seed = (int.from_bytes(os.urandom(8), 'little') >> 1) or 1
            integrator._seed = seed
        self._engine = _native.Engine(self._topo, device=dev, n_replicas=self._n_replicas, seed=seed)
        self._time = 0.0
        self._molecules = None
        self._barostat = None
        self._md_steps = 0
        if isinstance(integrator, LangevinIntegrator):
            for f in system.getForces():
                if isinstance(f, MonteCarloBarostat) and self._topo['nb_method'] != 0:
                    from .barostat import MonteCarloBarostatDriver
                    self._barostat = MonteCarloBarostatDriver(self._topo, f.pressure, f.temperature, f.frequency,
                                                              seed=seed % (2 ** 32))
        integrator._bind(self)

    # -- accessors ---------------------------------------------------------------------------------
    def getSystem(self):
        return self._system

    def getIntegrator(self):
        return self._integrator

    def getPlatform(self):
        return self._platform

    def getNumReplicas(self):
        return self._n_replicas

    # -- state in ----------------------------------------------------------------------------------
    def setPositions(self, positions, replica=-1):
        self._engine.set_positions(np.asarray(_strip(positions, u.nanometers), dtype=float), replica)

    def setVelocities(self, velocities, replica=-1):
        self._engine.set_velocities(np.asarray(_strip(velocities, u.nanometers / u.picoseconds), dtype=float), replica)

    def setPeriodicBoxVectors(self, a, b, c):
        va, vb, vc = (np.asarray(_strip(v, u.nanometers), dtype=float).reshape(3) for v in (a, b, c))
        if abs(va[1]) + abs(va[2]) + abs(vb[0]) + abs(vb[2]) + abs(vc[0]) + abs(vc[1]) > 1e-9:
            raise NotImplementedError('only orthorhombic periodic boxes are supported')
        if self._topo['nb_method'] != 0:
            self._engine.set_box([va[0], vb[1], vc[2]])

    def setVelocitiesToTemperature(self, temperature, randomSeed=None):
        if randomSeed is not None:
            self._engine.set_seed(int(randomSeed))
        self._engine.velocities_to_temperature(float(_strip(temperature, u.kelvin)))

    def setTime(self, t):
        self._time = float(_strip(t, u.picoseconds))

    def applyConstraints(self, tol=None):
        x = self._engine.get_positions(0)
        self._engine.set_positions(x)

    def reinitialize(self, preserveState=False):
        pass

    # parameters: lambda_sterics / lambda_electrostatics are slaved to the integrator's lambda_step table
    def getParameter(self, name):
        return self._engine.get_global(name)

    def getParameters(self):
        out = {}
        if len(self._topo['alch_atoms']):
            for k in ('lambda_sterics', 'lambda_electrostatics'):
                out[k] = self._engine.get_global(k)
        return out

    # -- state out ---------------------------------------------------------------------------------
    def _wrap_molecules(self, x, box):
        if self._molecules is None:
            n = self._topo['n_atoms']
            parent = list(range(n))

            def find(a):
                while parent[a] != a:
                    parent[a] = parent[parent[a]]
                    a = parent[a]
                return a

            for arr in (self._topo['bonds'], self._topo['constraints']):
                for i, j in arr:
                    ri, rj = find(int(i)), find(int(j))
                    if ri != rj:
                        parent[ri] = rj
            roots = np.asarray([find(a) for a in range(n)])
            _, self._molecules = np.unique(roots, return_inverse=True)
        mol = self._molecules
        nm = mol.max() + 1
        cnt = np.bincount(mol, minlength=nm).astype(float)
        cen = np.stack([np.bincount(mol, weights=x[:, k], minlength=nm) / cnt for k in range(3)], axis=1)
        shift = np.floor(cen / box) * box
        return x - shift[mol]

    def getState(self, getPositions=False, getVelocities=False, getForces=False, getEnergy=False, getParameters=False,
                 getParameterDerivatives=False, enforcePeriodicBox=False, groups=-1, replica=0):
        eng = self._engine
        box = eng.get_box() if self._topo['nb_method'] != 0 else np.asarray(self._topo['box'], float)
        pos = vel = frc = pe = ke = None
        if getPositions:
            pos = eng.get_positions(replica)
            if enforcePeriodicBox and self._topo['nb_method'] != 0:
                pos = self._wrap_molecules(pos, box)
        if getVelocities:
            vel = eng.get_velocities(replica)
        if getForces:
            frc = eng.get_forces(replica)
        if getEnergy:
            ep, ek = eng.get_energy(True, True)
            pe, ke = ep[replica], ek[replica]
        params = self.getParameters() if getParameters else None
        return State(pos, vel, frc, pe, ke, box, self._time, params)


def chunk_limit(reporters, simulation):
    """Steps until the next listed frame index of any ``frame_indices`` reporter that is not reporting right now
    (None: no such reporter).  Keeps device-resident chunks from jumping over an explicitly requested frame."""
    best = None
    for rep in reporters:
        fn = getattr(rep, 'stepsToNextFrameIndex', None)
        if fn is None:
            continue
        n = fn(simulation)
        if n is not None and n > 0 and (best is None or n < best):
            best = n
    return best


class Simulation(object):
    """``simtk.openmm.app.Simulation`` stand-in: steps the context in chunks bounded by the reporters'
    ``describeNextReport`` so no host round-trip happens between report steps."""

    def __init__(self, topology, system, integrator, platform=None, platformProperties=None, state=None,
                 n_replicas=None):
        self.topology = topology
        self.system = system
        self.integrator = integrator
        self.currentStep = 0
        self.currentIter = 0
        self.reporters = []
        self.context = Context(system, integrator, platform, platformProperties, n_replicas=n_replicas)
        self._usesPBC = system.usesPeriodicBoundaryConditions()

    def minimizeEnergy(self, tolerance=10.0, maxIterations=0):
        tol = _strip(tolerance, u.kilojoules_per_mole / u.nanometers) if u.is_quantity(tolerance) else tolerance
        self.context._engine.minimize(int(maxIterations), float(tol))

    def step(self, steps):
        self._simulate(endStep=self.currentStep + int(steps))

    def _simulate(self, endStep):
        while self.currentStep < endStep:
            nextSteps = endStep - self.currentStep
            anyReport = False
            nextReport = []
            for rep in self.reporters:
                r = rep.describeNextReport(self)
                nextReport.append(r)
                if 0 < r[0] <= nextSteps:
                    nextSteps = r[0]
                    anyReport = True
            # reporters with explicit frame_indices answer -1 until currentStep is a listed index (reference semantics,
            # blues/reporters.py:362-367, written for a one-step-at-a-time loop): stop the chunk there
            limit = chunk_limit(self.reporters, self)
            if limit is not None and limit < nextSteps:
                nextSteps, anyReport = limit, False
            self.integrator.step(nextSteps)
            self.currentStep += nextSteps
            if anyReport:
                wrapped = [r for rep, r in zip(self.reporters, nextReport) if r[0] == nextSteps]
                need = [any(r[k] for r in wrapped) for k in (1, 2, 3, 4)]
                pbc = any((len(r) > 5 and r[5]) for r in wrapped) or (self._usesPBC and any(len(r) <= 5 for r in wrapped))
                state = self.context.getState(getPositions=need[0], getVelocities=need[1], getForces=need[2],
                                              getEnergy=need[3], getParameters=True, enforcePeriodicBox=pbc)
                for rep, r in zip(self.reporters, nextReport):
                    if r[0] == nextSteps:
                        rep.report(self, state)
