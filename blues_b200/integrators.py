"""``AlchemicalExternalLangevinIntegrator`` with the reference's constructor and accessor surface.

Mirrors ``blues/integrators.py:8-249``.  The reference builds an OpenMM ``CustomIntegrator`` step *program*
(reset block, external-work bookkeeping, splitting pass, extra-propagation window, ``H`` step); here the same
program is executed by the native engine (``blues_b200/csrc/engine.cu``: ``bl_ncmc_run``), and this class only
holds the parameters, tabulates the ``alchemical_functions`` per ``lambda_step`` and forwards global-variable
reads/writes to the device-resident state.
"""
import numpy as np

from . import unit as u
from . import lepton

_OPENMM_ENERGY_UNIT = u.kilojoules_per_mole

_GLOBALS = ('lambda', 'lambda_step', 'step', 'protocol_work', 'shadow_work', 'heat', 'first_step', 'perturbed_pe',
            'unperturbed_pe', 'prop', 'nprop', 'prop_lambda_min', 'prop_lambda_max', 'Eold', 'Enew', 'debug',
            'lambda_sterics', 'lambda_electrostatics', 'n_lambda_steps', 'nsteps', 'kT')


def _strip(x, unit_):
    return x.value_in_unit(unit_) if u.is_quantity(x) else x


class AlchemicalExternalLangevinIntegrator(object):
    """Nonequilibrium Langevin switching with external-work accounting (see ``blues/integrators.py:9-96``).

    Parameters follow ``blues/integrators.py:98-111``: ``alchemical_functions`` (dict of Lepton strings in
    ``lambda``), ``splitting`` (tokens ``H V R O``), ``temperature``, ``collision_rate``, ``timestep``,
    ``constraint_tolerance``, ``measure_shadow_work``, ``measure_heat``, ``nsteps_neq``, ``nprop``,
    ``prop_lambda``.
    """

    def __init__(self, alchemical_functions, splitting="R V O H O V R", temperature=298.0 * u.kelvin,
                 collision_rate=1.0 / u.picoseconds, timestep=1.0 * u.femtoseconds, constraint_tolerance=1e-8,
                 measure_shadow_work=False, measure_heat=True, nsteps_neq=100, nprop=1, prop_lambda=0.3, *args,
                 **kwargs):
        if measure_shadow_work:
            raise NotImplementedError('measure_shadow_work is not supported (BLUES leaves it False)')
        self._alchemical_functions = dict(alchemical_functions)
        self._splitting = splitting
        tokens = splitting.split()
        for t in tokens:
            if t not in ('H', 'V', 'R', 'O'):
                raise ValueError("splitting token %r is not supported (use H, V, R, O)" % t)
        self._n_H = tokens.count('H')
        self._temperature = float(_strip(temperature, u.kelvin))
        self._collision_rate = float(_strip(collision_rate, u.picoseconds ** -1))
        self._timestep = float(_strip(timestep, u.picoseconds))
        self._constraint_tolerance = float(constraint_tolerance)
        self._measure_heat = measure_heat
        self._n_steps_neq = int(nsteps_neq)
        self._n_lambda_steps = self._n_steps_neq * self._n_H
        self._nprop = int(nprop)
        self._prop_lambda = self._get_prop_lambda(prop_lambda)
        self._seed = 0
        self._context = None
        # host copies of values that are only meaningful once bound
        self._pending = {}

    # -- reference helpers ------------------------------------------------------------------------------
    def _get_prop_lambda(self, prop_lambda):
        """Window of extra propagation around lambda = 0.5 (``blues/integrators.py:147-157``)."""
        hi = round(prop_lambda + 0.5, 4)
        lo = round(0.5 - prop_lambda, 4)
        if hi - lo <= 0.0:
            lo, hi = 2.0, -1.0      # outside [0, 1]: the window never opens
        return lo, hi

    @property
    def kT(self):
        return u.MOLAR_GAS_CONSTANT_R * (self._temperature * u.kelvin)

    def getTemperature(self):
        return self._temperature * u.kelvin

    def getStepSize(self):
        return self._timestep * u.picoseconds

    def getConstraintTolerance(self):
        return self._constraint_tolerance

    def setRandomNumberSeed(self, seed):
        self._seed = int(seed)
        if self._context is not None:
            self._context._engine.set_seed(self._seed)

    def getRandomNumberSeed(self):
        return self._seed

    # -- binding to a context --------------------------------------------------------------------------
    def addTabulatedFunction(self, name, function):
        """``CustomIntegrator.addTabulatedFunction``: ``name(x)`` becomes callable from the alchemical functions, e.g.
        ``{'lambda_sterics': 'sterics_tab(lambda*1000)'}`` with a ``Discrete1DFunction`` from
        ``utils.spreadLambdaProtocol`` (``blues/utils.py:306-325``).  The tables handed to the engine are rebuilt; on a bound
        integrator call it before the first step (the engine takes the tables when the Context is created)."""
        if not callable(function):
            raise TypeError('a tabulated function must be callable (Discrete1DFunction / Continuous1DFunction)')
        self._tabulated = getattr(self, '_tabulated', {})
        self._tabulated[str(name)] = function
        if getattr(self, '_context', None) is not None:
            self._bind(self._context)
        return len(self._tabulated) - 1

    def getNumTabulatedFunctions(self):
        return len(getattr(self, '_tabulated', {}))

    def getTabulatedFunctionName(self, index):
        return list(getattr(self, '_tabulated', {}))[index]

    def getTabulatedFunction(self, index):
        return list(getattr(self, '_tabulated', {}).values())[index]

    def _tables(self):
        n = self._n_lambda_steps
        funcs = getattr(self, '_tabulated', None)
        ls = lepton.tabulate(self._alchemical_functions.get('lambda_sterics', '1'), n, funcs)
        le = lepton.tabulate(self._alchemical_functions.get('lambda_electrostatics', '1'), n, funcs)
        return np.asarray(ls, float), np.asarray(le, float)

    def _bind(self, context):
        self._context = context
        ls, le = self._tables()
        context._engine.set_ncmc_integrator(self._temperature, self._collision_rate, self._timestep, self._splitting,
                                            self._n_steps_neq, self._nprop, self._prop_lambda[0], self._prop_lambda[1],
                                            ls, le, self._constraint_tolerance)
        for k, v in self._pending.items():
            context._engine.set_global(k, v)
        self._pending = {}

    def step(self, n):
        self._context._engine.ncmc_run(int(n), getattr(self, '_scheduled_move', None))
        self._context._time += n * self._timestep

    # -- global variables ------------------------------------------------------------------------------------
    def getNumGlobalVariables(self):
        return len(_GLOBALS)

    def getGlobalVariableName(self, i):
        return _GLOBALS[i]

    def getGlobalVariableByName(self, name, replica=0):
        if name not in _GLOBALS:
            raise Exception('Illegal global variable name: %s' % name)
        if self._context is None:
            defaults = {'nprop': self._nprop, 'prop': 1, 'prop_lambda_min': self._prop_lambda[0],
                        'prop_lambda_max': self._prop_lambda[1], 'n_lambda_steps': self._n_lambda_steps,
                        'nsteps': self._n_steps_neq}
            return float(self._pending.get(name, defaults.get(name, 0.0)))
        return self._context._engine.get_global(name, replica)

    def setGlobalVariableByName(self, name, value, replica=-1):
        if name not in _GLOBALS:
            raise Exception('Illegal global variable name: %s' % name)
        if self._context is None:
            self._pending[name] = float(value)
        else:
            self._context._engine.set_global(name, float(value), replica)

    def get_protocol_work(self, dimensionless=False, replica=0):
        w = self.getGlobalVariableByName('protocol_work', replica)
        if dimensionless:
            return w / self.kT.value_in_unit(_OPENMM_ENERGY_UNIT)
        return w * _OPENMM_ENERGY_UNIT

    def get_heat(self, dimensionless=False, replica=0):
        q = self.getGlobalVariableByName('heat', replica)
        return q / self.kT.value_in_unit(_OPENMM_ENERGY_UNIT) if dimensionless else q * _OPENMM_ENERGY_UNIT

    def getLogAcceptanceProbability(self, context=None, replica=0):
        """``-(protocol_work + shadow_work) / kT`` (``blues/integrators.py:233-238``)."""
        protocol = self.getGlobalVariableByName('protocol_work', replica)
        shadow = self.getGlobalVariableByName('shadow_work', replica)
        return -1.0 * (protocol + shadow) * _OPENMM_ENERGY_UNIT / self.kT

    def reset(self):
        """Zero the step / lambda / work accumulators (``blues/integrators.py:240-249``)."""
        if self._context is None:
            self._pending = {}
        else:
            self._context._engine.reset_ncmc()
