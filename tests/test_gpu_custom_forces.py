"""Generic Custom*Force terms on the CUDA engine, pinned by the reference's only known-answer system.

``blues/tests/data/ethylene_system.xml`` (read verbatim from ``tests/golden/reference_checkout``) holds a
``CustomNonbondedForce`` over an interaction group whose sigma / epsilon follow ``lambda_sterics`` /
``lambda_electrostatics`` and a ``CustomCentroidBondForce``.  ``XmlSerializer.deserialize`` + ``System.flatten`` lower
both to stack programs (``blues_b200/lepton.py``), the engine evaluates them with ``k_custom``.  The checker is the
hand-written analytic force field of ``tests/test_oracle_ethylene.py`` (independent of the expression compiler) inside
the oracle's integrator program.
"""
import math
import os

import numpy as np
import pytest

from tests.test_oracle_ethylene import EthyleneForceField, _fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, 'tests', 'golden', 'reference_checkout', 'blues', 'tests', 'data')
FUNCS = {'lambda_sterics': 'min(1, (1/0.3)*abs(lambda-0.5))',
         'lambda_electrostatics': 'step(0.2-lambda) - 1/0.2*lambda*step(0.2-lambda) + 1/0.2*(lambda-0.8)*step(lambda-0.8)'}


def _system():
    from blues_b200.system import XmlSerializer
    with open(os.path.join(DATA, 'ethylene_system.xml')) as fh:
        return XmlSerializer.deserialize(fh.read())


def test_compiled_programs_match_the_analytic_pair_energy():
    """CPU: the stack programs of the two custom forces against closed forms (value and derivative)."""
    from blues_b200 import lepton
    t = _system().flatten()
    assert len(t['custom_term']) == 13 and t['custom_n_params'] == 8
    ops, args = list(t['custom_code_op']), list(t['custom_code_arg'])
    p0, p1, p2 = t['custom_prog_start']
    par = t['custom_params'][0]
    for r, ls, le in ((0.31, 1.0, 1.0), (0.45, 0.4, 0.7), (0.8, 0.0, 0.0)):
        v, d = lepton.evaluate_program(ops[p0:p1], args[p0:p1], r, par, (ls, le))
        sig = 0.5 * (par[0] + par[4]) * ls
        eps = math.sqrt(par[1] * par[5]) * le
        q = par[2] * par[6]
        assert abs(v - (q / r ** 2 + 4 * eps * ((sig / r) ** 12 - (sig / r) ** 6))) < 1e-12
        assert abs(d - (-2 * q / r ** 3 + 4 * eps * (-12 * sig ** 12 / r ** 13 + 6 * sig ** 6 / r ** 7))) < 1e-10
    v, d = lepton.evaluate_program(ops[p1:p2], args[p1:p2], 0.2, t['custom_params'][12], (1.0, 1.0))
    assert abs(v - 0.5 * 100000.0 * 0.04) < 1e-9 and abs(d - 100000.0 * 0.2) < 1e-9


def _engine(nsteps=20, seed=5, T=200.0, dt=0.001, n_replicas=1):
    from blues_b200 import _native
    from tests.gpu_checks import lambda_tables
    topo = _system().flatten()
    ls, le = lambda_tables(nsteps, 2, FUNCS)
    eng = _native.Engine(topo, n_replicas=n_replicas, seed=seed)
    eng.set_ncmc_integrator(T, 1.0, dt, 'H V R O R V H', nsteps, 1, 0.2, 0.8, ls, le)
    return eng, topo, ls, le


@pytest.mark.gpu
def test_custom_force_energies_and_forces_on_the_engine():
    fx, topo_o = _fixture()
    ff = EthyleneForceField(fx)
    eng, topo, ls, le = _engine()
    rng = np.random.RandomState(3)
    for k in (0, 5, 13, 20, 31, 40):
        x = np.array(fx['positions_nm']) + 0.01 * rng.randn(8, 3)
        eng.set_global('lambda_step', k)
        eng.set_positions(x)
        E = eng.get_energy()[0][0]
        F = eng.get_forces()
        Eo, Fo, _ = ff.energy_forces(x, topo_o['box'], ls[k], le[k])
        assert abs(E - Eo) < 1e-5 * max(1.0, abs(Eo)), (k, E, Eo)
        assert np.max(np.abs(F - Fo)) < 1e-5 * max(1.0, np.max(np.abs(Fo))), (k, np.max(np.abs(F - Fo)))
    eng.close()


@pytest.mark.gpu
def test_ethylene_ncmc_trajectory_and_work_step_for_step():
    """The H V R O R V H program over the custom forces: positions, velocities and protocol work against the oracle's
    interpreter (same Philox noise), including a rigid rotation of the ligand at the protocol's midpoint."""
    from oracle import ncmc_oracle as orc
    fx, topo_o = _fixture()
    ff = EthyleneForceField(fx)
    nsteps, seed = 20, 5
    eng, topo, ls, le = _engine(nsteps, seed)
    x0 = np.array(fx['positions_nm'])
    eng.set_positions(x0)
    eng.velocities_to_temperature(200.0)
    o = orc.NCMCOracle(topo_o, FUNCS, 'H V R O R V H', 200.0, 1.0, 0.001, nsteps, 1, 0.3, seed, 0)
    o.ff = ff
    o.x = x0.copy()
    o.set_velocities_to_temperature(200.0, 0)
    assert np.max(np.abs(eng.get_velocities(0) - o.v)) < 1e-9
    lig = np.arange(2, 8)
    masses = np.array([12.01078, 12.01078, 1.007947, 1.007947, 1.007947, 1.007947])
    for step in range(nsteps):
        if step == nsteps // 2:
            R = orc.rotation_matrix_from_quaternion(orc.quaternion_from_uniforms(0.3, 0.6, 0.8))
            o.x = orc.rotate_ligand(o.x, lig, masses, R)
            eng.set_positions(o.x)
        eng.ncmc_run(1)
        o.step(1)
        assert np.max(np.abs(eng.get_positions(0) - o.x)) < 1e-7, step
        assert np.max(np.abs(eng.get_velocities(0) - o.v)) < 1e-5, step
        w, wo = eng.get_global('protocol_work'), o.g['protocol_work']
        assert abs(w - wo) < 1e-5 * max(1.0, abs(wo)), (step, w, wo)
    assert abs(eng.get_global('lambda') - 1.0) < 1e-12
    eng.close()


class _DistanceReporter(object):
    """Reporter protocol of blues/reporters.py:345-371: distance between atoms 0 and 2 every `interval` MD steps."""

    def __init__(self, interval):
        self.interval, self.values = interval, []

    def describeNextReport(self, simulation):
        steps = self.interval - simulation.currentStep % self.interval
        return (steps, True, False, False, False)

    def report(self, simulation, state):
        x = state.getPositions(asNumpy=True)._value
        self.values.append(float(np.linalg.norm(x[0] - x[2])))


@pytest.mark.gpu
def test_ethylene_two_state_populations_through_the_api():
    """blues/tests/test_ethylene.py:24-163 through this package's API on the CUDA engine: 5 runs x 100 iterations of
    20 NCMC + 20 MD steps at 200 K; populations of dist(0, 2) <= 0.49 nm / > 0.49 nm must come out 0.25 / 0.75."""
    from blues_b200 import mm, unit
    from blues_b200.structure import load_file
    from blues_b200.simulation import SystemFactory, SimulationFactory, BLUESSimulation
    from blues_b200.integrators import AlchemicalExternalLangevinIntegrator
    from blues_b200.moves import RandomLigandRotationMove, MoveEngine
    structure = load_file(os.path.join(DATA, 'ethylene_structure.pdb'))
    freqs, accepted = [], 0
    for run, seed in enumerate((11, 23, 37, 41, 59)):
        np.random.seed(seed)
        cfg = {'platform': 'CUDA', 'nprop': 1, 'propLambda': 0.3, 'dt': 1 * unit.femtoseconds,
               'friction': 1 / unit.picoseconds, 'temperature': 200 * unit.kelvin, 'nIter': 100, 'nstepsMD': 20,
               'nstepsNC': 20, 'propSteps': 20, 'moveStep': 10}
        mover = MoveEngine(RandomLigandRotationMove(structure, 'LIG', random_state=np.random.RandomState(seed)))
        system = _system()
        integrator = mm.LangevinIntegrator(cfg['temperature'], cfg['friction'], cfg['dt'])
        integrator.setRandomNumberSeed(seed)
        alch_integrator = mm.LangevinIntegrator(cfg['temperature'], cfg['friction'], cfg['dt'])
        alch_integrator.setRandomNumberSeed(seed)
        alch_system = SystemFactory.generateAlchSystem(system, [2, 3, 4, 5, 6, 7])
        ncmc_integrator = AlchemicalExternalLangevinIntegrator(
            nsteps_neq=cfg['nstepsNC'], alchemical_functions=FUNCS, splitting='H V R O R V H',
            temperature=cfg['temperature'], timestep=cfg['dt'])
        ncmc_integrator.setRandomNumberSeed(seed + 1)
        systems = SystemFactory(structure, [2, 3, 4, 5, 6, 7])
        systems.md, systems.alch = system, alch_system
        sims = SimulationFactory(systems, mover)
        sims.md = SimulationFactory.generateSimFromStruct(structure, system, integrator, 'CUDA')
        rep = _DistanceReporter(5)
        sims.md.reporters.append(rep)
        sims.alch = SimulationFactory.generateSimFromStruct(structure, system, alch_integrator, 'CUDA')
        sims.ncmc = SimulationFactory.generateSimFromStruct(structure, alch_system, ncmc_integrator, 'CUDA')
        blues = BLUESSimulation(sims, cfg)
        blues.run()
        d = np.asarray(rep.values)
        assert len(d) >= 100 * 4
        freqs.append([np.mean(d <= 0.49), np.mean(d > 0.49)])
        accepted += blues.accept
    avg = np.mean(freqs, axis=0)
    err = np.std(freqs, axis=0) / math.sqrt(len(freqs))
    print('populations', avg, '+-', err, 'accepted', accepted, 'of 500')
    assert accepted > 50
    assert abs(avg[0] - 0.25) < max(0.06, 2.5 * err[0]), (avg, err)
    assert abs(avg[1] - 0.75) < max(0.06, 2.5 * err[1]), (avg, err)
