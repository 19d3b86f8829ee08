"""The package's own API tests (tests/test_gpu_api.py) on a machine without a GPU: the same test functions, with
``tests/oracle_engine.OracleEngine`` (the CPU oracle behind the engine interface — test infrastructure) standing in for
the CUDA engine.  Covers the host layer end to end here — BLUESSimulation flows, YAML settings and reporters, both
paths of the water move, the barostat driver — so that host-side regressions show up before GPU time is spent.  The
``-m gpu`` run executes the very same functions against the CUDA engine."""
import pytest

from blues_b200 import _native
from blues_b200.structure import Structure
from tests.oracle_engine import OracleEngine
import tests.test_gpu_api as api


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    monkeypatch.setattr(_native, 'Engine', OracleEngine)


@pytest.fixture(scope='module')
def structure():
    import os
    return Structure.load_npz(os.path.join(api.GOLDEN, 'tol_parm.npz'))


@pytest.fixture()
def blues_sim(structure):
    from blues_b200 import utils
    from blues_b200.simulation import SystemFactory, SimulationFactory, BLUESSimulation
    from blues_b200.moves import MoveEngine
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, api.system_cfg())
    simulations = SimulationFactory(systems, MoveEngine(api.NoRandomLigandRotation(structure, 'LIG')), api.sim_cfg())
    b = BLUESSimulation(simulations)
    for sim in (b._md_sim, b._alch_sim, b._ncmc_sim):
        sim.minimizeEnergy()
    return b


def test_simulation_set_state_sync_and_iteration(blues_sim, structure):
    api.test_simulation_set(blues_sim, structure)
    api.test_state_and_sync(blues_sim)
    api.test_step_ncmc_accept_reject_md(blues_sim)


def test_random_rotation_move(structure):
    api.test_random_rotation_move(structure)


def test_run_from_yaml(structure, tmp_path):
    api.test_run_from_yaml(structure, tmp_path)


@pytest.mark.parametrize('on_device', [False, True])
def test_water_translation_move(structure, on_device):
    api.test_water_translation_move(structure, on_device)


def test_blues_run_with_water_translation(structure, tmp_path):
    api.test_blues_run_with_water_translation_on_device(structure, tmp_path)


def test_md_leg_with_monte_carlo_barostat(structure):
    api.test_md_leg_with_monte_carlo_barostat(structure)


def _run_example(name, func, tmp_path, monkeypatch, yaml='rotmove_b200.yml', fixture='tol_parm.npz', out='toluene-b200',
                 **overrides):
    """The scripts under examples/ with a short protocol, in a scratch directory (they write reporter files)."""
    import importlib.util
    import os
    import shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ex = tmp_path / 'examples'
    shutil.copytree(os.path.join(root, 'examples'), str(ex))
    os.makedirs(str(tmp_path / 'tests' / 'golden'))
    shutil.copy(os.path.join(api.GOLDEN, fixture), str(tmp_path / 'tests' / 'golden'))
    monkeypatch.chdir(str(ex))
    spec = importlib.util.spec_from_file_location(name, str(ex / (name + '.py')))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    kw = dict(nIter=2, nstepsNC=6, nstepsMD=4)
    kw.update(overrides)
    blues = getattr(mod, func)(yaml, **kw)
    assert blues.accept + blues.reject == kw['nIter']
    assert os.path.exists(str(ex / (out + '.log'))) and os.path.exists(str(ex / (out + '-ncmc.nc')))
    return blues


def test_example_rotmove(tmp_path, monkeypatch):
    _run_example('example_rotmove', 'rotmove', tmp_path, monkeypatch)


def test_example_water(tmp_path, monkeypatch):
    _run_example('example_water', 'watermove', tmp_path, monkeypatch)


def test_example_sidechain(tmp_path, monkeypatch):
    """examples/example_sidechain.py (the reference's examples/example_sidechain.py:1-37 with its analysis step): SideChainMove
    on the valine dipeptide, MD frames into a NetCDF trajectory, chi1 of every frame read back through the trajectory
    reader."""
    import numpy as np
    blues = _run_example('example_sidechain', 'sidechain', tmp_path, monkeypatch, yaml='sidechain_b200.yml',
                         fixture='vac_divaline.npz', out='divaline-b200', nIter=3, nstepsNC=6, nstepsMD=500)
    chi = np.asarray(blues.dihedrals)
    assert chi.shape == (6, 1) and np.all(np.isfinite(chi)) and np.all(np.abs(chi) <= np.pi + 1e-6)
    lines = open(str(tmp_path / 'examples' / 'divaline-b200-dihedrals.txt')).read().split()
    assert len(lines) == 6 and abs(float(lines[0]) - float(chi[0, 0])) < 1e-6


def test_monte_carlo_simulation_driver(structure):
    """MonteCarloSimulation (blues/simulation.py:1260-1335): plain Metropolis moves on the MD context, no NCMC."""
    import numpy as np
    from blues_b200 import utils
    from blues_b200.simulation import SystemFactory, SimulationFactory, MonteCarloSimulation
    from blues_b200.moves import MoveEngine, RandomLigandRotationMove
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, api.system_cfg())
    cfg = api.sim_cfg()
    cfg.update(nIter=2, mc_per_iter=3, nstepsMD=2)
    simulations = SimulationFactory(systems, MoveEngine(RandomLigandRotationMove(structure, 'LIG', 11)), cfg)
    simulations.md.minimizeEnergy(maxIterations=100)
    mc = MonteCarloSimulation(simulations, cfg)
    before = simulations.md.context.getState(getPositions=True).getPositions(asNumpy=True)._value
    mc.run()
    assert mc.accept + mc.reject == 6
    after = simulations.md.context.getState(getPositions=True).getPositions(asNumpy=True)._value
    assert np.all(np.isfinite(after)) and not np.array_equal(before, after)


def test_frame_indices_reporter(structure, tmp_path):
    api.test_frame_indices_reporter_without_an_interval_reporter(structure, tmp_path)


def test_distinct_seeds(structure):
    api.test_contexts_draw_distinct_seeds_unless_one_is_configured(structure)


def test_water_move_device_predicate_follows_all_three_hooks(structure):
    """a subclass overriding one hook takes all three hooks (and the midpoint move) to the host path."""
    from blues_b200 import unit
    from blues_b200.moves import WaterTranslationMove

    class Custom(WaterTranslationMove):
        def beforeMove(self, context):
            return context

    kw = dict(protein_selection='(index 0) or (index 1)', radius=0.9 * unit.nanometers)
    assert WaterTranslationMove(structure, **kw).device_move() is not None
    assert Custom(structure, **kw).device_move() is None


def test_sidechain_move():
    api.sidechain_move_contract()
