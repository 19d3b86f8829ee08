"""GPU tests of the drop-in Python API, modelled on blues/tests/test_simulation.py and test_randomrotation.py."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from blues_b200 import unit, utils                     # noqa: E402
from blues_b200 import mm as openmm                    # noqa: E402
from blues_b200.structure import Structure             # noqa: E402
from blues_b200.simulation import SystemFactory, SimulationFactory, BLUESSimulation   # noqa: E402
from blues_b200.moves import RandomLigandRotationMove, MoveEngine                     # noqa: E402
from blues_b200.reporters import ReporterConfig        # noqa: E402
from blues_b200.settings import Settings               # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
STATE_KEYS = {'getPositions': True, 'getVelocities': True, 'getForces': False, 'getEnergy': True,
              'getParameters': True, 'enforcePeriodicBox': True}


@pytest.fixture(scope='module')
def structure():
    return Structure.load_npz(os.path.join(GOLDEN, 'tol_parm.npz'))


def sim_cfg():
    return {'nprop': 1, 'propLambda': 0.3, 'dt': 0.002 * unit.picoseconds, 'friction': 1 * 1 / unit.picoseconds,
            'temperature': 300 * unit.kelvin, 'nIter': 1, 'nstepsMD': 10, 'nstepsNC': 10, 'platform': 'CUDA'}


def system_cfg():
    return {'nonbondedMethod': 'PME', 'nonbondedCutoff': 8.0 * unit.angstroms, 'constraints': 'HBonds'}


class NoRandomLigandRotation(RandomLigandRotationMove):
    def move(self, context):
        return context


@pytest.fixture(scope='module')
def blues_sim(structure):
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, system_cfg())
    engine = MoveEngine(NoRandomLigandRotation(structure, 'LIG'))
    simulations = SimulationFactory(systems, engine, sim_cfg())
    b = BLUESSimulation(simulations)
    b._md_sim.minimizeEnergy()
    b._alch_sim.minimizeEnergy()
    b._ncmc_sim.minimizeEnergy()
    return b


def test_simulation_set(blues_sim, structure):
    """tests/test_simulation.py:291-327"""
    for sim in (blues_sim._md_sim, blues_sim._alch_sim, blues_sim._ncmc_sim):
        assert isinstance(sim, openmm.Simulation)
        box = sim.context.getState().getPeriodicBoxVectors(asNumpy=True).value_in_unit(unit.angstroms)
        assert np.allclose(np.diag(box), structure.box[:3])
    integ = blues_sim._ncmc_sim.context._integrator
    assert integ._n_lambda_steps == 20 and integ._n_steps_neq == 10


def test_state_and_sync(blues_sim):
    """tests/test_simulation.py:331-383"""
    st = BLUESSimulation.getStateFromContext(blues_sim._md_sim.context, STATE_KEYS)
    assert set(st) == {'positions', 'velocities', 'potential_energy', 'kinetic_energy', 'box_vectors'}
    blues_sim._syncStatesMDtoNCMC()
    md = BLUESSimulation.getStateFromContext(blues_sim._md_sim.context, STATE_KEYS)
    nc = BLUESSimulation.getStateFromContext(blues_sim._ncmc_sim.context, STATE_KEYS)
    assert np.array_equal(md['positions']._value, nc['positions']._value)


def test_step_ncmc_accept_reject_md(blues_sim):
    """tests/test_simulation.py:385-428"""
    blues_sim._syncStatesMDtoNCMC()
    before = BLUESSimulation.getStateFromContext(blues_sim._ncmc_sim.context, STATE_KEYS)
    blues_sim._stepNCMC(10, 5)
    after = BLUESSimulation.getStateFromContext(blues_sim._ncmc_sim.context, STATE_KEYS)
    assert np.not_equal(before['positions']._value, after['positions']._value).all()
    assert blues_sim._ncmc_sim.integrator.getGlobalVariableByName('lambda') == pytest.approx(1.0)
    corr = blues_sim._computeAlchemicalCorrection()
    assert isinstance(corr, float)
    # force accept / reject through the work
    integ = blues_sim._ncmc_sim.context._integrator
    integ.setGlobalVariableByName('protocol_work', -1e6)
    blues_sim._acceptRejectMove()
    md = BLUESSimulation.getStateFromContext(blues_sim._md_sim.context, STATE_KEYS)
    assert np.allclose(md['positions']._value, after['positions']._value)
    # next iteration: a move carrying the 999999 sentinel is rejected and the MD state is left untouched
    blues_sim._resetSimulations()
    blues_sim._syncStatesMDtoNCMC()
    blues_sim._stepNCMC(10, 5)
    integ.setGlobalVariableByName('protocol_work', 999999)
    n_rej = blues_sim.reject
    md_before = BLUESSimulation.getStateFromContext(blues_sim._md_sim.context, STATE_KEYS)
    blues_sim._acceptRejectMove()
    assert blues_sim.reject == n_rej + 1
    md_after = BLUESSimulation.getStateFromContext(blues_sim._md_sim.context, STATE_KEYS)
    assert np.array_equal(md_before['positions']._value, md_after['positions']._value)
    nc_after = BLUESSimulation.getStateFromContext(blues_sim._ncmc_sim.context, STATE_KEYS)
    assert np.not_equal(md_after['positions']._value, nc_after['positions']._value).any()
    v0 = blues_sim._md_sim.context.getState(getVelocities=True).getVelocities(asNumpy=True)._value
    blues_sim._resetSimulations()
    v1 = blues_sim._md_sim.context.getState(getVelocities=True).getVelocities(asNumpy=True)._value
    assert np.not_equal(v0, v1).all()
    assert integ.getGlobalVariableByName('step') == 0 and integ.getGlobalVariableByName('protocol_work') == 0
    x0 = blues_sim._md_sim.context.getState(getPositions=True).getPositions(asNumpy=True)._value
    blues_sim._stepMD(10)
    x1 = blues_sim._md_sim.context.getState(getPositions=True).getPositions(asNumpy=True)._value
    assert np.not_equal(x0, x1).all()


def test_random_rotation_move(structure):
    """tests/test_randomrotation.py:52-61 — seeded host path and the on-device path both rotate every ligand atom."""
    system = structure.createSystem(nonbondedMethod='NoCutoff', constraints='HBonds')
    for seed in (3134, None):
        move = RandomLigandRotationMove(structure, 'LIG', seed)
        integ = openmm.LangevinIntegrator(300 * unit.kelvin, 1, 0.002 * unit.picoseconds)
        sim = SimulationFactory.generateSimFromStruct(structure, system, integ)
        x0 = sim.context.getState(getPositions=True).getPositions(asNumpy=True)._value[move.atom_indices]
        if seed is None:
            d = move.device_move()
            assert d is not None
            sim.context._engine.apply_move(d['kind'], d['atoms'], d['masses'])
            move._after_device_move(sim.context)
        else:
            assert move.device_move() is None
            sim.context = move.move(sim.context)
        x1 = sim.context.getState(getPositions=True).getPositions(asNumpy=True)._value[move.atom_indices]
        assert np.not_equal(x0, x1).all()
        d0 = np.linalg.norm(x0[0] - x0[5]); d1 = np.linalg.norm(x1[0] - x1[5])
        assert d0 == pytest.approx(d1, rel=1e-9)                  # rigid
        rest0 = sim.context.getState(getPositions=True).getPositions(asNumpy=True)._value[15:]
        assert np.allclose(rest0, structure.coordinates[15:] * 0.1)


def test_run_from_yaml(structure, tmp_path):
    """tests/test_simulation.py:430-491"""
    yaml_cfg = """
        output_dir: %s
        outfname: tol-test
        logger:
          level: info
          stream: False
        system:
          nonbondedMethod: PME
          nonbondedCutoff: 8.0 * angstroms
          constraints: HBonds
        simulation:
          dt: 0.002 * picoseconds
          friction: 1 * 1/picoseconds
          temperature: 300 * kelvin
          nIter: 2
          nstepsMD: 4
          nstepsNC: 4
          platform: CUDA
        md_reporters:
          restart:
            reportInterval: 4
          stream:
            title: md
            reportInterval: 1
            totalSteps: 8
            step: True
            speed: True
            progress: True
            remainingTime: True
            currentIter : True
        ncmc_reporters:
          traj_netcdf:
            frame_indices: [1, 0.5, -1]
            alchemicalLambda: True
            protocolWork: True
          stream:
            title: ncmc
            reportInterval: 1
            totalSteps: 4
            step: True
            speed: True
            progress: True
            remainingTime: True
            protocolWork : True
            alchemicalLambda : True
            currentIter : True
    """ % tmp_path
    cfg = Settings(yaml_cfg).asDict()
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, cfg['system'])
    engine = MoveEngine(RandomLigandRotationMove(structure, 'LIG'))
    simulations = SimulationFactory(systems, engine, cfg['simulation'], cfg['md_reporters'], cfg['ncmc_reporters'])
    blues = BLUESSimulation(simulations)
    blues._md_sim.minimizeEnergy()
    before = blues._md_sim.context.getState(getPositions=True).getPositions(asNumpy=True)._value
    blues.run()
    after = blues._md_sim.context.getState(getPositions=True).getPositions(asNumpy=True)._value
    assert np.not_equal(before, after).all()
    assert blues.accept + blues.reject == 2
    # the MD leg's restart file (NetCDF, as the reference's ReporterConfig asks) holds the final MD state
    from blues_b200.structure import Rst7, is_netcdf
    rst = os.path.join(str(tmp_path), 'tol-test.rst7')
    assert is_netcdf(rst)
    r = Rst7(rst)
    box = np.asarray(r.box[:3]) * 0.1
    d = r.positions.value_in_unit(unit.nanometers) - after
    d -= box * np.round(d / box)                                     # the reporter wraps molecules into the box
    assert np.max(np.abs(d)) < 1e-9 and r.hasvels
    log = open(os.path.join(str(tmp_path), 'tol-test.log')).read()
    assert 'ncmc:' in log and 'md:' in log and 'Acceptance Ratio' in log
    for rep in cfg['ncmc_reporters']:
        if hasattr(rep, 'close'):
            rep.close()
    from scipy.io import netcdf_file
    nc = netcdf_file(os.path.join(str(tmp_path), 'tol-test-ncmc.nc'), 'r', mmap=False)
    assert nc.variables['coordinates'].shape[1:] == (975, 3) and nc.variables['coordinates'].shape[0] >= 4
    assert 'protocolWork' in nc.variables and 'alchemicalLambda' in nc.variables
    nc.close()


def test_no_cpu_fallback_message():
    from blues_b200 import _native
    assert 'no CPU fallback' in (_native.load_library.__doc__ + open(_native.__file__).read())


def _selection_com(move, xyz_nm):
    from blues_b200.structure import geometry
    return np.asarray(geometry.center_of_mass(np.asarray(xyz_nm, np.float32)[move.protein_atoms], move.protein_masses),
                      float).reshape(3)


@pytest.mark.parametrize('on_device', [False, True])
def test_water_translation_move(structure, on_device):
    """tests/test_watertranslation.py:54-112: the alchemical water is swapped with one inside the sphere, translated to a
    point of the sphere at the midpoint, and a water left outside forces rejection (protocol_work = 999999).
    Host path (numpy RNG, state round-trips) and device path (bl_apply_move kernels) obey the same contract."""
    from blues_b200.moves import WaterTranslationMove
    np.random.seed(7)
    move = WaterTranslationMove(structure, protein_selection='(index 0) or (index 1)', radius=0.9 * unit.nanometers,
                                on_device=on_device)
    engine = MoveEngine(move)
    engine.selectMove()
    systems = SystemFactory(structure, move.atom_indices, system_cfg())
    cfg = sim_cfg()
    cfg['nstepsNC'] = 100
    simulations = SimulationFactory(systems, engine, cfg)
    ncmc = simulations.ncmc
    idx = move.atom_indices
    box = np.asarray(structure.box[:3]) * 0.1

    def positions():
        return ncmc.context.getState(getPositions=True).getPositions(asNumpy=True)

    def pdist(a, b):
        d = np.asarray(a, float) - np.asarray(b, float)
        d -= box * np.round(d / box)
        return np.linalg.norm(d)

    start = positions().value_in_unit(unit.nanometers)
    vel0 = ncmc.context.getState(getVelocities=True).getVelocities(asNumpy=True)._value
    com = _selection_com(move, start)
    before = start[idx, :]
    ncmc.context = move.beforeMove(ncmc.context)
    after_swap = positions().value_in_unit(unit.nanometers)
    swapped = after_swap[idx, :]
    assert move.go
    # the swap is a permutation of two waters: the alchemical slot holds the coordinates (and velocities) of a water
    # whose oxygen was inside the sphere, that water holds the alchemical water's
    partner = [w for w in move.water_residues if np.array_equal(start[w], swapped)]
    assert len(partner) == 1
    partner = partner[0]
    assert pdist(start[partner[0]], com) <= 0.9 + 1e-6
    assert np.array_equal(after_swap[partner], before)
    vel1 = ncmc.context.getState(getVelocities=True).getVelocities(asNumpy=True)._value
    assert np.array_equal(vel1[idx], vel0[partner]) and np.array_equal(vel1[partner], vel0[idx])
    others = np.setdiff1d(np.arange(len(start)), np.concatenate([idx, partner]))
    assert np.array_equal(after_swap[others], start[others])
    if partner != list(idx):
        assert np.not_equal(before, swapped).all()                   # another water took the alchemical slot
    ncmc.context = engine.runEngine(ncmc.context)
    moved = positions().value_in_unit(unit.nanometers)
    assert np.not_equal(swapped, moved[idx]).all()
    # the translated water oxygen sits inside the sphere around the selection's centre of mass (periodic distance)
    assert pdist(moved[idx[0]], com) <= 0.9 + 1e-6
    # rigid translation: the water geometry is unchanged, nothing else moved
    assert np.allclose(moved[idx[1]] - moved[idx[0]], swapped[1] - swapped[0], atol=1e-6)
    assert np.allclose(moved[idx[2]] - moved[idx[0]], swapped[2] - swapped[0], atol=1e-6)
    rest = np.setdiff1d(np.arange(len(start)), idx)
    assert np.array_equal(moved[rest], after_swap[rest])
    # in bounds: the work is untouched; pushed out of the sphere: afterMove forces rejection
    ncmc.context = move.afterMove(ncmc.context)
    assert ncmc.context._integrator.getGlobalVariableByName('protocol_work') == 0
    out = moved.copy()
    out[idx] = out[idx] - out[idx[0]] + (com + np.array([1.0, 0.0, 0.0]))    # 1.0 nm from the centre (< box / 2)
    ncmc.context.setPositions(out * unit.nanometers)
    ncmc.context = move.afterMove(ncmc.context)
    assert ncmc.context._integrator.getGlobalVariableByName('protocol_work') >= 999999


def test_water_translation_on_device_many_walkers(structure):
    """Every walker draws its own water and its own point of the sphere; a walker with no water in range keeps its
    coordinates through all three hooks (the reference's `go = False`, blues/moves.py:1003-1013)."""
    from blues_b200 import _native
    from blues_b200.moves import WaterTranslationMove
    from tests import gpu_checks as gc
    R = 6
    move = WaterTranslationMove(structure, protein_selection='(index 0) or (index 1)', radius=0.9 * unit.nanometers)
    _, _, topo, x0 = gc.load_case('tol_parm')
    eng = _native.Engine(topo, n_replicas=R, seed=11)
    lam_s, lam_e = gc.lambda_tables(10)
    eng.set_ncmc_integrator(300.0, 1.0, 0.002, 'H V R O R V H', 10, 1, 2.0, -1.0, lam_s, lam_e)
    x0 = np.asarray(x0, float)
    box = np.asarray(topo['box'], float).reshape(-1)[:3]
    idx = list(move.atom_indices)
    com = _selection_com(move, x0)
    for r in range(R):
        eng.set_positions(x0, r)
    desc = move._descriptor(_native.BL_MOVE_WATER_SWAP, with_waters=True)
    kind = desc.pop('kind'); desc.pop('step'); atoms = desc.pop('atoms')
    eng.apply_move(kind, atoms, None, **desc)
    from oracle import ncmc_oracle as orc
    # oracle: the candidates in residue order, choice = floor(u * n_inside) with the walker's first STREAM_MOVE draw
    inside = orc.waters_in_sphere(x0, box, move.water_residues, com, 0.9)
    partners = []
    for r in range(R):
        x = eng.get_positions(r)
        partner = [w for w in move.water_residues if np.array_equal(x0[w], x[idx])]
        assert len(partner) == 1
        u = orc.philox_uniform4(11, orc.STREAM_MOVE, r, 0, [1])[0][0]
        assert partner[0] == inside[min(int(u * len(inside)), len(inside) - 1)]
        partners.append(partner[0][0])
        d = x0[partner[0][0]] - com
        d -= box * np.round(d / box)
        assert np.linalg.norm(d) <= 0.9 + 1e-6
        assert np.array_equal(x[partner[0]], x0[idx])
    assert len(set(partners)) > 1                                        # independent draws per walker
    desc = move._descriptor(_native.BL_MOVE_WATER_TRANSLATE)
    kind = desc.pop('kind'); desc.pop('step'); atoms = desc.pop('atoms')
    xs = [eng.get_positions(r) for r in range(R)]
    eng.apply_move(kind, atoms, None, **desc)
    targets = []
    for r in range(R):
        x = eng.get_positions(r)
        d = x[idx[0]] - com
        assert np.linalg.norm(d) <= 0.9 + 1e-6                           # the target is centre + r * direction, unwrapped
        assert np.allclose(x[idx[1]] - x[idx[0]], xs[r][idx[1]] - xs[r][idx[0]], atol=1e-12)
        u = orc.philox_uniform4(11, orc.STREAM_MOVE, r, 1, [2])
        want = orc.water_translate(xs[r], box, idx, com.astype(np.float32).astype(float), 0.9, u[0][0], u[1][0], u[2][0])
        assert np.allclose(x, want, atol=1e-6)                # float32 centre: 1 ulp
        targets.append(tuple(np.round(x[idx[0]], 6)))
    assert len(set(targets)) == R
    # empty sphere: nothing is swapped, translated or flagged
    tiny = dict(move._descriptor(_native.BL_MOVE_WATER_SWAP, with_waters=True), radius=1e-4)
    kind = tiny.pop('kind'); tiny.pop('step'); atoms = tiny.pop('atoms')
    xs = [eng.get_positions(r) for r in range(R)]
    eng.apply_move(kind, atoms, None, **tiny)
    for k in (_native.BL_MOVE_WATER_TRANSLATE, _native.BL_MOVE_WATER_CHECK):
        dd = dict(move._descriptor(k), radius=1e-4)
        dd.pop('kind'); dd.pop('step'); a = dd.pop('atoms')
        eng.apply_move(k, a, None, **dd)
    for r in range(R):
        assert np.array_equal(eng.get_positions(r), xs[r])
        assert eng.get_global('protocol_work', r) == 0


def test_blues_run_with_water_translation_on_device(structure, tmp_path):
    """example_water.py's loop: WaterTranslationMove inside BLUESSimulation.run, hooks and midpoint move on the device."""
    from blues_b200.moves import WaterTranslationMove
    move = WaterTranslationMove(structure, protein_selection='(index 0) or (index 1)', radius=0.9 * unit.nanometers)
    engine = MoveEngine(move)
    systems = SystemFactory(structure, move.atom_indices, system_cfg())
    cfg = sim_cfg()
    cfg.update(nIter=2, nstepsNC=20, nstepsMD=4)
    simulations = SimulationFactory(systems, engine, cfg)
    for sim in (simulations.md, simulations.alch, simulations.ncmc):
        sim.minimizeEnergy(maxIterations=50)
    blues = BLUESSimulation(simulations)
    launches0 = simulations.ncmc.context._engine.launch_count()
    blues.run()
    assert blues.accept + blues.reject == 2
    assert simulations.ncmc.context._engine.launch_count() > launches0
    w = simulations.ncmc.context._integrator.getGlobalVariableByName('protocol_work')
    assert np.isfinite(w)


def test_md_leg_with_monte_carlo_barostat(structure):
    """`pressure` in the simulation config attaches a MonteCarloBarostat to the MD system only
    (blues/simulation.py:602-626, 781-785); the MD leg then attempts a volume move every `frequency` steps."""
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, system_cfg())
    cfg = sim_cfg()
    cfg['pressure'] = 1 * unit.atmospheres
    simulations = SimulationFactory(systems, MoveEngine(RandomLigandRotationMove(structure, 'LIG')), cfg)
    md, ncmc = simulations.md, simulations.ncmc
    assert type(md.system.getForces()[-1]).__name__ == 'MonteCarloBarostat'        # tests/test_simulation.py:241-250
    baro = md.context._barostat
    assert baro is not None and ncmc.context._barostat is None
    assert baro.frequency == 25 and baro.n_molecules == 321                        # toluene + 320 waters
    md.minimizeEnergy(maxIterations=100)
    md.context.setVelocitiesToTemperature(300 * unit.kelvin)
    box0 = np.diag(md.context.getState().getPeriodicBoxVectors(asNumpy=True).value_in_unit(unit.nanometers))
    e0 = md.context.getState(getEnergy=True).getPotentialEnergy()._value
    md.step(110)
    assert baro.total_attempted == 4                                               # steps 25, 50, 75, 100
    state = md.context.getState(getPositions=True, getEnergy=True)
    box1 = np.diag(state.getPeriodicBoxVectors(asNumpy=True).value_in_unit(unit.nanometers))
    assert np.isfinite(state.getPotentialEnergy()._value) and np.isfinite(e0)
    assert np.all(np.isfinite(state.getPositions(asNumpy=True)._value))
    if baro.total_accepted:
        assert not np.allclose(box0, box1, rtol=0, atol=1e-9)
    else:
        assert np.allclose(box0, box1, rtol=0, atol=1e-12)
    assert abs(np.prod(box1) / np.prod(box0) - 1.0) < 0.05                         # <= 4 moves of <= 1 % each
    assert np.allclose(box1 / box0, (box1 / box0)[0])                              # isotropic scaling
    # a rejected trial leaves positions and box exactly as they were
    x_before = md.context._engine.get_positions(0)
    b_before = md.context._engine.get_box()
    baro.volume_scale = 0.29 * np.prod(b_before)                                   # an absurd compression: rejected
    ok = baro.attempt(md.context._engine, uniforms=(0.0, 0.999999))
    assert ok is False
    assert np.array_equal(md.context._engine.get_box(), b_before)
    assert np.allclose(md.context._engine.get_positions(0), x_before, rtol=0, atol=0)


def test_frame_indices_reporter_without_an_interval_reporter(structure, tmp_path):
    """a reporter scheduled by explicit frame_indices (blues/reporters.py:362-367) must see every listed frame
    even when no interval-1 reporter forces single steps — the chunked stepper stops at the listed indices."""
    from blues_b200.reporters import NetCDF4Reporter
    from scipy.io import netcdf_file
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, system_cfg())
    cfg = sim_cfg()
    cfg.update(nstepsNC=6, nstepsMD=2)
    fname = os.path.join(str(tmp_path), 'frames.nc')
    rep = NetCDF4Reporter(fname, frame_indices=[1, 3, 6], protocolWork=True, alchemicalLambda=True)
    simulations = SimulationFactory(systems, MoveEngine(RandomLigandRotationMove(structure, 'LIG')), cfg,
                                    ncmc_reporters=[rep])
    b = BLUESSimulation(simulations)
    b._md_sim.minimizeEnergy(maxIterations=50)
    b._syncStatesMDtoNCMC()
    b._stepNCMC(6, 3)
    assert b._ncmc_sim.currentStep == 6
    rep.close() if hasattr(rep, 'close') else None
    nc = netcdf_file(fname, 'r', mmap=False)
    lam = np.array(nc.variables['alchemicalLambda'][:], float)
    nc.close()
    assert len(lam) == 3                                            # frames after steps 1, 3 and 6
    assert np.allclose(lam, [1 / 6.0, 3 / 6.0, 1.0], atol=1e-6)


class _NaNMove(RandomLigandRotationMove):
    """A host-path move that breaks one coordinate: the reference's flow then rejects the move and carries on."""

    def move(self, context):
        x = context.getState(getPositions=True).getPositions(asNumpy=True)
        if getattr(self, 'poison', False):
            x._value[self.atom_indices[0], 0] = np.nan
        context.setPositions(x)
        return context


def test_blues_iteration_survives_a_nan_move(structure):
    """after a NaN inside one NCMC leg `run()` must reject that move and keep iterating, as the
    reference does (blues/simulation.py:1082-1094, 1130-1140), instead of failing every later state query."""
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, system_cfg())
    cfg = sim_cfg()
    cfg.update(nIter=1, nstepsNC=6, nstepsMD=2)
    move = _NaNMove(structure, 'LIG', 3)
    simulations = SimulationFactory(systems, MoveEngine(move), cfg)
    b = BLUESSimulation(simulations)
    b._md_sim.minimizeEnergy(maxIterations=50)
    move.poison = True
    b.run()
    assert (b.accept, b.reject) == (0, 1)
    x = b._md_sim.context.getState(getPositions=True).getPositions(asNumpy=True)._value
    assert np.all(np.isfinite(x))
    move.poison = False
    b.run()
    assert b.accept + b.reject == 2
    assert np.isfinite(b._ncmc_sim.integrator.getGlobalVariableByName('protocol_work'))


def test_contexts_draw_distinct_seeds_unless_one_is_configured(structure):
    """seed 0 means a fresh seed per Context (OpenMM semantics); `seed` in the simulation config pins it."""
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, system_cfg())
    sims = SimulationFactory(systems, MoveEngine(RandomLigandRotationMove(structure, 'LIG')), sim_cfg())
    seeds = [s.integrator.getRandomNumberSeed() for s in (sims.md, sims.alch, sims.ncmc)]
    assert all(seeds) and len(set(seeds)) == 3
    cfg = sim_cfg()
    cfg['seed'] = 1234
    sims2 = SimulationFactory(systems, MoveEngine(RandomLigandRotationMove(structure, 'LIG')), cfg)
    assert [s.integrator.getRandomNumberSeed() for s in (sims2.md, sims2.alch, sims2.ncmc)] == [1234, 1235, 1236]
    v = [s.context.getState(getVelocities=True).getVelocities(asNumpy=True)._value for s in (sims2.md, sims2.alch)]
    assert not np.array_equal(v[0], v[1])


def _walker_sim(structure, n_walkers, seed=4321, **over):
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, system_cfg())
    cfg = sim_cfg()
    cfg.update(nIter=2, nstepsNC=10, nstepsMD=4, nReplicas=n_walkers, seed=seed)
    cfg.update(over)
    simulations = SimulationFactory(systems, MoveEngine(RandomLigandRotationMove(structure, 'LIG')), cfg)
    for sim in (simulations.md, simulations.alch, simulations.ncmc):
        sim.minimizeEnergy(maxIterations=100)
    return BLUESSimulation(simulations), simulations


def test_many_walker_blues_iteration_through_the_api(structure):
    """`simulation: nReplicas` runs R walkers through BLUESSimulation.run — device-to-device sync,
    NCMC with the on-device move, per-walker correction + Metropolis test, accepted walkers copied to the MD context —
    and every piece agrees with the single-walker bookkeeping recomputed on the host from the same states."""
    R = 4
    b, sims = _walker_sim(structure, R)
    md, alch, ncmc = (s.context._engine for s in (sims.md, sims.alch, sims.ncmc))
    assert ncmc.n_replicas == R and md.n_replicas == R
    sims.md.context.setVelocitiesToTemperature(300 * unit.kelvin)
    sims.md.step(5)                                                  # walkers decorrelate (independent noise)
    x_md0 = [md.get_positions(r) for r in range(R)]
    assert not np.allclose(x_md0[0], x_md0[1])
    e_md0 = md.get_energy(True, False)[0]
    rec = b._iterateWalkers(10, 5, 4, 300 * unit.kelvin)
    assert rec['accepted'].shape == (R,) and np.all(np.isfinite(rec['work_kT']))
    assert b.accept + b.reject == R
    assert np.std(rec['work_kT']) > 0
    # Metropolis rule per walker, with the engine's own uniform draws
    want = (rec['log_accept'] > rec['log_u'])
    assert np.array_equal(rec['accepted'], want)
    # log_accept = -work/kT + correction (blues/simulation.py:1130-1136)
    assert np.allclose(rec['log_accept'], -rec['work_kT'] + rec['correction'], rtol=0, atol=1e-9)
    # the integrator was reset, the MD leg ran
    assert ncmc.get_global('step') == 0 and ncmc.get_global('protocol_work') == 0
    x_md1 = [md.get_positions(r) for r in range(R)]
    for r in range(R):
        assert not np.array_equal(x_md1[r], x_md0[r])
    # a full run over two iterations keeps per-walker statistics
    b.run(nIter=2)
    assert len(b.walker_history) == 2 and len(b.walker_history[0]['walker']) == R
    assert 0.0 <= b.acceptRatio <= 1.0


def test_many_walker_correction_and_accept_copy_match_host_recomputation(structure):
    """The per-walker alchemical correction equals the four-energy formula evaluated walker by walker through the
    single-walker calls, and exactly the accepted walkers' MD positions are replaced by the NCMC end positions."""
    R = 3
    b, sims = _walker_sim(structure, R, seed=99)
    md, alch, ncmc = (s.context._engine for s in (sims.md, sims.alch, sims.ncmc))
    sims.md.context.setVelocitiesToTemperature(300 * unit.kelvin)
    sims.md.step(3)
    x0 = [md.get_positions(r) for r in range(R)]
    e_md0 = md.get_energy(True, False)[0].copy()
    kT = sims.ncmc.integrator.kT.value_in_unit(unit.kilojoules_per_mole)

    class Spy(object):
        pass
    spy = Spy()
    orig = ncmc.accept_reject

    def accept_reject(correction=None):
        spy.x1 = [ncmc.get_positions(r) for r in range(R)]
        spy.e_nc1 = ncmc.get_energy(True, False)[0].copy()
        spy.corr = np.array(correction, float)
        out = orig(correction)
        spy.acc = out[0].astype(bool)
        return out
    ncmc.accept_reject = accept_reject
    # force walker 1 to accept, walker 2 to reject, walker 0 by its own work
    rec = b._iterateWalkers(10, 5, 0 + 1, 300 * unit.kelvin)
    for r in range(R):
        ncmc.set_positions(x0[r], r)
    ncmc.reset_ncmc()
    e_nc0 = ncmc.get_energy(True, False)[0]
    for r in range(R):
        alch.set_positions(spy.x1[r], r)
    e_alch1 = alch.get_energy(True, False)[0]
    want = -(e_nc0 - e_md0 + e_alch1 - spy.e_nc1) / kT
    assert np.allclose(spy.corr, want, rtol=0, atol=2e-3)
    # accepted walkers carry the NCMC end positions into the MD leg (which then ran 1 step); rejected keep x0
    x_md = [md.get_positions(r) for r in range(R)]
    for r in range(R):
        ref = spy.x1[r] if spy.acc[r] else x0[r]
        assert np.max(np.abs(x_md[r] - ref)) < 0.05                 # one MD step away from the right starting point
        other = x0[r] if spy.acc[r] else spy.x1[r]
        assert np.max(np.abs(x_md[r] - other)) > np.max(np.abs(x_md[r] - ref))


def sidechain_move_contract(golden_dir=GOLDEN):
    """``blues/tests/test_sidechain.py:30-88`` (which upstream can only run with an OpenEye licence): divaline in vacuum,
    one rotatable bond in valine with 11 listed atoms, and a move that displaces everything but the two axis atoms."""
    from blues_b200.simulation import SystemFactory, SimulationFactory
    from blues_b200.moves import SideChainMove, MoveEngine
    from blues_b200 import system as sysmod
    struct = Structure.load_npz(os.path.join(golden_dir, 'vac_divaline.npz'))
    sidechain = SideChainMove(struct, [1])
    engine = MoveEngine(sidechain)
    engine.selectMove()
    vals = [v for v in sidechain.rot_atoms[1].values()][0]
    assert len(vals) == 11
    assert len(sidechain.rot_bonds) == 1
    systems = SystemFactory(struct, sidechain.atom_indices, {'nonbondedMethod': sysmod.NoCutoff, 'constraints': sysmod.HBonds})
    cfg = {'dt': 0.002 * unit.picoseconds, 'friction': 1 / unit.picoseconds, 'temperature': 300 * unit.kelvin,
           'nIter': 1, 'nstepsMD': 1, 'nstepsNC': 4, 'platform': 'CUDA',
           'alchemical_functions': {
               'lambda_sterics': 'step(0.199999-lambda) + step(lambda-0.2)*step(0.8-lambda)*abs(lambda-0.5)*1/0.3 + step(lambda-0.800001)',
               'lambda_electrostatics': 'step(0.2-lambda)- 1/0.2*lambda*step(0.2-lambda) + 1/0.2*(lambda-0.8)*step(lambda-0.8)'}}
    simulations = SimulationFactory(systems, engine, cfg)
    atom_indices = vals
    before = simulations.ncmc.context.getState(getPositions=True).getPositions(asNumpy=True)[atom_indices, :]
    simulations.ncmc.context = engine.runEngine(simulations.ncmc.context)
    after = simulations.ncmc.context.getState(getPositions=True).getPositions(asNumpy=True)[atom_indices, :]
    b, a = np.asarray(before._value), np.asarray(after._value)
    assert np.not_equal(b, a)[2:, :].all()                 # every rotated atom moved ...
    assert np.allclose(b[:2], a[:2], atol=1e-12)           # ... the two axis atoms did not
    d0 = np.linalg.norm(b[2:] - b[1], axis=1)
    d1 = np.linalg.norm(a[2:] - a[1], axis=1)
    assert np.allclose(d0, d1, atol=1e-9)                  # a rigid rotation about the bond
    ax = (b[0] - b[1]) / np.linalg.norm(b[0] - b[1])
    assert np.allclose((b[2:] - b[1]) @ ax, (a[2:] - a[1]) @ ax, atol=1e-9)


@pytest.mark.gpu
def test_sidechain_move():
    sidechain_move_contract()


def test_example_sidechain_script(tmp_path, monkeypatch):
    """examples/example_sidechain.py on the CUDA engine (the reference's examples/example_sidechain.py:1-37 with its
    analysis step): SideChainMove + NCMC on the valine dipeptide, MD frames into NetCDF, chi1 of every frame read back."""
    import importlib.util
    import shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ex = tmp_path / 'examples'
    shutil.copytree(os.path.join(root, 'examples'), str(ex))
    os.makedirs(str(tmp_path / 'tests' / 'golden'))
    shutil.copy(os.path.join(GOLDEN, 'vac_divaline.npz'), str(tmp_path / 'tests' / 'golden'))
    monkeypatch.chdir(str(ex))
    spec = importlib.util.spec_from_file_location('example_sidechain', str(ex / 'example_sidechain.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    blues = mod.sidechain('sidechain_b200.yml', nIter=3, nstepsNC=20, nstepsMD=500)
    assert blues.accept + blues.reject == 3
    chi = np.asarray(blues.dihedrals)
    assert chi.shape == (6, 1) and np.all(np.isfinite(chi)) and np.all(np.abs(chi) <= np.pi + 1e-6)
    assert os.path.exists(str(ex / 'divaline-b200-ncmc.nc'))
