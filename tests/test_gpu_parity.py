"""GPU parity tests: the CUDA path (through the C ABI) against the float64 oracle on identical inputs.

Tolerances (BASELINE.json north_star): forces / energies within 1e-4 relative in mixed precision; neighbour and
exclusion sets bit-exact; protocol work step for step (the engine and the oracle share the Philox noise stream).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests import gpu_checks as gc   # noqa: E402

FORCE_TOL = 1e-4      # max over atoms of |dF| / max(|F_atom|, rms|F|)
ENERGY_TOL = 1e-4     # relative, on the total potential energy


@pytest.mark.parametrize('name', ['vac_divaline', 'tol_parm', 'wat_divaline', 't4l_surrogate'])
def test_md_system_forces_and_energy_terms(name):
    out = gc.compare_forces(name, alchemical=False)
    assert out['energy_rel'] < ENERGY_TOL
    assert out['force_max_rel'] < FORCE_TOL
    for k, (a, b) in out['terms'].items():
        assert abs(a - b) <= 1e-5 * max(abs(b), 1.0) + 1e-3, (k, a, b)


@pytest.mark.parametrize('name', ['vac_divaline', 'tol_parm', 'wat_divaline'])
@pytest.mark.parametrize('lam_index', [0, 3, 7, 10, 15, 20])
def test_alchemical_system_at_several_lambda(name, lam_index):
    out = gc.compare_forces(name, alchemical=True, lam_index=lam_index, nsteps=10)
    assert out['energy_rel'] < ENERGY_TOL
    assert out['force_max_rel'] < FORCE_TOL
    for k in ('alch_sterics', 'alch_electrostatics', 'alch_exceptions'):
        a, b = out['terms'][k]
        assert abs(a - b) <= 1e-6 * max(abs(b), 1.0) + 1e-5, (k, a, b)


def test_t4l_alchemical_midpoint():
    out = gc.compare_forces('t4l_surrogate', alchemical=True, lam_index=5000, nsteps=5000)
    assert out['energy_rel'] < ENERGY_TOL and out['force_max_rel'] < FORCE_TOL


@pytest.mark.parametrize('name', ['vac_divaline', 'tol_parm', 'wat_divaline', 't4l_surrogate'])
def test_neighbour_and_exclusion_sets_bit_exact(name):
    out = gc.compare_neighbors(name)
    assert out['duplicates'] == 0
    # a pair may differ only if its float64 distance is within float32 rounding of the cutoff
    assert out['only_engine'] + out['only_oracle'] == 0 or out['edge'] < 5e-7, out
    assert out['n_engine'] == out['n_oracle'] or out['edge'] < 5e-7


def test_tiled_system_above_65535_atoms_uses_32bit_list_indices():
    # watDivaline tiled 3 x 3 x 3 = 69 957 atoms: the neighbour rows switch from uint16 to int32 entries, the PME grid
    # and the cell grid grow; forces / energy against the oracle's C twin, neighbour set against the numpy oracle
    out = gc.compare_tiled('wat_divaline', (3, 3, 3))
    assert out['n_atoms'] == 69957
    assert out['duplicates'] == 0
    # coordinates up to ~9 nm in float32: a pair may differ only within 5e-6 nm of the cutoff
    assert out['only_engine'] + out['only_oracle'] == 0 or out['edge'] < 5e-6, out
    assert out['energy_rel'] < ENERGY_TOL and out['force_max_rel'] < FORCE_TOL, out


@pytest.mark.parametrize('name,kw', [('wat_divaline', dict(steps=80, dt=0.002)),
                                     ('t4l_surrogate', dict(steps=60, dt=0.004, minimize=60))])
def test_neighbour_sets_stay_exact_through_prunes_and_rebuilds(name, kw):
    # dual Verlet lists: inner list pruned from the outer one every few steps, cell search every ~10; after hot dynamics
    # the list in use must still contain every pair inside the cutoff, for every walker of a batched context
    for out in gc.compare_neighbors_dynamic(name, n_replicas=2, **kw):
        assert out['rebuilds'] >= 3, out
        assert out['duplicates'] == 0
        assert out['only_engine'] + out['only_oracle'] == 0 or out['edge'] < 5e-6, out


@pytest.mark.parametrize('name,kw', [('vac_divaline', {}), ('tol_parm', dict(minimize=True)),
                                     ('vac_divaline', dict(splitting='V H R O R H V')),
                                     ('vac_divaline', dict(splitting='R V O H O V R', nsteps=8)),
                                     ('vac_divaline', dict(nprop=3, prop_lambda=0.3, nsteps=10))])
def test_noisy_trajectory_and_work_step_for_step(name, kw):
    kw = dict(kw)
    nsteps = kw.pop('nsteps', 6)
    out = gc.compare_trajectory(name, nsteps=nsteps, **kw)
    assert out['dv0'] < 1e-9                      # identical Maxwell-Boltzmann draw + velocity constraints
    for r in out['rows']:
        assert r['dx'] < 5e-6 and r['dv'] < 5e-4, r
        assert abs(r['work_engine'] - r['work_oracle']) < 1e-4 * max(1.0, abs(r['work_oracle'])), r


def test_chunked_equals_single_steps_and_graph_equals_direct():
    """Device-resident chunks, per-step calls, graph replay and direct launches give the same trajectory."""
    res = []
    for mode in ('chunk', 'single', 'nograph'):
        eng, orc, topo = gc.make_ncmc_pair('vac_divaline', nsteps=10, seed=21)
        if mode == 'nograph':
            eng.use_graphs(False)
        if mode == 'single':
            for _ in range(10):
                eng.ncmc_run(1)
        else:
            eng.ncmc_run(10)
        res.append((eng.get_positions(0), eng.get_global('protocol_work')))
        eng.close()
    for x, w in res[1:]:
        assert np.array_equal(x, res[0][0]) and w == res[0][1]      # fixed-point accumulation → bitwise equal


def test_rotation_move_on_device_matches_oracle_and_external_work():
    from oracle import ncmc_oracle as orc
    from blues_b200 import _native
    eng, o, topo = gc.make_ncmc_pair('vac_divaline', nsteps=10, seed=5)
    atoms = np.arange(16, 35)
    masses = np.linspace(1.0, 12.0, len(atoms))
    x0 = eng.get_positions(0)
    eng.apply_move(_native.BL_MOVE_ROTATE, atoms, masses)
    x1 = eng.get_positions(0)
    u0, u1, u2, _ = orc.philox_uniform4(5, orc.STREAM_MOVE, 0, 0, [0])
    R = orc.rotation_matrix_from_quaternion(orc.quaternion_from_uniforms(u0[0], u1[0], u2[0]))
    ref = orc.rotate_ligand(x0, atoms, masses, R)
    assert np.max(np.abs(x1 - ref)) < 1e-6
    assert np.array_equal(x1[:16], x0[:16])
    eng.close()
    # the move inside the protocol: work bookkeeping equals the oracle's with the same rotation at moveStep
    eng, o, topo = gc.make_ncmc_pair('vac_divaline', nsteps=10, seed=5)
    eng.ncmc_run(10, dict(kind=_native.BL_MOVE_ROTATE, step=5, atoms=atoms, masses=masses))
    o.step(5)
    o.x = orc.rotate_ligand(o.x, atoms, masses, R)
    o.step(5)
    assert eng.get_global('step') == 10 and eng.get_global('lambda') == pytest.approx(1.0)
    assert np.max(np.abs(eng.get_positions(0) - o.x)) < 5e-6
    assert eng.get_global('protocol_work') == pytest.approx(o.g['protocol_work'], rel=1e-4, abs=1e-4)
    eng.close()


def test_external_work_for_host_side_coordinate_change():
    """blues/integrators.py:184-191: a coordinate change between steps enters protocol_work as E_after - E_before."""
    from oracle import ncmc_oracle as orc
    eng, o, topo = gc.make_ncmc_pair('vac_divaline', nsteps=10, seed=9)
    eng.ncmc_run(3)
    o.step(3)
    x = eng.get_positions(0)
    x[20] += np.array([0.01, -0.02, 0.015])
    eng.set_positions(x)
    o.x = x.copy()
    eng.ncmc_run(2)
    o.step(2)
    assert eng.get_global('protocol_work') == pytest.approx(o.g['protocol_work'], rel=1e-4, abs=1e-4)
    assert eng.get_global('unperturbed_pe') == pytest.approx(o.g['unperturbed_pe'], rel=1e-6)
    eng.close()


def test_work_distribution_over_walkers_matches_oracle():
    """Fixed-seed ensemble: per-walker protocol work equals the oracle's walker by walker (same noise stream),
    hence the distributions are indistinguishable (two-sample KS p > 0.05)."""
    from scipy.stats import ks_2samp
    from oracle import ncmc_oracle as orc
    from blues_b200 import _native
    R, nsteps = 48, 10
    s, system, topo, x = gc.load_case('vac_divaline', True)
    ls, le = gc.lambda_tables(nsteps)
    eng = _native.Engine(topo, n_replicas=R, seed=77)
    eng.set_ncmc_integrator(300.0, 1.0, 0.001, 'H V R O R V H', nsteps, 1, 0.2, 0.8, ls, le)
    eng.set_positions(x)
    eng.velocities_to_temperature(300.0)
    eng.ncmc_run(nsteps)
    w_gpu = np.array([eng.get_global('protocol_work', r) for r in range(R)])
    acc, logp, logu = eng.accept_reject()
    w_cpu = []
    for r in range(R):
        o = orc.NCMCOracle(topo, gc.DEFAULT_FUNCS, 'H V R O R V H', 300.0, 1.0, 0.001, nsteps, 1, 0.3, 77, r)
        o.x = x.copy()
        o.set_velocities_to_temperature(300.0, 0)
        o.step(nsteps)
        w_cpu.append(o.g['protocol_work'])
        lu = np.log(orc.philox_uniform4(77, orc.STREAM_ACCEPT, r, 0, [0])[0][0])
        assert logu[r] == pytest.approx(lu, rel=1e-12)
        assert bool(acc[r]) == orc.metropolis_accept(o.log_acceptance_probability(), 0.0, lu) or \
            abs(o.log_acceptance_probability() - lu) < 1e-3
    w_cpu = np.array(w_cpu)
    assert np.std(w_gpu) > 0                                  # walkers really are independent
    assert np.max(np.abs(w_gpu - w_cpu)) < 1e-3 * max(1.0, np.max(np.abs(w_cpu)))
    assert ks_2samp(w_gpu, w_cpu).pvalue > 0.05
    eng.close()


def test_md_leg_matches_oracle():
    from oracle import ncmc_oracle as orc
    from blues_b200 import _native
    s, system, topo, x = gc.load_case('vac_divaline', False)
    eng = _native.Engine(topo, n_replicas=1, seed=13)
    eng.set_langevin_integrator(300.0, 1.0, 0.002, 1e-10)
    eng.set_positions(x)
    eng.velocities_to_temperature(300.0)
    o = orc.LangevinMDOracle(topo, 300.0, 1.0, 0.002, 13, 0)
    o.x = x.copy()
    o.v = eng.get_velocities(0)
    eng.md_run(5)
    o.step(5)
    assert np.max(np.abs(eng.get_positions(0) - o.x)) < 5e-6
    eng.close()


def test_full_size_properties_t4l():
    """Size-independent properties at the BASELINE configuration: Newton's third law on the direct-space sum,
    constraint residuals after integration, energy bookkeeping consistency."""
    from blues_b200 import _native
    s, system, topo, x = gc.load_case('t4l_surrogate', True)
    ls, le = gc.lambda_tables(5000)
    eng = _native.Engine(topo, n_replicas=2, seed=3)
    eng.set_ncmc_integrator(300.0, 1.0, 0.004, 'H V R O R V H', 5000, 1, 0.2, 0.8, ls, le)
    eng.set_positions(x)
    eng.minimize(30, 10.0)
    eng.velocities_to_temperature(300.0)
    eng.ncmc_run(40)
    for r in range(2):
        xr = eng.get_positions(r)
        c = topo['constraints']
        d = np.linalg.norm(xr[c[:, 0]] - xr[c[:, 1]], axis=1)
        assert np.max(np.abs(d - topo['constraint_d']) / topo['constraint_d']) < 1e-7
        assert np.isfinite(eng.get_global('protocol_work', r))
    assert eng.get_global('step') == 40 and eng.get_global('lambda') == pytest.approx(40 / 5000)
    assert not np.allclose(eng.get_positions(0), eng.get_positions(1))       # independent noise per walker
    terms = eng.get_energy_terms(0)
    ep, ek = eng.get_energy()
    assert sum(terms.values()) == pytest.approx(ep[0], rel=1e-9)
    F = eng.get_forces(0)
    assert np.all(np.isfinite(F))
    eng.close()


def test_alchemical_run_is_bitwise_reproducible():
    """Fixed-point accumulation makes every sum independent of the order in which threads arrive; the one list that is
    filled through an atomic cursor (the alchemical pair list) is sorted before use.  Two walkers with the same start
    relax to bitwise equal energies, and two engines with the same seed produce bitwise equal trajectories and work."""
    from blues_b200 import _native
    s, system, topo, x = gc.load_case('t4l_surrogate', True)
    ls, le = gc.lambda_tables(5000)
    outs = []
    for rep in range(2):
        eng = _native.Engine(topo, n_replicas=2, seed=5)
        eng.set_ncmc_integrator(300.0, 1.0, 0.004, 'H V R O R V H', 5000, 1, 0.2, 0.8, ls, le)
        eng.set_positions(x)
        eng.minimize(20, 10.0)
        ep, _ = eng.get_energy()
        assert ep[0] == ep[1]
        eng.velocities_to_temperature(300.0)
        eng.ncmc_run(40)
        outs.append((eng.get_positions(0), eng.get_positions(1), eng.get_global('protocol_work', 0),
                     eng.get_global('protocol_work', 1)))
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2] and outs[0][3] == outs[1][3]


def test_madelung_constant_of_rock_salt_on_the_engine():
    """Literature anchor (tests/test_oracle.py has the same check for the oracle): lattice energy per ion pair of NaCl
    = -1.747565 e^2 / (4 pi eps0 r0), from the engine's erfc pair sum + smooth PME + self term; no net force on an ion."""
    from blues_b200 import _native
    from tests.test_oracle import _rock_salt
    a0, cells = 0.564, 4
    r0 = 0.5 * a0
    expect = -1.747565 * 138.935456 / r0
    topo, x = _rock_salt(cells, a0, 5e-4, cutoff=0.9)
    eng = _native.Engine(topo, n_replicas=1, seed=1)
    eng.set_langevin_integrator(300.0, 1.0, 0.002)
    eng.set_positions(x)
    ep, _ = eng.get_energy()
    f = eng.get_forces()
    eng.close()
    assert ep[0] / (topo['n_atoms'] // 2) == pytest.approx(expect, rel=2e-3)
    assert np.max(np.abs(f)) < 2e-3 * abs(expect) / r0


# ---- round 2: dynamic parity on solvated PME systems, frozen / restrained systems, full M1 protocol ensemble ----------

def test_long_noisy_trajectory_on_solvated_pme_system():
    """50 noisy NCMC steps of the solvated dipeptide (PME, SETTLE waters, 2591 atoms) against the oracle's C twin: the
    float32 pair / PME arithmetic must not pull the trajectory or the accumulated work away from the float64 reference."""
    out = gc.compare_trajectory_c('wat_divaline', nsteps=50, stride=10, dt=0.002, seed=17, minimize=100)
    assert out['dv0'] < 1e-9
    for r in out['rows']:
        assert r['dx'] < 2e-4 and r['dv'] < 2e-2, r
        assert abs(r['work_engine'] - r['work_oracle']) < 2e-3 * max(1.0, abs(r['work_oracle'])) + 2e-3, r


@pytest.mark.parametrize('name,kw', [('tol_parm', dict(freeze_beyond_nm=0.6, minimize=100)),
                                     ('wat_divaline', dict(freeze_beyond_nm=0.5, minimize=100)),
                                     ('wat_divaline', dict(restrain='(@CA,C,N)', minimize=100))])
def test_frozen_and_restrained_systems_match_oracle(name, kw):
    """freeze_radius / freeze_atoms zero the masses (blues/simulation.py:364-480), restrain_positions adds the harmonic
    CustomExternalForce (:319-362): forces, a 20-step noisy trajectory and the work against the oracle; frozen atoms keep
    their coordinates bit for bit and carry zero velocity."""
    from oracle.ncmc_oracle import ForceField
    s, system, topo, x = gc.load_case(name, True, **{k: v for k, v in kw.items() if k != 'minimize'})
    mass = np.asarray(topo['mass'], float)
    if 'freeze_beyond_nm' in kw:
        assert 0 < np.count_nonzero(mass == 0) < len(mass) - len(gc.CASES[name]['alch'])
    else:
        assert len(topo['restraint_atoms']) > 0
    cmp_ = gc.compare_forces(name, alchemical=True, lam_index=3, nsteps=10, **{k: v for k, v in kw.items() if k != 'minimize'})
    assert cmp_['energy_rel'] < ENERGY_TOL and cmp_['force_max_rel'] < FORCE_TOL
    a, b = cmp_['terms']['restraint']
    assert abs(a - b) <= 1e-6 * max(abs(b), 1.0) + 1e-6
    out = gc.compare_trajectory_c(name, nsteps=20, stride=5, dt=0.002, seed=23, **kw)
    for r in out['rows']:
        assert r['dx'] < 5e-5 and r['dv'] < 5e-3, r
        assert abs(r['work_engine'] - r['work_oracle']) < 1e-3 * max(1.0, abs(r['work_oracle'])) + 1e-3, r
    fz = out['frozen']
    if len(fz):
        assert np.array_equal(out['x_engine'][fz], out['x0'][fz])


def test_t4l_freeze_radius_variant_runs_with_frozen_atoms_untouched():
    """The example's default variant (examples/rotmove_cuda.yml:42-45, docs/BLUES_tutorial.ipynb:718): freeze_radius 5 A
    leaves 275 mobile atoms of 22 340; the engine integrates those, keeps the other 22 065 bit for bit and their
    constraints satisfied, and its forces still match the oracle's on every atom."""
    from blues_b200 import _native, unit
    from blues_b200.simulation import SystemFactory
    from oracle.c_oracle import COracle
    s, system, topo, x = gc.load_case('t4l_surrogate', False)
    system = SystemFactory.freeze_radius(s, system, freeze_distance=5.0 * unit.angstroms, freeze_center=':LIG',
                                         freeze_solvent=':HOH,NA,CL,Cl-')
    from blues_b200.alchemy import AbsoluteAlchemicalFactory, AlchemicalRegion
    system = AbsoluteAlchemicalFactory().create_alchemical_system(
        system, AlchemicalRegion(alchemical_atoms=gc.CASES['t4l_surrogate']['alch']))
    topo = system.flatten()
    mass = np.asarray(topo['mass'], float)
    assert np.count_nonzero(mass == 0) == 22065
    ls, le = gc.lambda_tables(5000)
    eng = _native.Engine(topo, n_replicas=2, seed=3)
    eng.set_ncmc_integrator(300.0, 1.0, 0.004, 'H V R O R V H', 5000, 1, 0.2, 0.8, ls, le)
    eng.set_positions(x)
    eng.minimize(30, 10.0)
    x0 = eng.get_positions(0)
    F = eng.get_forces(0)
    Eo, Fo = COracle(topo, ls, le, nsteps_neq=5000).energy_forces(x0, ls[0], le[0])[:2]
    fmax, frms = gc.rel_force_error(F, Fo)
    assert fmax < FORCE_TOL
    eng.velocities_to_temperature(300.0)
    v = eng.get_velocities(0)
    assert np.all(v[mass == 0] == 0.0) and np.any(v[mass > 0] != 0.0)
    eng.ncmc_run(40)
    for r in range(2):
        xr = eng.get_positions(r)
        assert np.array_equal(xr[mass == 0], x0[mass == 0])
        assert not np.array_equal(xr[mass > 0], x0[mass > 0])
        c = topo['constraints']
        if len(c):
            d = np.linalg.norm(xr[c[:, 0]] - xr[c[:, 1]], axis=1)
            assert np.max(np.abs(d - topo['constraint_d']) / topo['constraint_d']) < 1e-6
        assert np.isfinite(eng.get_global('protocol_work', r))
    eng.close()


@pytest.mark.parametrize('name,frz,rotate', [('tol_parm', 0.6, True), ('wat_divaline', 0.5, False)])
def test_frozen_fast_paths_do_not_change_a_single_bit(name, frz, rotate):
    """Skipping frozen rows in the pair kernel, frozen clusters in the integrator, frozen atoms in the PME gather and
    keeping the frozen atoms' share of the PME charge grid (fixed-point sums: share + share == sum over all) must leave
    the mobile atoms' trajectory and the work bit for bit what the engine computes with all of that switched off
    (BLUES_B200_SKIP_FROZEN=0) — across a rotation move, a host coordinate upload (which invalidates the stored grid)
    and an energy query in between.  (No rotation for the dipeptide: re-coupling it on top of the water within 20 steps
    blows the walker up whichever path computes it.)"""
    import os
    from blues_b200 import _native
    s, system, topo, x = gc.load_case(name, True, freeze_beyond_nm=frz)
    mass = np.asarray(topo['mass'], float)
    alch = np.asarray(gc.CASES[name]['alch'], np.int32)
    assert np.count_nonzero(mass == 0) > 0 and np.all(mass[alch] > 0)
    ls, le = gc.lambda_tables(40)
    res = {}
    old = os.environ.get('BLUES_B200_SKIP_FROZEN')
    try:
        for flag in ('1', '0'):
            os.environ['BLUES_B200_SKIP_FROZEN'] = flag
            eng = _native.Engine(topo, n_replicas=2, seed=5)
            eng.set_ncmc_integrator(300.0, 1.0, 0.002, 'H V R O R V H', 40, 1, 0.2, 0.8, ls, le)
            eng.set_positions(x)
            eng.minimize(40, 10.0)
            eng.velocities_to_temperature(300.0)
            # the rotation at lambda = 0.5, where the ligand is decoupled (blues/simulation.py:1039-1098)
            eng.ncmc_run(24, move=dict(kind=_native.BL_MOVE_ROTATE, step=20, atoms=alch, masses=mass[alch]) if rotate else None)
            e_mid = eng.get_energy()[0].copy()
            xm = eng.get_positions(1)
            xm[mass > 0] += 1e-3                      # host upload for one walker: every stored grid share is dropped
            eng.set_positions(xm, replica=1)
            eng.ncmc_run(16)
            res[flag] = dict(x=[eng.get_positions(r) for r in range(2)], v=[eng.get_velocities(r) for r in range(2)],
                             w=[eng.get_global('protocol_work', r) for r in range(2)], e=e_mid,
                             e_end=eng.get_energy()[0].copy(), f=eng.get_forces(0),
                             pairs=np.sort(np.asarray(eng.neighbor_pairs(1))))
            eng.close()
    finally:
        if old is None:
            os.environ.pop('BLUES_B200_SKIP_FROZEN', None)
        else:
            os.environ['BLUES_B200_SKIP_FROZEN'] = old
    a, b = res['1'], res['0']
    for r in range(2):
        assert np.array_equal(a['x'][r], b['x'][r]) and np.array_equal(a['v'][r], b['v'][r])
        assert a['w'][r] == b['w'][r] and np.isfinite(a['w'][r])
    assert np.array_equal(a['e'], b['e']) and np.array_equal(a['e_end'], b['e_end'])
    # a host force query after force-only launches re-evaluates every row (forces_partial): frozen atoms get their forces
    assert np.array_equal(a['f'], b['f']) and np.any(a['f'][mass == 0] != 0.0)
    assert len(a['pairs']) > 0 and np.array_equal(a['pairs'], b['pairs'])


def test_m1_full_protocol_ensemble_work_and_acceptance():
    """North-star criterion 3 on M1 = BASELINE configs[0]: 64 walkers of toluene in TIP3P (TOL-parm, PME, HBonds),
    the full nstepsNC = 100 protocol with the rotation at moveStep = 50, the alchemical correction and the Metropolis
    test (blues/simulation.py:1039-1166).  Engine walkers and oracle walkers share seeds, noise streams and rotations:
    per-walker work agrees while trajectories coincide, the work distributions are KS-indistinguishable and the
    acceptance counts are consistent (binomial test)."""
    from scipy.stats import ks_2samp, binomtest
    from oracle import ncmc_oracle as orc
    from oracle.c_oracle import COracle
    from blues_b200 import _native
    R, nsteps, move_step, seed, dt, T = 64, 100, 50, 4242, 0.002, 300.0
    s, system_md, topo_md, x = gc.load_case('tol_parm', False)
    _, _, topo, _ = gc.load_case('tol_parm', True)
    ls, le = gc.lambda_tables(nsteps)
    atoms = np.arange(15)
    masses = np.asarray(topo_md['mass'], float)[atoms]
    kT = orc.KB * T
    eng = _native.Engine(topo, n_replicas=R, seed=seed)
    eng.set_ncmc_integrator(T, 1.0, dt, 'H V R O R V H', nsteps, 1, 0.2, 0.8, ls, le)
    eng.set_positions(x)
    eng.minimize(200, 10.0)
    x0 = eng.get_positions(0)
    eng.set_positions(x0)                                    # every walker starts from the same relaxed frame
    eng.velocities_to_temperature(T)
    md = _native.Engine(topo_md, n_replicas=R, seed=seed + 1)
    md.set_langevin_integrator(T, 1.0, dt)
    md.set_positions(x0)
    e_md0 = md.get_energy(True, False)[0]
    e_nc0 = eng.get_energy(True, False)[0]
    eng.ncmc_run(nsteps, dict(kind=_native.BL_MOVE_ROTATE, step=move_step, atoms=atoms, masses=masses))
    w_gpu = np.array([eng.get_global('protocol_work', r) for r in range(R)])
    e_nc1 = eng.get_energy(True, False)[0]
    x1 = [eng.get_positions(r) for r in range(R)]
    for r in range(R):
        md.set_positions(x1[r], r)
    e_md1 = md.get_energy(True, False)[0]
    corr_gpu = -(e_nc0 - e_md0 + e_md1 - e_nc1) / kT
    acc_gpu, logp_gpu, logu = eng.accept_reject(corr_gpu)
    # oracle walkers
    o_md = COracle(topo_md)
    w_cpu, corr_cpu, acc_cpu, dx = [], [], [], []
    for r in range(R):
        c = COracle(topo, ls, le, 'H V R O R V H', T, 1.0, dt, nsteps, 1, 0.2, 0.8, seed=seed, replica=r)
        c.set_state(x0)
        c.velocities_to_temperature(T)
        e0n = c.energy_forces(x0, ls[0], le[0])[0]
        e0m = o_md.energy_forces(x0)[0]
        c.step(move_step)
        u0, u1, u2, _ = orc.philox_uniform4(seed, orc.STREAM_MOVE, r, 0, [0])
        Rm = orc.rotation_matrix_from_quaternion(orc.quaternion_from_uniforms(u0[0], u1[0], u2[0]))
        c.x = np.ascontiguousarray(orc.rotate_ligand(c.x, atoms, masses, Rm))
        c.step(nsteps - move_step)
        w_cpu.append(c.get('protocol_work'))
        e1n = c.energy_forces(c.x, ls[-1], le[-1])[0]
        e1m = o_md.energy_forces(c.x)[0]
        corr_cpu.append(orc.alchemical_correction(e0n, e0m, e1m, e1n, kT))
        lu = np.log(orc.philox_uniform4(seed, orc.STREAM_ACCEPT, r, 0, [0])[0][0])
        assert logu[r] == pytest.approx(lu, rel=1e-12)
        acc_cpu.append(orc.metropolis_accept(-w_cpu[-1] / kT, corr_cpu[-1], lu))
        dx.append(float(np.max(np.abs(c.x - x1[r]))))
    w_cpu, corr_cpu, acc_cpu = np.array(w_cpu), np.array(corr_cpu), np.array(acc_cpu, bool)
    assert np.std(w_gpu) > 0
    # (i) walker by walker: the float32 engine stays on the float64 oracle's trajectory through the whole protocol
    assert np.median(dx) < 1e-3, (np.median(dx), np.max(dx))
    assert np.median(np.abs(w_gpu - w_cpu)) < 0.05 * kT, np.abs(w_gpu - w_cpu)
    # (ii) the alchemical correction (difference of four ~1e4 kJ/mol totals) to 1e-3 kT per walker where they coincide
    same = np.asarray(dx) < 1e-4
    assert same.sum() >= R // 2
    assert np.max(np.abs(corr_gpu - corr_cpu)[same]) < 0.05, np.abs(corr_gpu - corr_cpu)[same]
    # (iii) distributions and acceptance
    assert ks_2samp(w_gpu, w_cpu).pvalue > 0.05
    k_gpu, k_cpu = int(np.sum(acc_gpu)), int(np.sum(acc_cpu))
    p_ref = min(max(k_cpu / R, 0.5 / R), 1 - 0.5 / R)
    assert binomtest(k_gpu, R, p_ref).pvalue > 0.01, (k_gpu, k_cpu)
    decided = np.abs((-w_cpu / kT + corr_cpu) - logu) > 0.1          # walkers whose decision is not on the edge
    assert np.array_equal(acc_gpu.astype(bool)[decided], acc_cpu[decided])
    eng.close()
    md.close()


def test_alchemical_correction_value_matches_oracle():
    """`_computeAlchemicalCorrection` (blues/simulation.py:1100-1119) through the public API on TOL-parm: the value —
    a difference of four totals of order 1e4 kJ/mol — against oracle.alchemical_correction on the same coordinates."""
    import os
    from oracle import ncmc_oracle as orc
    from oracle.c_oracle import COracle
    from blues_b200 import unit, utils
    from blues_b200.structure import Structure
    from blues_b200.simulation import SystemFactory, SimulationFactory, BLUESSimulation
    from blues_b200.moves import RandomLigandRotationMove, MoveEngine
    import tests.test_gpu_api as api
    structure = Structure.load_npz(os.path.join(gc.GOLDEN, 'tol_parm.npz'))
    idx = utils.atomIndexfromTop('LIG', structure.topology)
    systems = SystemFactory(structure, idx, api.system_cfg())
    cfg = api.sim_cfg()
    cfg.update(nstepsNC=20, seed=99)
    simulations = SimulationFactory(systems, MoveEngine(RandomLigandRotationMove(structure, 'LIG')), cfg)
    b = BLUESSimulation(simulations)
    b._md_sim.minimizeEnergy(maxIterations=200)
    b._syncStatesMDtoNCMC()
    b._stepNCMC(20, 10)
    corr = b._computeAlchemicalCorrection()
    st = b.stateTable
    x0 = st['md']['state0']['positions'].value_in_unit(unit.nanometers)
    x1 = st['ncmc']['state1']['positions'].value_in_unit(unit.nanometers)
    topo_md, topo_nc = systems.md.flatten(), systems.alch.flatten()
    o_md, o_nc = COracle(topo_md), COracle(topo_nc)
    e_md0, e_md1 = o_md.energy_forces(x0)[0], o_md.energy_forces(x1)[0]
    e_nc0, e_nc1 = o_nc.energy_forces(x0, 1.0, 1.0)[0], o_nc.energy_forces(x1, 1.0, 1.0)[0]
    kT = orc.KB * 300.0
    want = orc.alchemical_correction(e_nc0, e_md0, e_md1, e_nc1, kT)
    assert np.isfinite(corr)
    assert corr == pytest.approx(want, abs=1e-3)                     # 1e-3 kT
    # each of the four totals to 1e-6 relative
    for got, ref in ((st['md']['state0']['potential_energy']._value, e_md0),
                     (st['ncmc']['state0']['potential_energy']._value, e_nc0),
                     (st['ncmc']['state1']['potential_energy']._value, e_nc1)):
        assert got == pytest.approx(ref, rel=2e-6)


def test_nan_in_one_ncmc_leg_is_rejected_and_the_next_iteration_recovers():
    """a non-finite coordinate raises from the stepping call (OpenMM: 'Particle coordinate is nan'), the state
    stays readable, the work reads NaN so the move can only be rejected, and fresh coordinates clear the latch."""
    from blues_b200 import _native
    eng, o, topo = gc.make_ncmc_pair('vac_divaline', nsteps=10, seed=9)
    good = eng.get_positions(0)
    bad = good.copy()
    bad[3, 0] = np.nan
    eng.set_positions(bad)
    with pytest.raises(_native.EngineError, match='nan'):
        eng.ncmc_run(2)
    assert np.isnan(eng.get_global('protocol_work'))
    eng.get_energy()                                          # state queries do not raise on the flagged walker
    acc, logp, logu = eng.accept_reject()
    assert acc[0] == 0
    eng.reset_ncmc()
    eng.set_positions(good)
    eng.velocities_to_temperature(300.0)
    eng.ncmc_run(4)
    assert np.isfinite(eng.get_global('protocol_work')) and eng.get_global('step') == 4
    eng.close()


def test_energy_query_between_host_move_and_step_keeps_the_external_work():
    """setPositions followed by getState(getEnergy=True) must not swallow perturbed_pe - unperturbed_pe."""
    from oracle import ncmc_oracle as orc
    eng, o, topo = gc.make_ncmc_pair('vac_divaline', nsteps=10, seed=9)
    eng.ncmc_run(3)
    o.step(3)
    x = eng.get_positions(0)
    x[20] += np.array([0.01, -0.02, 0.015])
    eng.set_positions(x)
    eng.get_energy()                                          # the query a user Move subclass might make
    eng.get_forces(0)
    o.x = x.copy()
    eng.ncmc_run(2)
    o.step(2)
    assert eng.get_global('protocol_work') == pytest.approx(o.g['protocol_work'], rel=1e-4, abs=1e-4)
    eng.close()


@pytest.mark.parametrize('what', ['huge_velocity', 'huge_coordinate'])
def test_blown_up_walker_is_flagged_like_nan_and_stays_memory_safe(what):
    """Found by tests/gpu_stress_probe.py on the T4L surrogate (one blow-up in ~2e5 steps): coordinates of ~1e33 nm are
    finite, so the isfinite latch stayed silent, the float -> int conversions of the list builder saturated and
    `z0 + ncz` wrapped around into an out-of-bounds read (compute-sanitizer: build_group_runs).  A walker whose
    coordinates leave +-1e6 nm now raises the NaN flag at once, its mirror is parked inside the box, and the builder
    clamps its scan range; periodic PME system, many steps after the blow-up, neighbour rebuilds included."""
    from blues_b200 import _native
    eng, o, topo = gc.make_ncmc_pair('tol_parm', nsteps=400, seed=4, dt=0.002, minimize=True)
    good = eng.get_positions(0)
    eng.ncmc_run(5)
    if what == 'huge_velocity':
        v = eng.get_velocities(0)
        v[17] = [3.0e33, -2.0e33, 1.0e33]
        v[400:420] *= 1.0e30
        eng.set_velocities(v, 0)
    else:
        x = eng.get_positions(0)
        x[17] = [4.0e33, 1.0e17, -2.0e33]
        x[500] = [1.0e9, 2.0, 3.0]
        eng.set_positions(x)
    with pytest.raises(_native.EngineError, match='nan'):
        eng.ncmc_run(60)                                      # keeps stepping the dead walker: must not fault
    assert np.isnan(eng.get_global('protocol_work'))
    eng.reset_ncmc()
    eng.set_positions(good)
    eng.velocities_to_temperature(300.0)
    eng.ncmc_run(20)
    assert np.isfinite(eng.get_global('protocol_work')) and eng.get_global('step') == 20
    eng.close()
