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
