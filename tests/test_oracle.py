"""CPU tests of the oracle itself: golden vectors, analytic anchors, RNG known answers."""
import os
import numpy as np
import pytest

from blues_b200 import unit as u
from blues_b200.structure import Structure
from blues_b200.alchemy import AbsoluteAlchemicalFactory, AlchemicalRegion
from oracle import ncmc_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _case(name, alch=None, **kw):
    s = Structure.load_npz(os.path.join(GOLDEN, name + '.npz'))
    system = s.createSystem(**kw)
    if alch is not None:
        system = AbsoluteAlchemicalFactory().create_alchemical_system(system, AlchemicalRegion(alchemical_atoms=alch))
    return s, system.flatten(), s.coordinates * 0.1


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def kat(c, k):
        return [int(x) for x in orc.philox4x32(*[np.uint32(v) for v in c], k[0], k[1])]
    assert kat([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    xi = orc.philox_normal3(5, 0, 0, 0, 20000)
    assert abs(xi.mean()) < 0.02 and abs(xi.std() - 1) < 0.02


def test_golden_vectors_reproduce():
    vec = np.load(os.path.join(GOLDEN, 'oracle_vectors.npz'))
    s, t, x = _case('tol_parm', nonbondedMethod='PME', nonbondedCutoff=8.0 * u.angstroms, constraints='HBonds')
    E, F, comp = orc.ForceField(t).energy_forces(x, t['box'])
    assert E == pytest.approx(float(vec['tol_parm/md/energy']), rel=1e-12)
    assert np.allclose(F, vec['tol_parm/md/forces'], rtol=1e-9, atol=1e-6)
    s, t, x = _case('vac_divaline', alch=list(range(16, 35)), nonbondedMethod='NoCutoff', constraints='HBonds')
    E, F, comp = orc.ForceField(t).energy_forces(x, t['box'], 0.5, 0.25)
    assert E == pytest.approx(float(vec['vac_divaline/alch_0.5_0.25/energy']), rel=1e-12)


def test_forces_are_energy_gradients():
    s, t, x = _case('vac_divaline', alch=list(range(16, 35)), nonbondedMethod='NoCutoff')
    ff = orc.ForceField(t)
    E, F, _ = ff.energy_forces(x, t['box'], 0.4, 0.3)
    h = 1e-6
    for a, k in ((0, 0), (17, 2), (30, 1)):
        xp, xm = x.copy(), x.copy()
        xp[a, k] += h
        xm[a, k] -= h
        fd = -(ff.energy(xp, t['box'], 0.4, 0.3) - ff.energy(xm, t['box'], 0.4, 0.3)) / (2 * h)
        assert fd == pytest.approx(F[a, k], rel=1e-5, abs=1e-4)


def test_ewald_sum_is_alpha_independent_and_self_energy():
    s = Structure.load_npz(os.path.join(GOLDEN, 'tol_parm.npz'))
    x = s.coordinates * 0.1
    tot = []
    for tol_, rc in ((5e-4, 0.8), (1e-5, 1.0)):
        t = s.createSystem(nonbondedMethod='PME', nonbondedCutoff=rc * 10 * u.angstroms, ewaldErrorTolerance=tol_).flatten()
        comp = orc.ForceField(t).energy_forces(x, t['box'])[2]
        tot.append(sum(comp[k] for k in ('coulomb_direct', 'pme_reciprocal', 'ewald_self', 'ewald_exclusion', 'exceptions')))
    assert tot[0] == pytest.approx(tot[1], rel=2e-5)
    t = s.createSystem(nonbondedMethod='PME', nonbondedCutoff=10 * u.angstroms, ewaldErrorTolerance=0.005).flatten()
    comp = orc.ForceField(t).energy_forces(x, t['box'])[2]
    assert comp["ewald_self"] == pytest.approx(-56191.99, abs=0.2)       # SURVEY.md Appendix B (quoted to ~1e-6 rel)


def test_alchemical_endpoints():
    """lambda = 1: alchemical system = MD system minus the ligand's reciprocal-space / dispersion terms."""
    s, t_md, x = _case('vac_divaline', nonbondedMethod='NoCutoff')
    _, t_al, _ = _case('vac_divaline', alch=list(range(16, 35)), nonbondedMethod='NoCutoff')
    e_md = orc.ForceField(t_md).energy(x, t_md['box'])
    e_al = orc.ForceField(t_al).energy(x, t_al['box'], 1.0, 1.0)
    assert e_al == pytest.approx(e_md, rel=1e-10)          # no cutoff → identical at lambda = 1
    e0 = orc.ForceField(t_al).energy_forces(x, t_al['box'], 0.0, 0.0)[2]
    assert e0['alch_electrostatics'] == 0.0


def test_constraints_and_energy_conservation():
    s, t, x = _case('vac_divaline', alch=list(range(16, 35)), nonbondedMethod='NoCutoff', constraints='HBonds')
    o = orc.NCMCOracle(t, {'lambda_sterics': '1', 'lambda_electrostatics': '1'}, 'V R R V', 300.0, 1.0, 0.0005, 200)
    o.x = x.copy()
    o.set_velocities_to_temperature(300.0, 0)
    o.step(1)
    e0 = o.energy() + o.kinetic_energy()
    o.step(60)
    e1 = o.energy() + o.kinetic_energy()
    c = t['constraints']
    d = np.linalg.norm(o.x[c[:, 0]] - o.x[c[:, 1]], axis=1)
    assert np.max(np.abs(d - t['constraint_d'])) < 1e-10
    rel_v = np.einsum('ij,ij->i', o.x[c[:, 0]] - o.x[c[:, 1]], o.v[c[:, 0]] - o.v[c[:, 1]])
    assert np.max(np.abs(rel_v)) < 1e-10
    assert abs(e1 - e0) < 0.05 * o.kT * 3            # velocity-Verlet without thermostat conserves energy
    assert o.g['protocol_work'] == pytest.approx(0.0, abs=1e-9)


def test_program_bookkeeping_and_symmetric_protocol():
    s, t, x = _case('vac_divaline', alch=list(range(16, 35)), nonbondedMethod='NoCutoff', constraints='HBonds')
    o = orc.NCMCOracle(t, None, 'H V R O R V H', 300.0, 1.0, 0.001, 10, seed=3)
    assert o.n_lambda_steps == 20
    o.x = x.copy()
    o.set_velocities_to_temperature(300.0, 0)
    o.step(5)
    assert o.g['lambda_'] == pytest.approx(0.5) and o.lam_s == pytest.approx(0.0) and o.lam_e == pytest.approx(0.0)
    w_mid = o.g['protocol_work']
    # a rigid rotation of the decoupled ligand costs no external work at lambda = 0.5
    R = orc.rotation_matrix_from_quaternion(orc.quaternion_from_uniforms(0.3, 0.6, 0.1))
    o.x = orc.rotate_ligand(o.x, np.arange(16, 35), np.ones(19), R)
    o.step(1)
    # external-work term must vanish (intramolecular energy is rotation invariant; environment decoupled)
    assert abs(o.g['perturbed_pe'] - (o.g['perturbed_pe'])) == 0
    o.step(4)
    assert o.g['step'] == 10 and o.g['lambda_'] == pytest.approx(1.0)
    o.step(3)                                   # `if step < nsteps` guard: nothing happens
    assert o.g['step'] == 10
    assert np.isfinite(o.g['protocol_work']) and np.isfinite(w_mid)
    assert o.log_acceptance_probability() == pytest.approx(-o.g['protocol_work'] / o.kT)
    o.reset()
    assert o.g['step'] == 0 and o.g['protocol_work'] == 0.0 and o.g['lambda_'] == 0.0


def test_moves_and_acceptance_rules():
    R = orc.rotation_matrix_from_quaternion(orc.quaternion_from_uniforms(0.2, 0.7, 0.9))
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
    x = np.random.RandomState(0).rand(10, 3)
    y = orc.rotate_ligand(x, np.arange(4), np.array([12.0, 1.0, 1.0, 16.0]), R)
    d0 = np.linalg.norm(x[0] - x[3]); d1 = np.linalg.norm(y[0] - y[3])
    assert d0 == pytest.approx(d1) and np.all(y[4:] == x[4:])
    p = orc.random_sphere_point(0.9, np.zeros(3), 0.5, 0.25, 0.75)
    assert np.linalg.norm(p) == pytest.approx(0.9 * 0.5 ** (1 / 3))
    assert orc.metropolis_accept(float('nan'), 5.0, -100.0) is False        # NaN work is always rejected
    assert orc.metropolis_accept(-1.0, 0.0, -2.0) is True
    assert orc.metropolis_accept(-999999 / 2.49, 0.0, -2.0) is False        # the 999999 sentinel forces rejection
    assert orc.alchemical_correction(1.0, 2.0, 3.0, 4.0, 2.0) == pytest.approx(1.0)


def test_water_translation_contract():
    """blues/tests/test_watertranslation.py:54-112 on the oracle's restatement of the three hooks: the alchemical water
    trades places with a water inside the sphere, lands on a point of the sphere, protocol_work stays 0 in bounds and
    becomes 999999 out of bounds (and only when the swap happened)."""
    s, topo, x = _case('tol_parm', nonbondedMethod='PME', nonbondedCutoff=8.0 * u.angstroms, constraints='HBonds')
    box = np.asarray(topo['box'], float).reshape(-1)[:3]
    waters = [[a.index for a in r.atoms()] for r in s.topology.residues() if r.name in ('WAT', 'HOH')]
    alch = waters[0]
    masses = np.array([12.011, 12.011])
    com = orc.center_of_mass_f32(x, [0, 1], masses)
    assert np.allclose(com, 0.5 * (x[0] + x[1]), atol=1e-6)
    inside = orc.waters_in_sphere(x, box, waters, com, 0.9)
    assert 0 < len(inside) < len(waters)
    for w in inside:
        d = x[w[0]] - com
        d -= box * np.round(d / box)
        assert np.linalg.norm(d) <= 0.9 + 1e-6
    v = np.random.RandomState(1).normal(size=x.shape)
    chosen = inside[len(inside) // 2]
    xs, vs = orc.water_swap(x, v, alch, chosen)
    assert np.array_equal(xs[alch], x[chosen]) and np.array_equal(xs[chosen], x[alch])
    assert np.array_equal(vs[alch], v[chosen]) and np.array_equal(vs[chosen], v[alch])
    xt = orc.water_translate(xs, box, alch, com, 0.9, 0.3, 0.6, 0.8)
    assert orc.periodic_distance_f32(xt[alch[0]], com, box) <= 0.9
    assert np.allclose(xt[alch[1]] - xt[alch[0]], xs[alch[1]] - xs[alch[0]], atol=1e-12)      # rigid translation
    rest = np.setdiff1d(np.arange(len(x)), alch)
    assert np.array_equal(xt[rest], xs[rest])
    assert orc.water_after_move(xt, box, alch, com, 0.9, True, 0.0) == 0.0
    out = xt.copy()
    out[alch] = out[alch] - out[alch[0]] + (com + np.array([1.0, 0.0, 0.0]))
    assert orc.water_after_move(out, box, alch, com, 0.9, True, 0.0) >= 999999
    assert orc.water_after_move(out, box, alch, com, 0.9, False, 0.0) == 0.0
    # a water at or beyond the radius is not translated (blues/moves.py:1037-1040)
    assert np.array_equal(orc.water_translate(out, box, alch, com, 0.9, 0.3, 0.6, 0.8), out)


class _OracleEngine(object):
    """Duck type of blues_b200._native.Engine for host-logic tests on the CPU: state in numpy, energies from the oracle."""

    def __init__(self, topo, x):
        from oracle.c_oracle import COracle
        self.topo = dict(topo)
        self.n_replicas = 1
        self.x = np.asarray(x, float).copy()
        self.box = np.asarray(topo['box'], float).reshape(-1)[:3].copy()
        self._mk = lambda t: COracle(t)

    def get_box(self):
        return self.box.copy()

    def set_box(self, box):
        self.box = np.asarray(box, float).copy()

    def get_positions(self, replica=0):
        return self.x.copy()

    def set_positions(self, x, replica=-1):
        self.x = np.asarray(x, float).copy()

    def energy_at(self, x, box):
        t = dict(self.topo)
        t['box'] = np.asarray(box, float)
        return self._mk(t).energy_forces(x)[0]

    def get_energy(self, potential=True, kinetic=True):
        return np.array([self.energy_at(self.x, self.box)]), np.zeros(1)


def test_monte_carlo_barostat_host_logic_matches_oracle_restatement():
    """blues_b200.barostat (the MD leg's MonteCarloBarostat, blues/simulation.py:602-626) against the oracle's
    independent restatement of OpenMM's volume move: same trial coordinates, box, work and decision for the same
    uniforms; a rejected move restores the state; the step size adapts after 10 attempts."""
    from blues_b200.barostat import MonteCarloBarostatDriver, molecule_ids
    s, topo, x = _case('wat_divaline', nonbondedMethod='CutoffPeriodic', nonbondedCutoff=9.0 * u.angstroms,
                       constraints='HBonds')
    eng = _OracleEngine(topo, x)
    drv = MonteCarloBarostatDriver(topo, 1.01325, 300.0, 25, seed=5)
    mol = molecule_ids(topo)
    molecules = [np.nonzero(mol == m)[0].tolist() for m in range(mol.max() + 1)]
    assert drv.n_molecules == len(molecules) and len(molecules) > 100
    assert sorted(len(m) for m in molecules)[len(molecules) // 2] == 3          # mostly waters
    rs = np.random.RandomState(3)
    n_acc = 0
    for k in range(12):
        u_vol, u_acc = rs.random_sample(2)
        x0, box0 = eng.get_positions(), eng.get_box()
        scale = drv.volume_scale if drv.volume_scale is not None else 0.01 * np.prod(box0)
        ok_ref, x_ref, box_ref, w = orc.mc_barostat_trial(x0, box0, molecules, eng.energy_at, 1.01325, 300.0, scale, u_vol, u_acc)
        ok = drv.attempt(eng, (u_vol, u_acc))
        assert ok == ok_ref, (k, w)
        assert np.allclose(eng.get_box(), box_ref, rtol=0, atol=1e-12)
        assert np.allclose(eng.get_positions(), x_ref, rtol=0, atol=1e-10)
        if not ok:
            assert np.array_equal(eng.get_positions(), x0) and np.array_equal(eng.get_box(), box0)
        n_acc += int(ok)
        # rigid translation of whole molecules: intramolecular geometry untouched
        m = molecules[7]
        assert np.allclose(eng.get_positions()[m] - eng.get_positions()[m][0], x[m] - x[m][0], atol=1e-9)
    assert drv.total_attempted == 12 and drv.total_accepted == n_acc
    # OpenMM resets its window only when it retunes the step (< 25 % or > 75 % accepted after >= 10 attempts)
    assert drv.attempted in (12, 2) and drv.volume_scale > 0
    if drv.attempted == 2:
        assert drv.volume_scale != pytest.approx(0.01 * np.prod(np.asarray(topo['box'], float).reshape(-1)[:3]))


def _rock_salt(cells, a0, method_tol=5e-4, cutoff=None):
    """NaCl lattice, cells^3 conventional cells of edge a0 (nm): flat topology for the oracles."""
    from blues_b200.system import System, NonbondedForce, PME
    basis_na = [(0, 0, 0), (.5, .5, 0), (.5, 0, .5), (0, .5, .5)]
    basis_cl = [(.5, 0, 0), (0, .5, 0), (0, 0, .5), (.5, .5, .5)]
    xyz, q = [], []
    for i in range(cells):
        for j in range(cells):
            for k in range(cells):
                for b, charge in ((basis_na, 1.0), (basis_cl, -1.0)):
                    for f in b:
                        xyz.append(((i + f[0]) * a0, (j + f[1]) * a0, (k + f[2]) * a0))
                        q.append(charge)
    n = len(q)
    system = System(n)
    system.masses[:] = 22.99
    nb = NonbondedForce(n)
    nb.charge[:] = q
    nb.sigma[:] = 0.3
    nb.epsilon[:] = 0.0                                     # Coulomb only
    nb.method = PME
    nb.cutoff = cutoff or 0.45 * cells * a0
    nb.ewald_tol = method_tol
    nb.use_dispersion_correction = False
    system.addForce(nb)
    system.box = np.array([cells * a0] * 3)
    return system.flatten(), np.asarray(xyz, float)


def test_madelung_constant_of_rock_salt():
    """Literature anchor for the Ewald / smooth-PME arithmetic (direct erfc sum + reciprocal space + self term): the
    lattice energy per ion pair of NaCl is -M e^2 / (4 pi eps0 r0) with the Madelung constant M = 1.747565.  Both
    oracle implementations (numpy, C) reproduce it at two Ewald tolerances, within the accuracy each tolerance buys."""
    from oracle.c_oracle import COracle
    a0, cells = 0.564, 4                                    # 512 ions, box 2.256 nm
    r0 = 0.5 * a0
    expect = -1.747565 * orc.ONE_4PI_EPS0 / r0 if hasattr(orc, 'ONE_4PI_EPS0') else -1.747565 * 138.935456 / r0
    for tol, rel in ((5e-4, 2e-3), (1e-5, 1e-4)):
        topo, x = _rock_salt(cells, a0, tol)
        pairs = topo['n_atoms'] // 2
        e_np = orc.ForceField(topo).energy(x, np.asarray(topo['box'], float))
        e_c = COracle(topo).energy_forces(x)[0]
        assert e_np / pairs == pytest.approx(expect, rel=rel), (tol, e_np / pairs, expect)
        assert e_c == pytest.approx(e_np, rel=1e-10)
        # a perfect lattice: no net force on any ion
        f = COracle(topo).energy_forces(x)[1]
        assert np.max(np.abs(f)) < 1e-3 * abs(expect) / r0


def test_c_oracle_one_evaluation_mode_is_the_same_trajectory():
    """The bench's "optimised CPU" row (oracle/c_oracle.py: set_fast): one full evaluation per coordinate set plus the
    alchemical pairs on lambda changes gives the trajectory and work of the reference's 3-evaluation program."""
    from oracle.c_oracle import COracle
    from blues_b200.workloads import lambda_tables
    from tests import gpu_checks as gc
    s, system, topo, x = gc.load_case('wat_divaline', True)
    ls, le = lambda_tables(10)
    out = []
    for fast in (False, True):
        c = COracle(topo, ls, le, 'H V R O R V H', 300.0, 1.0, 0.001, 10, 1, 0.2, 0.8, seed=5)
        if fast:
            c.set_fast(True)
        c.set_state(x)
        c.velocities_to_temperature(300.0)
        c.step(6)
        out.append((c.x.copy(), c.get('protocol_work'), c.get('n_evals')))
    assert np.max(np.abs(out[0][0] - out[1][0])) < 1e-10
    assert out[0][1] == pytest.approx(out[1][1], abs=1e-7)
    assert out[1][2] < 0.5 * out[0][2]
