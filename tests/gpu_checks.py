"""Engine-vs-oracle comparisons shared by the ``-m gpu`` tests, ``__graft_entry__.smoke()`` and a diagnostic CLI.

``python -m tests.gpu_checks`` prints every comparison (used for the first bring-up runs on the GPU box).
Everything here reads only ``tests/golden/`` — ``/root/reference`` does not exist on the GPU box.
"""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

from blues_b200 import unit as u  # noqa: E402
from blues_b200.structure import Structure  # noqa: E402
from blues_b200.alchemy import AbsoluteAlchemicalFactory, AlchemicalRegion  # noqa: E402
from blues_b200 import _native  # noqa: E402

from blues_b200.workloads import CASES, DEFAULT_FUNCS, tile_structure  # noqa: E402,F401


def far_residue_atoms(s, center_atoms, distance_nm):
    """Atoms of every residue that has no atom within ``distance_nm`` of the centre atoms (periodic): the complement of
    the reference's ``:LIG<:d`` residue selection (blues/simulation.py:395-480), usable on fixtures without a protein."""
    x = np.asarray(s.coordinates, float) * 0.1
    box = np.asarray(s.box[:3], float) * 0.1
    res = np.asarray(s.to_arrays()['atom_residue'])
    near = np.zeros(len(x), bool)
    for c in center_atoms:
        d = x - x[c]
        d -= box * np.round(d / box)
        near |= (np.sum(d * d, axis=1) <= distance_nm ** 2)
    keep = np.zeros(res.max() + 1, bool)
    keep[res[near]] = True
    return np.nonzero(~keep[res])[0]


def load_case(name, alchemical=False, freeze_beyond_nm=None, restrain=None, **overrides):
    """freeze_beyond_nm: zero the masses (utils.zero_masses, as freeze_radius does) of all residues farther than that from
    the alchemical atoms; restrain: Amber mask for SystemFactory.restrain_positions."""
    s = Structure.load_npz(os.path.join(GOLDEN, name + '.npz'))
    kw = dict(CASES[name]['kw'])
    kw.update(overrides)
    system = s.createSystem(**kw)
    if restrain:
        from blues_b200.simulation import SystemFactory
        system = SystemFactory.restrain_positions(s, system, selection=restrain, weight=5.0)
    if freeze_beyond_nm is not None:
        from blues_b200 import utils
        system = utils.zero_masses(system, far_residue_atoms(s, CASES[name]['alch'], freeze_beyond_nm))
    if alchemical:
        system = AbsoluteAlchemicalFactory().create_alchemical_system(
            system, AlchemicalRegion(alchemical_atoms=CASES[name]['alch']))
    topo = system.flatten()
    return s, system, topo, s.coordinates * 0.1


def lambda_tables(nsteps, n_H=2, funcs=None):
    from oracle.ncmc_oracle import eval_lambda_function
    funcs = funcs or DEFAULT_FUNCS
    n = nsteps * n_H
    lam = [k / n for k in range(n + 1)]
    ls = [eval_lambda_function(funcs.get('lambda_sterics', '1'), x) for x in lam]
    le = [eval_lambda_function(funcs.get('lambda_electrostatics', '1'), x) for x in lam]
    return np.asarray(ls), np.asarray(le)


def rel_force_error(F, Fref):
    """max over atoms of |dF| / max(|Fref_atom|, rms|Fref|)  and the rms-relative error"""
    d = np.linalg.norm(F - Fref, axis=1)
    fn = np.linalg.norm(Fref, axis=1)
    rms = np.sqrt(np.mean(fn ** 2)) + 1e-30
    return float(np.max(d / np.maximum(fn, rms))), float(np.sqrt(np.mean(d ** 2)) / rms)


ORACLE_TERMS = {
    'bond': ['bond'], 'angle': ['angle'], 'torsion': ['torsion'], 'restraint': ['restraint'],
    'pair_direct': ['lj', 'coulomb_direct'], 'exceptions': ['exceptions', 'ewald_exclusion'],
    'pme_reciprocal': ['pme_reciprocal'], 'ewald_self': ['ewald_self', 'ewald_plasma'], 'dispersion': ['dispersion'],
    'alch_sterics': ['alch_sterics'], 'alch_electrostatics': ['alch_electrostatics'], 'alch_exceptions': ['alch_exceptions']}


def compare_forces(name, alchemical=False, lam_index=None, nsteps=10, verbose=False, **overrides):
    """Energies by term and forces of the engine vs the oracle at the fixture coordinates."""
    from oracle.ncmc_oracle import ForceField
    s, system, topo, x = load_case(name, alchemical, **overrides)
    eng = _native.Engine(topo, n_replicas=1, seed=1)
    ls_tab, le_tab = lambda_tables(nsteps)
    lam_s = lam_e = 1.0
    if alchemical:
        eng.set_ncmc_integrator(300.0, 1.0, 0.002, 'H V R O R V H', nsteps, 1, 0.2, 0.8, ls_tab, le_tab)
        k = 0 if lam_index is None else lam_index
        eng.set_global('lambda_step', k)
        lam_s, lam_e = ls_tab[k], le_tab[k]
    else:
        eng.set_langevin_integrator(300.0, 1.0, 0.002)
    eng.set_positions(x)
    t0 = time.time()
    terms = eng.get_energy_terms()
    F = eng.get_forces()
    ep, ek = eng.get_energy()
    t_gpu = time.time() - t0
    t0 = time.time()
    Eo, Fo, comp = ForceField(topo).energy_forces(x, topo['box'], lam_s, lam_e)
    t_cpu = time.time() - t0
    out = {'energy': float(ep[0]), 'energy_oracle': float(Eo), 'terms': {}, 't_gpu': t_gpu, 't_cpu': t_cpu}
    for k, names in ORACLE_TERMS.items():
        ref = sum(comp.get(nm, 0.0) for nm in names)
        out['terms'][k] = (terms[k], ref)
    out['force_max_rel'], out['force_rms_rel'] = rel_force_error(F, Fo)
    out['energy_rel'] = abs(ep[0] - Eo) / max(abs(Eo), 1.0)
    if verbose:
        print('--- %s alchemical=%s lambda_s=%.3f lambda_e=%.3f' % (name, alchemical, lam_s, lam_e))
        print('   E engine %.6f  oracle %.6f  rel %.2e   (gpu %.2fs cpu %.2fs)' % (ep[0], Eo, out['energy_rel'], t_gpu, t_cpu))
        for k, (a, b) in out['terms'].items():
            if a != 0 or b != 0:
                print('   %-20s engine %16.6f oracle %16.6f  diff %.3e' % (k, a, b, a - b))
        print('   forces: max rel %.3e  rms rel %.3e' % (out['force_max_rel'], out['force_rms_rel']))
    eng.close()
    return out


def compare_neighbors(name, verbose=False, **overrides):
    from oracle.ncmc_oracle import ForceField
    s, system, topo, x = load_case(name, False, **overrides)
    eng = _native.Engine(topo, n_replicas=1, seed=1)
    eng.set_langevin_integrator(300.0, 1.0, 0.002)
    eng.set_positions(x)
    codes = eng.neighbor_pairs()
    ff = ForceField(topo)
    ref = ff.neighbor_pairs(x, topo['box'])
    # pairs whose float32 distance straddles the cutoff may legitimately differ: list them
    only_e = np.setdiff1d(codes, ref)
    only_o = np.setdiff1d(ref, codes)
    n = topo['n_atoms']

    def dist(code):
        i, j = code // n, code % n
        d = x[i] - x[j]
        if topo['nb_method'] != 0:
            d -= topo['box'] * np.round(d / topo['box'])
        return np.linalg.norm(d, axis=1)

    edge = 0.0
    if len(only_e) + len(only_o):
        edge = float(np.max(np.abs(dist(np.concatenate([only_e, only_o])) - topo['cutoff'])))
    tiles, rebuilds = eng.neighbor_stats()
    eng.close()
    if verbose:
        print('--- neighbours %s: engine %d oracle %d, only-engine %d only-oracle %d, max |r-rc| of mismatches %.2e, items %d'
              % (name, len(codes), len(ref), len(only_e), len(only_o), edge, tiles))
    return dict(n_engine=len(codes), n_oracle=len(ref), only_engine=len(only_e), only_oracle=len(only_o), edge=edge,
                duplicates=len(codes) - len(np.unique(codes)))


def compare_neighbors_dynamic(name, steps=60, n_replicas=2, dt=0.004, minimize=0, verbose=False, **overrides):
    """After `steps` of hot dynamics (several prunes and cell-search rebuilds) the list the last evaluation used must
    still hold every non-excluded pair inside the cutoff at the current coordinates, for every walker."""
    from oracle.ncmc_oracle import ForceField
    s, system, topo, x = load_case(name, True, **overrides)
    ls, le = lambda_tables(5000)
    eng = _native.Engine(topo, n_replicas=n_replicas, seed=3)
    eng.set_ncmc_integrator(300.0, 1.0, dt, 'H V R O R V H', 5000, 1, 0.2, 0.8, ls, le)
    eng.set_positions(x)
    if minimize:
        eng.minimize(minimize, 10.0)
    eng.velocities_to_temperature(300.0)
    eng.ncmc_run(steps)
    ff = ForceField(topo)
    out = []
    for r in range(n_replicas):
        xr = eng.get_positions(r)
        codes = eng.neighbor_pairs(r)
        ref = ff.neighbor_pairs(xr, topo['box'])
        only_e, only_o = np.setdiff1d(codes, ref), np.setdiff1d(ref, codes)
        n = topo['n_atoms']
        edge = 0.0
        if len(only_e) + len(only_o):
            c = np.concatenate([only_e, only_o])
            d = xr[c // n] - xr[c % n]
            if topo['nb_method'] != 0:
                d -= topo['box'] * np.round(d / topo['box'])
            edge = float(np.max(np.abs(np.linalg.norm(d, axis=1) - topo['cutoff'])))
        out.append(dict(n_engine=len(codes), n_oracle=len(ref), only_engine=len(only_e), only_oracle=len(only_o), edge=edge,
                        duplicates=len(codes) - len(np.unique(codes)), rebuilds=eng.neighbor_stats(r)[1]))
        if verbose:
            print('--- dynamic neighbours %s walker %d:' % (name, r), out[-1])
    eng.close()
    return out


def compare_tiled(name='tol_parm', reps=(4, 4, 5), verbose=False, **overrides):
    """Forces, energy and neighbour sets on a tiled copy of a fixture (> 65 535 atoms: 32-bit list indices, larger
    PME grid, many cells) against the oracle's C twin (forces) and the numpy oracle (pairs)."""
    from oracle.ncmc_oracle import ForceField
    from oracle.c_oracle import COracle
    s0 = Structure.load_npz(os.path.join(GOLDEN, name + '.npz'))
    s = tile_structure(s0, reps)
    kw = dict(CASES[name]['kw'])
    kw.update(overrides)
    topo = s.createSystem(**kw).flatten()
    x = s.coordinates * 0.1
    eng = _native.Engine(topo, n_replicas=1, seed=1)
    eng.set_langevin_integrator(300.0, 1.0, 0.002)
    eng.set_positions(x)
    ep, ek = eng.get_energy()
    F = eng.get_forces()
    codes = eng.neighbor_pairs()
    eng.close()
    Eo, Fo = COracle(topo).energy_forces(x)[:2]
    ref = ForceField(topo).neighbor_pairs(x, topo['box'])
    only_e, only_o = np.setdiff1d(codes, ref), np.setdiff1d(ref, codes)
    edge = 0.0
    if len(only_e) + len(only_o):
        c = np.concatenate([only_e, only_o])
        n = topo['n_atoms']
        dd = x[c // n] - x[c % n]
        dd -= topo['box'] * np.round(dd / topo['box'])
        edge = float(np.max(np.abs(np.linalg.norm(dd, axis=1) - topo['cutoff'])))
    out = {'n_atoms': topo['n_atoms'], 'energy_rel': abs(ep[0] - Eo) / max(abs(Eo), 1.0), 'pairs': len(ref),
           'only_engine': len(only_e), 'only_oracle': len(only_o), 'duplicates': len(codes) - len(np.unique(codes)),
           'edge': edge, 'pme_grid': [int(v) for v in topo['pme_grid']]}
    out['force_max_rel'], out['force_rms_rel'] = rel_force_error(F, Fo)
    if verbose:
        print('--- tiled', name, reps, out)
    return out


def make_ncmc_pair(name, nsteps=10, dt=0.002, splitting='H V R O R V H', nprop=1, prop_lambda=0.3, seed=7,
                   temperature=300.0, n_replicas=1, minimize=False, **overrides):
    """Engine and oracle initialised identically for step-for-step comparisons."""
    from oracle.ncmc_oracle import NCMCOracle, get_prop_lambda
    s, system, topo, x = load_case(name, True, **overrides)
    n_H = splitting.split().count('H')
    ls_tab, le_tab = lambda_tables(nsteps, n_H)
    pmin, pmax = get_prop_lambda(prop_lambda)
    eng = _native.Engine(topo, n_replicas=n_replicas, seed=seed)
    eng.set_ncmc_integrator(temperature, 1.0, dt, splitting, nsteps, nprop, pmin, pmax, ls_tab, le_tab)
    eng.set_positions(x)
    if minimize:
        eng.minimize(200, 10.0)
        x = eng.get_positions(0)
        eng.set_positions(x)
    eng.velocities_to_temperature(temperature)
    orc = NCMCOracle(topo, DEFAULT_FUNCS, splitting, temperature, 1.0, dt, nsteps, nprop, prop_lambda, seed, 0)
    orc.x = x.copy()
    orc.set_velocities_to_temperature(temperature, 0)
    return eng, orc, topo


def make_ncmc_pair_c(name, nsteps=10, dt=0.002, splitting='H V R O R V H', seed=7, temperature=300.0, n_replicas=1,
                     minimize=0, **overrides):
    """Engine and the oracle's C twin (fast enough for long protocols on solvated systems) initialised identically."""
    from oracle.c_oracle import COracle
    s, system, topo, x = load_case(name, True, **overrides)
    n_H = splitting.split().count('H')
    ls_tab, le_tab = lambda_tables(nsteps, n_H)
    eng = _native.Engine(topo, n_replicas=n_replicas, seed=seed)
    eng.set_ncmc_integrator(temperature, 1.0, dt, splitting, nsteps, 1, 0.2, 0.8, ls_tab, le_tab)
    eng.set_positions(x)
    if minimize:
        eng.minimize(minimize, 10.0)
        x = eng.get_positions(0)
        eng.set_positions(x)
    eng.velocities_to_temperature(temperature)
    orcs = []
    for r in range(n_replicas):
        c = COracle(topo, ls_tab, le_tab, splitting, temperature, 1.0, dt, nsteps, 1, 0.2, 0.8, seed=seed, replica=r)
        c.set_state(x)
        c.velocities_to_temperature(temperature)
        orcs.append(c)
    return eng, orcs, topo, x


def compare_trajectory_c(name, nsteps=50, stride=10, verbose=False, **kw):
    """Noisy trajectory of the engine against the C oracle, compared every ``stride`` steps."""
    eng, orcs, topo, x0 = make_ncmc_pair_c(name, nsteps=nsteps, **kw)
    orc = orcs[0]
    dv0 = float(np.max(np.abs(eng.get_velocities(0) - orc.v)))
    rows = []
    for k in range(0, nsteps, stride):
        eng.ncmc_run(stride)
        orc.step(stride)
        xe, ve = eng.get_positions(0), eng.get_velocities(0)
        rows.append(dict(step=k + stride, dx=float(np.max(np.abs(xe - orc.x))), dv=float(np.max(np.abs(ve - orc.v))),
                         work_engine=eng.get_global('protocol_work'), work_oracle=orc.get('protocol_work')))
        if verbose:
            print('   step %3d  max|dx| %.2e  max|dv| %.2e  work engine %.6f oracle %.6f' %
                  (rows[-1]['step'], rows[-1]['dx'], rows[-1]['dv'], rows[-1]['work_engine'], rows[-1]['work_oracle']))
    mass = np.asarray(topo['mass'], float)
    out = dict(dv0=dv0, rows=rows, x0=x0, x_engine=eng.get_positions(0), frozen=np.nonzero(mass == 0)[0], topo=topo)
    eng.close()
    return out


def compare_trajectory(name, nsteps=6, verbose=False, **kw):
    eng, orc, topo = make_ncmc_pair(name, nsteps=nsteps, **kw)
    v0 = eng.get_velocities(0)
    dv0 = float(np.max(np.abs(v0 - orc.v)))
    rows = []
    for k in range(nsteps):
        eng.ncmc_run(1)
        orc.step(1)
        xe, ve = eng.get_positions(0), eng.get_velocities(0)
        rows.append(dict(step=k + 1, dx=float(np.max(np.abs(xe - orc.x))), dv=float(np.max(np.abs(ve - orc.v))),
                         work_engine=eng.get_global('protocol_work'), work_oracle=orc.g['protocol_work'],
                         lam=eng.get_global('lambda')))
        if verbose:
            r = rows[-1]
            print('   step %2d lambda %.3f  max|dx| %.2e  max|dv| %.2e  work engine %.6f oracle %.6f' %
                  (r['step'], r['lam'], r['dx'], r['dv'], r['work_engine'], r['work_oracle']))
    eng.close()
    return dict(dv0=dv0, rows=rows)


def smoke():
    """One small NCMC invocation (toluene in water) on cuda:0, checked against the oracle."""
    out = compare_forces('tol_parm', alchemical=True, lam_index=3, nsteps=10)
    assert out['energy_rel'] < 1e-4, out
    assert out['force_rms_rel'] < 1e-3, out
    eng, orc, topo = make_ncmc_pair('tol_parm', nsteps=4, minimize=True)
    eng.ncmc_run(4)
    w = eng.get_global('protocol_work')
    orc.step(4)
    assert np.isfinite(w)
    assert abs(w - orc.g['protocol_work']) < 1e-3 * max(1.0, abs(orc.g['protocol_work'])), (w, orc.g['protocol_work'])
    print('smoke ok: protocol_work engine %.6f oracle %.6f, launches %d' % (w, orc.g['protocol_work'], eng.launch_count()))
    eng.close()


def main():
    names = sys.argv[1:] or ['vac_divaline', 'tol_parm', 'wat_divaline']
    for nm in names:
        compare_forces(nm, False, verbose=True)
        for k in (0, 3, 10):
            compare_forces(nm, True, lam_index=k, nsteps=10, verbose=True)
        compare_neighbors(nm, verbose=True)
    print('--- trajectory tol_parm (minimised start)')
    compare_trajectory('tol_parm', nsteps=6, verbose=True, minimize=True)
    print('--- trajectory vac_divaline')
    compare_trajectory('vac_divaline', nsteps=6, verbose=True)


if __name__ == '__main__':
    main()
