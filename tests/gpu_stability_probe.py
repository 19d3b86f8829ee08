"""T4L surrogate stability probe (not a test): python -m tests.gpu_stability_probe [replicas] [steps] [seed]"""
import sys
import numpy as np
from tests.gpu_checks import load_case, lambda_tables
from blues_b200 import _native


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 11
    s, system, topo, x = load_case('t4l_surrogate', True)
    ls, le = lambda_tables(5000)
    ndof = 3 * topo['n_atoms'] - len(topo['constraints']) - 3
    eng = _native.Engine(topo, n_replicas=R, seed=seed)
    eng.set_ncmc_integrator(300.0, 1.0, 0.004, 'H V R O R V H', 5000, 1, 0.2, 0.8, ls, le)
    eng.set_positions(x)
    eng.minimize(60, 10.0)
    eng.velocities_to_temperature(300.0)
    done = 0
    try:
        while done < steps:
            eng.ncmc_run(250)
            done += 250
            ep, ek = eng.get_energy()
            T = 2 * np.asarray(ek) / ndof / 0.0083144720
            vmax = max(np.linalg.norm(eng.get_velocities(r), axis=1).max() for r in range(R))
            print('step %5d  T %s  Epot[0] %.0f  max|v| %.1f  work %s  rebuilds %d' % (
                done, np.round(T, 1), ep[0], vmax, np.round([eng.get_global('protocol_work', r) for r in range(R)], 1),
                eng.neighbor_stats()[1]), flush=True)
    except Exception as e:
        print('FAILED after', done, 'steps:', e)
        for r in range(R):
            v = eng.get_velocities(r)
            xx = eng.get_positions(r)
            vm = np.linalg.norm(v, axis=1)
            bad = np.where(~np.isfinite(xx).all(axis=1))[0]
            print('  walker', r, 'nonfinite atoms', len(bad), bad[:10], [(s.atom_names[i], s.residue_names[s.atom_residue[i]]) for i in bad[:6]],
                  'fastest', [(int(i), s.atom_names[i], round(float(vm[i]), 1)) for i in np.argsort(np.nan_to_num(vm))[-4:]])
    eng.close()


if __name__ == '__main__':
    main()
