"""T4L surrogate stability probe: python -m tests.gpu_stability_probe"""
import sys, time
import numpy as np
from tests.gpu_checks import load_case, lambda_tables
from blues_b200 import _native

def main():
    s, system, topo, x = load_case('t4l_surrogate', True)
    ls, le = lambda_tables(5000)
    ndof = 3 * topo['n_atoms'] - len(topo['constraints']) - 3
    for seed in (20261017, 11, 5):
        eng = _native.Engine(topo, n_replicas=1, seed=seed)
        eng.set_ncmc_integrator(300.0, 1.0, 0.004, 'H V R O R V H', 5000, 1, 0.2, 0.8, ls, le)
        eng.set_positions(x)
        eng.minimize(100, 10.0)
        print('seed', seed, 'E after minimize', eng.get_energy()[0][0])
        eng.velocities_to_temperature(300.0)
        try:
            for blk in range(10):
                eng.ncmc_run(200)
                ep, ek = eng.get_energy()
                v = eng.get_velocities(0)
                vmax = np.linalg.norm(v, axis=1)
                F = eng.get_forces(0)
                fmax = np.linalg.norm(F, axis=1)
                print('  step %4d  Epot %.1f  T %.1f  max|v| %.2f (atom %d %s %s)  max|F| %.0f (atom %d %s)  items %s' % (
                    (blk + 1) * 200, ep[0], 2 * ek[0] / ndof / 0.0083144720, vmax.max(), vmax.argmax(), s.atom_names[vmax.argmax()],
                    s.residue_names[s.atom_residue[vmax.argmax()]], fmax.max(), fmax.argmax(), s.atom_names[fmax.argmax()], eng.neighbor_stats()))
        except Exception as e:
            print('  FAILED', e)
            v = eng.get_velocities(0); vmax = np.linalg.norm(v, axis=1)
            idx = np.argsort(vmax)[-8:]
            print('  fastest atoms', [(int(i), s.atom_names[i], s.residue_names[s.atom_residue[i]], round(float(vmax[i]), 1)) for i in idx])
        eng.close()

if __name__ == '__main__':
    main()
