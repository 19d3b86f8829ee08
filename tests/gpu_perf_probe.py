"""Quick performance probe on the T4L surrogate (not a test): python -m tests.gpu_perf_probe [replicas] [steps]"""
import sys, time
import numpy as np
from tests.gpu_checks import load_case, lambda_tables, compare_forces, compare_neighbors
from blues_b200 import _native

def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    nsteps = 5000
    s, system, topo, x = load_case('t4l_surrogate', True)
    ls, le = lambda_tables(nsteps)
    eng = _native.Engine(topo, n_replicas=R, seed=11)
    eng.set_ncmc_integrator(300.0, 1.0, 0.004, 'H V R O R V H', nsteps, 1, 0.2, 0.8, ls, le)
    eng.set_positions(x)
    t0 = time.time(); eng.minimize(60, 10.0); print('minimize %.2fs' % (time.time() - t0), 'E', eng.get_energy()[0][:2])
    eng.velocities_to_temperature(300.0)
    eng.ncmc_run(50)
    eng.synchronize()
    for rep in range(3):
        t0 = time.time(); eng.ncmc_run(steps); eng.synchronize(); dt = time.time() - t0
        print('R=%d graphs: %d steps in %.3fs -> %.1f steps/s/replica, %.1f aggregate; %.1f us/step' % (R, steps, dt, steps / dt, R * steps / dt, 1e6 * dt / steps))
    print('work', eng.get_global('protocol_work'), 'lambda', eng.get_global('lambda'), 'items,rebuilds', eng.neighbor_stats(), 'T_kin', eng.get_energy()[1][0] * 2 / (3 * topo['n_atoms'] - len(topo['constraints'])) / 0.0083144720)
    eng.set_profiling(True)
    t0 = time.time(); eng.ncmc_run(100); eng.synchronize(); dt = time.time() - t0
    print('profiling mode (direct launches): %.1f us/step' % (1e6 * dt / 100))
    tot = 0
    for k in _native.KERNEL_IDS:
        ms, n = eng.kernel_time(k)
        tot += ms
        print('   %-14s %8.3f ms total over %5d launches = %8.2f us/launch, %7.2f us/step' % (k, ms, n, 1e3 * ms / max(n, 1), 1e3 * ms / 100))
    print('   sum %.2f us/step' % (1e3 * tot / 100))
    eng.set_profiling(False)
    eng.close()

if __name__ == '__main__':
    if len(sys.argv) > 3:
        compare_forces('t4l_surrogate', False, verbose=True)
        compare_forces('t4l_surrogate', True, lam_index=5000, nsteps=5000, verbose=True)
        compare_neighbors('t4l_surrogate', verbose=True)
    main()
