"""Diagnostic: stability of the TOL-parm NCMC protocol on the engine (python -m tests.gpu_tolparm_probe [protocols])."""
import sys
import numpy as np
from tests.gpu_checks import load_case, lambda_tables
from blues_b200 import _native

n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 20
s, system, topo, x = load_case('tol_parm', True)
ls, le = lambda_tables(100)
eng = _native.Engine(topo, n_replicas=1, seed=3)
eng.set_ncmc_integrator(300.0, 1.0, 0.002, 'H V R O R V H', 100, 1, 0.2, 0.8, ls, le)
eng.set_positions(x)
eng.minimize(100, 10.0)
x0 = eng.get_positions(0)
print('E after minimize', eng.get_energy()[0][0])
eng.velocities_to_temperature(300.0)
v0 = eng.get_velocities(0)
ndof = 3 * topo['n_atoms'] - len(topo['constraints'])
bad = 0
for p in range(n_prot):
    eng.reset_ncmc()
    eng.set_positions(x0)
    eng.set_velocities(v0, 0)
    try:
        trace = []
        for k in range(10):
            eng.ncmc_run(10 if k < 9 else 8)
            ep, ek = eng.get_energy()
            trace.append('%.0f/%.0fK' % (ep[0], 2 * ek[0] / ndof / 0.0083144626))
        print('protocol', p, 'ok work %.2f' % eng.get_global('protocol_work'), ' '.join(trace))
    except Exception as e:                                    # noqa: BLE001
        bad += 1
        print('protocol', p, 'FAILED at step', eng.get_global('step'), str(e)[:60], ' '.join(trace))
print('blown up: %d of %d' % (bad, n_prot))

# bench.py's pattern: one 20-step call, one 78-step call, no energy queries in between; velocities of the restart taken
# from the end of a previous protocol
for mode in ('fresh_v0', 'late_v0'):
    bad = 0
    vv = v0 if mode == 'fresh_v0' else None
    for p in range(n_prot):
        eng.reset_ncmc()
        eng.set_positions(x0)
        eng.set_velocities(v0 if vv is None else vv, 0)
        try:
            eng.ncmc_run(20)
            eng.ncmc_run(78)
            if mode == 'late_v0':
                vv = eng.get_velocities(0)
        except Exception as e:                                # noqa: BLE001
            bad += 1
            print(mode, 'protocol', p, 'FAILED at step', eng.get_global('step'), str(e)[:60])
    print(mode, 'blown up: %d of %d' % (bad, n_prot))
