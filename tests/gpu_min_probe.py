import numpy as np
from tests.gpu_checks import load_case, lambda_tables
from blues_b200 import _native
s, system, topo, x = load_case('t4l_surrogate', True)
ls, le = lambda_tables(5000)
eng = _native.Engine(topo, n_replicas=1, seed=1)
eng.set_ncmc_integrator(300.0, 1.0, 0.004, 'H V R O R V H', 5000, 1, 0.2, 0.8, ls, le)
eng.set_positions(x)
print('E0', eng.get_energy()[0][0], eng.neighbor_stats())
for k in range(14):
    try:
        eng.minimize(10, 10.0)
        xx = eng.get_positions(0)
        print(k, 'E', eng.get_energy()[0][0], 'items', eng.neighbor_stats(), 'max disp', np.abs(xx - x).max(), 'finite', np.isfinite(xx).all())
    except Exception as e:
        print(k, 'ERR', e, eng.neighbor_stats())
        xx = eng.get_positions(0)
        print('   max disp', np.nanmax(np.abs(xx - x)), 'finite', np.isfinite(xx).all())
        break
