"""Small driver for ncu captures: python -m tests.gpu_ncu_target [replicas] [steps]"""
import sys
from tests.gpu_checks import load_case, lambda_tables
from blues_b200 import _native
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
s, system, topo, x = load_case('t4l_surrogate', True)
ls, le = lambda_tables(5000)
eng = _native.Engine(topo, n_replicas=R, seed=11)
eng.set_ncmc_integrator(300.0, 1.0, 0.004, 'H V R O R V H', 5000, 1, 0.2, 0.8, ls, le)
eng.set_positions(x)
eng.minimize(30, 10.0)
eng.velocities_to_temperature(300.0)
eng.use_graphs(False)
eng.ncmc_run(steps)
eng.synchronize()
print('done', eng.get_global('protocol_work'))
