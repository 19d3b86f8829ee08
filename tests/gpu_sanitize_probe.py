"""compute-sanitizer target: python -m tests.gpu_sanitize_probe [case] — one evaluation and a few NCMC steps."""
import sys
from tests.gpu_checks import load_case, lambda_tables
from blues_b200 import _native
case = sys.argv[1] if len(sys.argv) > 1 else 'wat_divaline'
s, system, topo, x = load_case(case, True)
ls, le = lambda_tables(100)
eng = _native.Engine(topo, n_replicas=1, seed=1)
eng.set_ncmc_integrator(300.0, 1.0, 0.002, 'H V R O R V H', 100, 1, 0.2, 0.8, ls, le)
eng.set_positions(x)
print('E0', eng.get_energy()[0][0], eng.neighbor_stats())
eng.use_graphs(False)
eng.velocities_to_temperature(300.0)
eng.ncmc_run(6)
eng.synchronize()
print('done', eng.get_global('protocol_work'))
