"""Two (or more) independent engines on one GPU, each driven by its own host thread (ctypes releases the GIL, the
engines' streams overlap on the device): python -m tests.gpu_multi_engine_probe [engines] [replicas_each] [steps]"""
import sys, time, threading
from tests.gpu_checks import load_case, lambda_tables
from blues_b200 import _native

E = int(sys.argv[1]) if len(sys.argv) > 1 else 2
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 160
s, system, topo, x = load_case('t4l_surrogate', True)
ls, le = lambda_tables(5000)
engs = []
for e in range(E):
    eng = _native.Engine(topo, n_replicas=R, seed=11 + e)
    eng.set_ncmc_integrator(300.0, 1.0, 0.004, 'H V R O R V H', 5000, 1, 0.2, 0.8, ls, le)
    eng.set_positions(x)
    eng.minimize(60, 10.0)
    eng.velocities_to_temperature(300.0)
    eng.ncmc_run(40)
    engs.append(eng)
for eng in engs:
    eng.synchronize()
for rep in range(3):
    t0 = time.time()
    th = [threading.Thread(target=lambda e=eng: (e.ncmc_run(steps), e.synchronize())) for eng in engs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.time() - t0
    print('%d engines x %d walkers: %d steps in %.3fs -> %.1f walker-steps/s aggregate' % (E, R, steps, dt, E * R * steps / dt))
