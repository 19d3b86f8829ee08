"""Latency of the API calls that bracket an NCMC leg (diagnostic): python -m tests.gpu_api_latency_probe"""
import time
import numpy as np
import torch
from blues_b200 import mm, unit
from blues_b200.integrators import AlchemicalExternalLangevinIntegrator
from blues_b200.workloads import load_workload, DEFAULT_FUNCS

wl = load_workload('t4l')
integ = AlchemicalExternalLangevinIntegrator(DEFAULT_FUNCS, splitting='H V R O R V H', temperature=300 * unit.kelvin,
                                             timestep=wl['dt'] * unit.picoseconds, nsteps_neq=wl['nsteps_nc'], nprop=1, prop_lambda=0.3)
ctx = mm.Context(wl['system'], integ, mm.Platform.getPlatformByName('CUDA'), {'DeviceIndex': 0})
ctx.setPositions(wl['x'] * unit.nanometers)
eng = ctx._engine
eng.minimize(100, 10.0)
ctx.setVelocitiesToTemperature(300 * unit.kelvin)
integ.step(50)
x = eng.get_positions(0)
v = eng.get_velocities(0)
xq, vq = x * unit.nanometers, v * (unit.nanometers / unit.picoseconds)


def timeit(name, fn, n=30):
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print('%-46s %8.1f us' % (name, 1e6 * (time.perf_counter() - t0) / n))


timeit('ctx.setPositions(Quantity)', lambda: ctx.setPositions(xq))
timeit('eng.set_positions(ndarray)', lambda: eng.set_positions(x))
timeit('ctx.setVelocities(Quantity)', lambda: ctx.setVelocities(vq))
timeit('ctx.getState(getPositions).getPositions(asNumpy)', lambda: ctx.getState(getPositions=True).getPositions(asNumpy=True))
timeit('eng.get_positions(0)', lambda: eng.get_positions(0))
timeit('integ.get_protocol_work(dimensionless)', lambda: integ.get_protocol_work(dimensionless=True))
timeit('eng.accept_reject()', lambda: eng.accept_reject())
timeit('integ.step(1) (forces valid)', lambda: integ.step(1))
timeit('integ.step(4)', lambda: integ.step(4))
timeit('integ.step(20)', lambda: integ.step(20))
timeit('setPositions + step(1)', lambda: (ctx.setPositions(xq), integ.step(1)))
timeit('eng.get_energy()', lambda: eng.get_energy())
