"""ctypes binding of the oracle's C twin (oracle/ncmc_oracle.c) — test infrastructure and CPU baseline only."""
import ctypes as C
import os
import subprocess
import numpy as np

from blues_b200._native import BlIntegratorParams, BlTopology, build_topology, BL_INTEGRATOR_NCMC

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'liboracle.so')
_dp = C.POINTER(C.c_double)
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(['make', '-s', '-C', HERE])
        lib = C.CDLL(LIB)
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.POINTER(BlTopology), C.POINTER(BlIntegratorParams), C.c_uint64, C.c_int]
        lib.orc_energy_forces.restype = C.c_double
        lib.orc_energy_forces.argtypes = [C.c_void_p, _dp, C.c_double, C.c_double, _dp, _dp]
        lib.orc_step.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        lib.orc_velocities_to_temperature.argtypes = [C.c_void_p, _dp, _dp, C.c_double]
        lib.orc_reset.argtypes = [C.c_void_p]
        lib.orc_get.restype = C.c_double
        lib.orc_get.argtypes = [C.c_void_p, C.c_char_p]
        lib.orc_set_lambda.argtypes = [C.c_void_p, C.c_double, C.c_double]
        lib.orc_num_threads.restype = C.c_int
        lib.orc_set_num_threads.argtypes = [C.c_int]
        lib.orc_set_fast.argtypes = [C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def set_threads(n):
    """OpenMP threads used by oracles created from now on (torchrun exports OMP_NUM_THREADS=1)."""
    load().orc_set_num_threads(int(n))
    return load().orc_num_threads()


def _p(a):
    return a.ctypes.data_as(_dp)


class COracle(object):
    """Reference-semantics NCMC on the CPU (3 full evaluations per step), float64, OpenMP."""

    def __init__(self, topo, lambda_sterics=None, lambda_electrostatics=None, splitting='H V R O R V H',
                 temperature=300.0, collision_rate=1.0, timestep=0.002, nsteps_neq=0, nprop=1, prop_lambda_min=2.0,
                 prop_lambda_max=-1.0, seed=0, replica=0):
        self.lib = load()
        self.n = int(topo['n_atoms'])
        t, self._keep = build_topology(topo)
        p = BlIntegratorParams()
        p.kind = BL_INTEGRATOR_NCMC
        p.temperature, p.friction, p.timestep, p.constraint_tol = temperature, collision_rate, timestep, 1e-8
        p.splitting = splitting.encode()
        p.nsteps_neq, p.nprop = int(nsteps_neq), int(nprop)
        p.prop_lambda_min, p.prop_lambda_max = prop_lambda_min, prop_lambda_max
        n_H = splitting.split().count('H')
        n = nsteps_neq * n_H + 1
        self._ls = np.ascontiguousarray(lambda_sterics if lambda_sterics is not None else np.ones(n), np.float64)
        self._le = np.ascontiguousarray(lambda_electrostatics if lambda_electrostatics is not None else np.ones(n), np.float64)
        assert len(self._ls) == n and len(self._le) == n
        p.n_lambda = n
        p.lambda_sterics, p.lambda_electrostatics = _p(self._ls), _p(self._le)
        self.h = self.lib.orc_create(C.byref(t), C.byref(p), C.c_uint64(seed), int(replica))
        self.x = np.zeros((self.n, 3))
        self.v = np.zeros((self.n, 3))

    @property
    def threads(self):
        return self.lib.orc_num_threads()

    def energy_forces(self, x, lam_s=1.0, lam_e=1.0):
        x = np.ascontiguousarray(x, np.float64)
        F = np.zeros((self.n, 3))
        terms = np.zeros(12)
        E = self.lib.orc_energy_forces(self.h, _p(x), lam_s, lam_e, _p(F), _p(terms))
        return E, F, terms

    def set_state(self, x, v=None):
        self.x = np.ascontiguousarray(x, np.float64).copy()
        if v is not None:
            self.v = np.ascontiguousarray(v, np.float64).copy()

    def velocities_to_temperature(self, T):
        self.lib.orc_velocities_to_temperature(self.h, _p(self.x), _p(self.v), float(T))

    def step(self, n=1):
        self.lib.orc_step(self.h, _p(self.x), _p(self.v), int(n))

    def reset(self):
        self.lib.orc_reset(self.h)

    def set_fast(self, on=True):
        """One full evaluation per coordinate set, alchemical pairs re-evaluated on lambda changes (bench context row:
        the lambda-separable evaluation the engine uses; same trajectory and work as the 3-evaluation program)."""
        self.lib.orc_set_fast(self.h, 1 if on else 0)

    def get(self, name):
        return self.lib.orc_get(self.h, name.encode())
