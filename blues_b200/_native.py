"""ctypes binding of ``libblues_b200.so`` (the C ABI in ``include/blues_b200.h``).

There is no CPU fallback: importing this module succeeds without a GPU (so host-side code and symbol
checks work), but creating an :class:`Engine` raises unless a CUDA device is present and the
in-tree shared library has been built (``python -c 'import __graft_entry__ as g; g.build()'``).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libblues_b200.so')

BL_INTEGRATOR_NCMC, BL_INTEGRATOR_LANGEVIN = 1, 2
BL_MOVE_NONE, BL_MOVE_ROTATE = 0, 1
BL_MOVE_WATER_SWAP, BL_MOVE_WATER_TRANSLATE, BL_MOVE_WATER_CHECK = 2, 3, 4
ENERGY_TERMS = ['bond', 'angle', 'torsion', 'restraint', 'pair_direct', 'exceptions', 'pme_reciprocal', 'ewald_self',
                'dispersion', 'alch_sterics', 'alch_electrostatics', 'alch_exceptions']
KERNEL_IDS = {'pair': 0, 'integrate': 1, 'pme_spread': 2, 'pme_gather': 3, 'pme_convolve': 4, 'bonded': 5, 'alch': 6,
              'neighbor': 7, 'fft': 8}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class BlTopology(C.Structure):
    _fields_ = [
        ('n_atoms', C.c_int32), ('mass', _dp), ('charge', _dp), ('sigma', _dp), ('epsilon', _dp),
        ('n_bonds', C.c_int32), ('bonds', _ip), ('bond_k', _dp), ('bond_r0', _dp),
        ('n_angles', C.c_int32), ('angles', _ip), ('angle_k', _dp), ('angle_t0', _dp),
        ('n_torsions', C.c_int32), ('torsions', _ip), ('torsion_k', _dp), ('torsion_n', _ip), ('torsion_phase', _dp),
        ('n_excl', C.c_int32), ('excl_pairs', _ip), ('excl_qq', _dp), ('excl_sigma', _dp), ('excl_eps', _dp),
        ('n_constraints', C.c_int32), ('constraints', _ip), ('constraint_d', _dp),
        ('box', C.c_double * 3), ('nb_method', C.c_int32), ('cutoff', C.c_double), ('ewald_alpha', C.c_double),
        ('pme_grid', C.c_int32 * 3), ('dispersion_coeff', C.c_double), ('remove_cm', C.c_int32),
        ('n_restraints', C.c_int32), ('restraint_atoms', _ip), ('restraint_k', _dp), ('restraint_x0', _dp),
        ('n_alch', C.c_int32), ('alch_atoms', _ip), ('alch_charge', _dp), ('alch_sigma', _dp), ('alch_eps', _dp),
        ('n_alch_exc', C.c_int32), ('alch_exc_pairs', _ip), ('alch_exc_qq', _dp), ('alch_exc_sigma', _dp),
        ('alch_exc_eps', _dp),
        ('softcore_alpha', C.c_double), ('softcore_a', C.c_double), ('softcore_b', C.c_double),
        ('softcore_c', C.c_double), ('annihilate_sterics', C.c_int32), ('annihilate_electrostatics', C.c_int32),
        ('n_custom_terms', C.c_int32), ('custom_term', _ip), ('custom_cutoff', _dp),
        ('custom_n_params', C.c_int32), ('custom_params', _dp),
        ('n_custom_groups', C.c_int32), ('custom_group_start', _ip), ('custom_group_atoms', _ip),
        ('custom_group_weights', _dp),
        ('n_custom_progs', C.c_int32), ('custom_prog_start', _ip), ('custom_code_op', _ip), ('custom_code_arg', _dp),
    ]


class BlIntegratorParams(C.Structure):
    _fields_ = [
        ('kind', C.c_int32), ('temperature', C.c_double), ('friction', C.c_double), ('timestep', C.c_double),
        ('constraint_tol', C.c_double), ('splitting', C.c_char_p), ('nsteps_neq', C.c_int32), ('nprop', C.c_int32),
        ('prop_lambda_min', C.c_double), ('prop_lambda_max', C.c_double), ('n_lambda', C.c_int32),
        ('lambda_sterics', _dp), ('lambda_electrostatics', _dp),
    ]


class BlMove(C.Structure):
    _fields_ = [('kind', C.c_int32), ('step', C.c_int32), ('n_atoms', C.c_int32), ('atoms', _ip), ('masses', _dp),
                ('n_waters', C.c_int32), ('water_atoms', _ip), ('n_center', C.c_int32), ('center_atoms', _ip),
                ('center_masses', _dp), ('radius', C.c_double)]


# every symbol include/blues_b200.h declares: name → (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    'bl_create': (C.c_int, [C.POINTER(BlTopology), C.c_int, C.c_int, C.c_uint64, C.POINTER(_H)]),
    'bl_destroy': (C.c_int, [_H]),
    'bl_last_error': (C.c_char_p, [_H]),
    'bl_num_replicas': (C.c_int, [_H]),
    'bl_num_atoms': (C.c_int, [_H]),
    'bl_set_integrator': (C.c_int, [_H, C.POINTER(BlIntegratorParams)]),
    'bl_set_seed': (C.c_int, [_H, C.c_uint64]),
    'bl_set_positions': (C.c_int, [_H, C.c_int, _dp]),
    'bl_set_velocities': (C.c_int, [_H, C.c_int, _dp]),
    'bl_set_box': (C.c_int, [_H, _dp]),
    'bl_get_positions': (C.c_int, [_H, C.c_int, _dp]),
    'bl_get_velocities': (C.c_int, [_H, C.c_int, _dp]),
    'bl_get_forces': (C.c_int, [_H, C.c_int, _dp]),
    'bl_get_box': (C.c_int, [_H, _dp]),
    'bl_get_energy': (C.c_int, [_H, _dp, _dp]),
    'bl_get_energy_terms': (C.c_int, [_H, C.c_int, _dp]),
    'bl_copy_state': (C.c_int, [_H, _H, C.c_int]),
    'bl_copy_state_masked': (C.c_int, [_H, _H, C.c_int, _ip]),
    'bl_velocities_to_temperature': (C.c_int, [_H, C.c_double]),
    'bl_get_global': (C.c_int, [_H, C.c_int, C.c_char_p, _dp]),
    'bl_set_global': (C.c_int, [_H, C.c_int, C.c_char_p, C.c_double]),
    'bl_reset_ncmc': (C.c_int, [_H]),
    'bl_ncmc_run': (C.c_int, [_H, C.c_int, C.POINTER(BlMove)]),
    'bl_md_run': (C.c_int, [_H, C.c_int]),
    'bl_apply_move': (C.c_int, [_H, C.POINTER(BlMove)]),
    'bl_accept_reject': (C.c_int, [_H, _dp, _ip, _dp, _dp]),
    'bl_minimize': (C.c_int, [_H, C.c_int, C.c_double]),
    'bl_neighbor_pairs': (C.c_int, [_H, C.c_int, C.POINTER(C.c_int64), C.c_size_t, C.POINTER(C.c_size_t)]),
    'bl_neighbor_stats': (C.c_int, [_H, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'bl_launch_count': (C.c_uint64, [_H]),
    'bl_set_profiling': (C.c_int, [_H, C.c_int]),
    'bl_get_kernel_time': (C.c_int, [_H, C.c_int, _dp, C.POINTER(C.c_int64)]),
    'bl_use_graphs': (C.c_int, [_H, C.c_int]),
    'bl_stream': (C.c_void_p, [_H]),
    'bl_synchronize': (C.c_int, [_H]),
    'bl_version': (C.c_char_p, []),
    'bl_measure_fp32_peak': (C.c_int, [C.c_int, _dp]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def load_library():
    """Load the in-tree shared library and type every entry point; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(blues_b200 has no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _arr(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def build_topology(topo):
    """Flat dictionary (``System.flatten()``) → (BlTopology, keep-alive dict of the numpy buffers it points to)."""
    keep = {}

    def dp(key, dtype=np.float64):
        keep[key] = _arr(topo[key], dtype)
        return _p(keep[key], C.c_double if dtype == np.float64 else C.c_int32)

    t = BlTopology()
    t.n_atoms = int(topo['n_atoms'])
    t.mass, t.charge, t.sigma, t.epsilon = dp('mass'), dp('charge'), dp('sigma'), dp('epsilon')
    t.n_bonds = len(topo['bonds'])
    t.bonds, t.bond_k, t.bond_r0 = dp('bonds', np.int32), dp('bond_k'), dp('bond_r0')
    t.n_angles = len(topo['angles'])
    t.angles, t.angle_k, t.angle_t0 = dp('angles', np.int32), dp('angle_k'), dp('angle_t0')
    t.n_torsions = len(topo['torsions'])
    t.torsions, t.torsion_k, t.torsion_n, t.torsion_phase = (dp('torsions', np.int32), dp('torsion_k'),
                                                             dp('torsion_n', np.int32), dp('torsion_phase'))
    t.n_excl = len(topo['excl_pairs'])
    t.excl_pairs, t.excl_qq, t.excl_sigma, t.excl_eps = (dp('excl_pairs', np.int32), dp('excl_qq'),
                                                         dp('excl_sigma'), dp('excl_eps'))
    t.n_constraints = len(topo['constraints'])
    t.constraints, t.constraint_d = dp('constraints', np.int32), dp('constraint_d')
    for k in range(3):
        t.box[k] = float(topo['box'][k])
        t.pme_grid[k] = int(topo['pme_grid'][k])
    t.nb_method = int(topo['nb_method'])
    t.cutoff = float(topo['cutoff'])
    t.ewald_alpha = float(topo['ewald_alpha'])
    t.dispersion_coeff = float(topo['dispersion_coeff'])
    t.remove_cm = int(topo['remove_cm'])
    t.n_restraints = len(topo['restraint_atoms'])
    t.restraint_atoms, t.restraint_k, t.restraint_x0 = (dp('restraint_atoms', np.int32), dp('restraint_k'),
                                                        dp('restraint_x0'))
    t.n_alch = len(topo['alch_atoms'])
    t.alch_atoms, t.alch_charge, t.alch_sigma, t.alch_eps = (dp('alch_atoms', np.int32), dp('alch_charge'),
                                                             dp('alch_sigma'), dp('alch_eps'))
    t.n_alch_exc = len(topo['alch_exc_pairs'])
    t.alch_exc_pairs, t.alch_exc_qq, t.alch_exc_sigma, t.alch_exc_eps = (
        dp('alch_exc_pairs', np.int32), dp('alch_exc_qq'), dp('alch_exc_sigma'), dp('alch_exc_eps'))
    t.softcore_alpha, t.softcore_a = float(topo['softcore_alpha']), float(topo['softcore_a'])
    t.softcore_b, t.softcore_c = float(topo['softcore_b']), float(topo['softcore_c'])
    t.annihilate_sterics = int(topo['annihilate_sterics'])
    t.annihilate_electrostatics = int(topo['annihilate_electrostatics'])
    # generic Custom*Force terms (absent from flat dictionaries written before they existed)
    n_custom = len(topo.get('custom_term', ()))
    t.n_custom_terms = n_custom
    if n_custom:
        t.custom_term, t.custom_cutoff = dp('custom_term', np.int32), dp('custom_cutoff')
        t.custom_n_params = int(topo['custom_n_params'])
        t.custom_params = dp('custom_params')
        t.n_custom_groups = len(topo['custom_group_start']) - 1
        t.custom_group_start, t.custom_group_atoms = dp('custom_group_start', np.int32), dp('custom_group_atoms', np.int32)
        t.custom_group_weights = dp('custom_group_weights')
        t.n_custom_progs = len(topo['custom_prog_start']) - 1
        t.custom_prog_start, t.custom_code_op = dp('custom_prog_start', np.int32), dp('custom_code_op', np.int32)
        t.custom_code_arg = dp('custom_code_arg')
    return t, keep


def measure_fp32_peak(device=0):
    """Measured FP32 FMA throughput of the device in TFLOP/s (register-only microbenchmark inside the library)."""
    v = C.c_double()
    rc = load_library().bl_measure_fp32_peak(int(device), C.byref(v))
    if rc != 0:
        raise EngineError('bl_measure_fp32_peak failed (%d)' % rc)
    return v.value


class Engine(object):
    """One native handle: R independent walkers of one flattened system on one GPU."""

    created = 0          # native handles created by this process (bl_create calls that succeeded)

    def __init__(self, topo, device=0, n_replicas=1, seed=0):
        self.lib = load_library()
        self.topo = topo
        self.n_atoms = int(topo['n_atoms'])
        self.n_replicas = int(n_replicas)
        t, self._keep = build_topology(topo)
        h = _H()
        rc = self.lib.bl_create(C.byref(t), int(device), self.n_replicas, C.c_uint64(int(seed) & (2 ** 64 - 1)), C.byref(h))
        if rc != 0:
            raise EngineError('bl_create failed (%d): %s' % (rc, (self.lib.bl_last_error(None) or b'').decode()))
        self.h = h
        Engine.created += 1

    # -- plumbing -----------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = (self.lib.bl_last_error(self.h) or b'').decode()
            raise EngineError('%s (status %d)' % (msg, rc))

    def close(self):
        if getattr(self, 'h', None):
            self.lib.bl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- integrator ---------------------------------------------------------------------------------
    def set_ncmc_integrator(self, temperature, friction, timestep, splitting, nsteps_neq, nprop, prop_lambda_min,
                            prop_lambda_max, lambda_sterics, lambda_electrostatics, constraint_tol=1e-8):
        p = BlIntegratorParams()
        p.kind = BL_INTEGRATOR_NCMC
        p.temperature, p.friction, p.timestep, p.constraint_tol = temperature, friction, timestep, constraint_tol
        p.splitting = splitting.encode()
        p.nsteps_neq, p.nprop = int(nsteps_neq), int(nprop)
        p.prop_lambda_min, p.prop_lambda_max = float(prop_lambda_min), float(prop_lambda_max)
        ls, le = _arr(lambda_sterics, np.float64), _arr(lambda_electrostatics, np.float64)
        p.n_lambda = len(ls)
        p.lambda_sterics, p.lambda_electrostatics = _p(ls, C.c_double), _p(le, C.c_double)
        self._check(self.lib.bl_set_integrator(self.h, C.byref(p)))

    def set_langevin_integrator(self, temperature, friction, timestep, constraint_tol=1e-5):
        p = BlIntegratorParams()
        p.kind = BL_INTEGRATOR_LANGEVIN
        p.temperature, p.friction, p.timestep, p.constraint_tol = temperature, friction, timestep, constraint_tol
        p.splitting = None
        self._check(self.lib.bl_set_integrator(self.h, C.byref(p)))

    def set_seed(self, seed):
        self._check(self.lib.bl_set_seed(self.h, C.c_uint64(int(seed) & (2 ** 64 - 1))))

    # -- state ------------------------------------------------------------------------------------------
    def set_positions(self, xyz, replica=-1):
        a = _arr(xyz, np.float64).reshape(self.n_atoms, 3)
        self._check(self.lib.bl_set_positions(self.h, replica, _p(a, C.c_double)))

    def set_velocities(self, v, replica=-1):
        a = _arr(v, np.float64).reshape(self.n_atoms, 3)
        self._check(self.lib.bl_set_velocities(self.h, replica, _p(a, C.c_double)))

    def set_box(self, box):
        a = _arr(box, np.float64).reshape(3)
        self._check(self.lib.bl_set_box(self.h, _p(a, C.c_double)))

    def _get3(self, fn, replica):
        out = np.empty((self.n_atoms, 3), np.float64)
        self._check(fn(self.h, replica, _p(out, C.c_double)))
        return out

    def get_positions(self, replica=0):
        return self._get3(self.lib.bl_get_positions, replica)

    def get_velocities(self, replica=0):
        return self._get3(self.lib.bl_get_velocities, replica)

    def get_forces(self, replica=0):
        return self._get3(self.lib.bl_get_forces, replica)

    def get_box(self):
        out = np.empty(3, np.float64)
        self._check(self.lib.bl_get_box(self.h, _p(out, C.c_double)))
        return out

    def get_energy(self, potential=True, kinetic=True):
        ep = np.zeros(self.n_replicas, np.float64)
        ek = np.zeros(self.n_replicas, np.float64)
        self._check(self.lib.bl_get_energy(self.h, _p(ep, C.c_double) if potential else None,
                                           _p(ek, C.c_double) if kinetic else None))
        return ep, ek

    def get_energy_terms(self, replica=0):
        out = np.zeros(len(ENERGY_TERMS), np.float64)
        self._check(self.lib.bl_get_energy_terms(self.h, replica, _p(out, C.c_double)))
        return dict(zip(ENERGY_TERMS, out.tolist()))

    def copy_state_from(self, other, positions=True, velocities=True, box=True, mask=None):
        """Device-to-device copy of the walkers' state from another engine on the same GPU (all walkers, or those with
        mask[r] != 0; the box is shared by the walkers and only copied by the unmasked form)."""
        flags = (1 if positions else 0) | (2 if velocities else 0) | (4 if box else 0)
        if mask is None:
            self._check(self.lib.bl_copy_state(self.h, other.h, flags))
        else:
            m = _arr(mask, np.int32).reshape(self.n_replicas)
            self._check(self.lib.bl_copy_state_masked(self.h, other.h, flags & 3, _p(m, C.c_int32)))

    def velocities_to_temperature(self, temperature):
        self._check(self.lib.bl_velocities_to_temperature(self.h, float(temperature)))

    # -- globals ------------------------------------------------------------------------------------------
    def get_global(self, name, replica=0):
        v = C.c_double()
        self._check(self.lib.bl_get_global(self.h, replica, name.encode(), C.byref(v)))
        return v.value

    def set_global(self, name, value, replica=-1):
        self._check(self.lib.bl_set_global(self.h, replica, name.encode(), float(value)))

    def reset_ncmc(self):
        self._check(self.lib.bl_reset_ncmc(self.h))

    # -- hot path -----------------------------------------------------------------------------------------
    def _move(self, kind, step, atoms, masses=None, waters=None, center_atoms=None, center_masses=None, radius=0.0):
        m = BlMove()
        m.kind, m.step = int(kind), int(step)
        a = _arr(atoms, np.int32)
        keep = [a]
        m.n_atoms = len(a)
        m.atoms = _p(a, C.c_int32)
        if masses is not None:
            ms = _arr(masses, np.float64).reshape(-1)
            m.masses = _p(ms, C.c_double)
            keep.append(ms)
        if waters is not None:
            w = _arr(waters, np.int32).reshape(-1, len(a))
            m.n_waters, m.water_atoms = len(w), _p(w, C.c_int32)
            keep.append(w)
        if center_atoms is not None:
            ca = _arr(center_atoms, np.int32).reshape(-1)
            cm = _arr(center_masses, np.float64).reshape(-1)
            if len(ca) != len(cm):
                raise ValueError('center_atoms and center_masses differ in length')
            m.n_center, m.center_atoms, m.center_masses = len(ca), _p(ca, C.c_int32), _p(cm, C.c_double)
            keep += [ca, cm]
        m.radius = float(radius)
        return m, keep

    def ncmc_run(self, n_steps, move=None):
        """move: None or dict(kind=BL_MOVE_ROTATE, step=k, atoms=[...], masses=[...])"""
        if move is None:
            self._check(self.lib.bl_ncmc_run(self.h, int(n_steps), None))
        else:
            m, keep = self._move(**move)
            self._check(self.lib.bl_ncmc_run(self.h, int(n_steps), C.byref(m)))

    def md_run(self, n_steps):
        self._check(self.lib.bl_md_run(self.h, int(n_steps)))

    def apply_move(self, kind, atoms, masses=None, **water):
        """Apply a move now.  ``water``: waters=, center_atoms=, center_masses=, radius= for the BL_MOVE_WATER_* kinds."""
        m, keep = self._move(kind, 0, atoms, masses, **water)
        self._check(self.lib.bl_apply_move(self.h, C.byref(m)))

    def accept_reject(self, correction=None):
        acc = np.zeros(self.n_replicas, np.int32)
        logp = np.zeros(self.n_replicas, np.float64)
        logu = np.zeros(self.n_replicas, np.float64)
        corr = None if correction is None else _arr(correction, np.float64).reshape(self.n_replicas)
        self._check(self.lib.bl_accept_reject(self.h, None if corr is None else _p(corr, C.c_double),
                                              _p(acc, C.c_int32), _p(logp, C.c_double), _p(logu, C.c_double)))
        return acc, logp, logu

    def minimize(self, max_iterations=0, tolerance=10.0):
        self._check(self.lib.bl_minimize(self.h, int(max_iterations), float(tolerance)))

    # -- introspection --------------------------------------------------------------------------------------
    def neighbor_pairs(self, replica=0, capacity=None):
        cap = int(capacity or max(1024, self.n_atoms * 600))
        codes = np.empty(cap, np.int64)
        n = C.c_size_t()
        self._check(self.lib.bl_neighbor_pairs(self.h, replica, codes.ctypes.data_as(C.POINTER(C.c_int64)), cap, C.byref(n)))
        if n.value > cap:
            return self.neighbor_pairs(replica, n.value)
        return codes[:n.value].copy()

    def neighbor_stats(self, replica=0):
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.bl_neighbor_stats(self.h, replica, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self):
        return int(self.lib.bl_launch_count(self.h))

    def set_profiling(self, on):
        self._check(self.lib.bl_set_profiling(self.h, 1 if on else 0))

    def kernel_time(self, name):
        ms, n = C.c_double(), C.c_int64()
        self._check(self.lib.bl_get_kernel_time(self.h, KERNEL_IDS[name], C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def use_graphs(self, on):
        self._check(self.lib.bl_use_graphs(self.h, 1 if on else 0))

    def synchronize(self):
        self._check(self.lib.bl_synchronize(self.h))
