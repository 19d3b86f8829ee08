"""Test infrastructure: the interface of ``blues_b200._native.Engine`` answered by the CPU oracle.

Lets the *host* layer (``BLUESSimulation``, ``SimulationFactory``, moves, reporters — everything above the C ABI) run on
a machine without a GPU, so that the reference's own test files can exercise it unmodified
(``tests/test_reference_suite.py``).  The CUDA engine is compared with the same oracle in the ``-m gpu`` tests.  Never
imported by the package: ``blues_b200`` itself has no CPU path (``bl_create`` fails without a device).
"""
import math

import numpy as np

from oracle import ncmc_oracle as orc
from oracle.c_oracle import COracle


class _FastForceField(object):
    """``ForceField``'s call signature on top of the oracle's C twin (the two agree to round-off, tests/test_oracle.py)."""

    def __init__(self, topo):
        self.topo = dict(topo)
        self.n = int(topo['n_atoms'])
        self._c = None
        self._box = None

    def _oracle(self, box):
        box = np.asarray(box, float).reshape(-1)[:3]
        if self._c is None or not np.array_equal(box, self._box):
            t = dict(self.topo)
            t['box'] = box.copy()
            self._c, self._box = COracle(t), box.copy()
        return self._c

    def energy_forces(self, x, box, lam_s=1.0, lam_e=1.0):
        out = self._oracle(box).energy_forces(x, lam_s, lam_e)
        if len(self.topo.get('custom_term', ())):
            ec, fc = self._custom(np.asarray(x, float), np.asarray(box, float).reshape(-1)[:3], lam_s, lam_e)
            out = (out[0] + ec, out[1] + fc) + tuple(out[2:])
        return out

    def _custom(self, x, box, lam_s, lam_e):
        """Generic Custom*Force terms (the ``custom_*`` tables of ``bl_topology``): E(r) between the weighted centroids
        of two atom groups from the stack program, by the host interpreter of ``blues_b200/lepton.py``."""
        from blues_b200.lepton import evaluate_program
        t = self.topo
        gs, ga, gw = t['custom_group_start'], t['custom_group_atoms'], t['custom_group_weights']
        ps, op, arg = t['custom_prog_start'], t['custom_code_op'], t['custom_code_arg']
        par = np.asarray(t['custom_params'], float).reshape(len(t['custom_term']), -1)
        e, f = 0.0, np.zeros_like(x)
        for k, (a, b, prog, flags) in enumerate(np.asarray(t['custom_term']).reshape(-1, 4)):
            ia, wa = ga[gs[a]:gs[a + 1]], np.asarray(gw[gs[a]:gs[a + 1]], float)
            ib, wb = ga[gs[b]:gs[b + 1]], np.asarray(gw[gs[b]:gs[b + 1]], float)
            dv = wa @ x[ia] - wb @ x[ib]
            if flags & 1:
                dv -= box * np.round(dv / box)
            r = math.sqrt(float(dv @ dv))
            cut = float(t['custom_cutoff'][k])
            if cut > 0.0 and r >= cut:
                continue
            v, d = evaluate_program(list(op[ps[prog]:ps[prog + 1]]), list(arg[ps[prog]:ps[prog + 1]]), r, par[k],
                                    (lam_s, lam_e))
            e += v
            if r > 0.0 and d != 0.0:
                g = -d / r * dv
                np.add.at(f, ia, wa[:, None] * g)
                np.add.at(f, ib, -wb[:, None] * g)
        return e, f

    def energy(self, x, box, lam_s=1.0, lam_e=1.0):
        return self.energy_forces(x, box, lam_s, lam_e)[0]


class _TableNCMC(orc.NCMCOracle):
    """The oracle's integrator program driven by tabulated lambda functions, as the C ABI receives them."""

    def __init__(self, topo, ff, cons, tables, **kw):
        self._tables = tables
        orc.NCMCOracle.__init__(self, topo, **kw)
        self.ff, self.cons = ff, cons

    def _update_alch(self):
        k = int(round(self.g['lambda_'] * self.n_lambda_steps)) if hasattr(self, 'n_lambda_steps') else 0
        k = max(0, min(k, len(self._tables[0]) - 1))
        self.lam_s, self.lam_e = float(self._tables[0][k]), float(self._tables[1][k])


_GLOBAL_KEYS = {'lambda': 'lambda_'}


class OracleEngine(object):
    def __init__(self, topo, device=0, n_replicas=1, seed=0):
        if int(n_replicas) != 1:
            raise NotImplementedError('the oracle engine holds one walker')
        self.topo = topo
        self.n_atoms = int(topo['n_atoms'])
        self.n_replicas = 1
        self.seed = int(seed)
        self.x = np.zeros((self.n_atoms, 3))
        self.v = np.zeros((self.n_atoms, 3))
        self.box = np.asarray(topo['box'], float).reshape(-1)[:3].copy()
        self.ff = _FastForceField(topo)
        self.cons = orc.Constraints(topo)
        self.mass = np.asarray(topo['mass'], float)
        self.mobile = self.mass > 0
        self.invm = np.where(self.mobile, 1.0 / np.where(self.mobile, self.mass, 1.0), 0.0)
        self.integ = None
        self.kind = None
        self.vel_counter = self.move_counter = self.accept_counter = 0
        self._launches = 0
        self._water = None

    # -- plumbing ---------------------------------------------------------------------------------------
    def close(self):
        pass

    def synchronize(self):
        pass

    def set_profiling(self, on):
        pass

    def use_graphs(self, on):
        pass

    def launch_count(self):
        return self._launches

    def set_seed(self, seed):
        self.seed = int(seed)
        if self.integ is not None:
            self.integ.seed = self.seed

    # -- integrators --------------------------------------------------------------------------------------
    def set_ncmc_integrator(self, temperature, friction, timestep, splitting, nsteps_neq, nprop, prop_lambda_min,
                            prop_lambda_max, lambda_sterics, lambda_electrostatics, constraint_tol=1e-8):
        old = self.integ.g if self.kind == 'ncmc' else None
        self.integ = _TableNCMC(self.topo, self.ff, self.cons, (np.asarray(lambda_sterics, float),
                                                                  np.asarray(lambda_electrostatics, float)),
                                splitting=splitting, temperature=temperature, collision_rate=friction, timestep=timestep,
                                nsteps_neq=nsteps_neq, nprop=nprop, prop_lambda=0.3, seed=self.seed, replica=0)
        self.integ.prop_lambda_min, self.integ.prop_lambda_max = float(prop_lambda_min), float(prop_lambda_max)
        self.integ._update_alch()
        if old is not None:
            self.integ.g.update(old)
        self.kind = 'ncmc'

    def set_langevin_integrator(self, temperature, friction, timestep, constraint_tol=1e-5):
        self.integ = orc.LangevinMDOracle(self.topo, temperature, friction, timestep, self.seed, 0)
        self.integ.ff, self.integ.cons = self.ff, self.cons
        self.kind = 'md'

    def _push(self):
        self.integ.x, self.integ.v, self.integ.box = self.x.copy(), self.v.copy(), self.box.copy()

    def _pull(self):
        self.x, self.v = self.integ.x.copy(), self.integ.v.copy()
        if not np.all(np.isfinite(self.x)):
            from blues_b200._native import EngineError
            raise EngineError('Particle coordinate is nan (walker 0) (status -3)')

    def _lambdas(self):
        if self.kind == 'ncmc':
            return self.integ.lam_s, self.integ.lam_e
        return 1.0, 1.0

    # -- state ----------------------------------------------------------------------------------------------
    def set_positions(self, xyz, replica=-1):
        self.x = np.asarray(xyz, float).reshape(self.n_atoms, 3).copy()

    def set_velocities(self, v, replica=-1):
        self.v = np.asarray(v, float).reshape(self.n_atoms, 3).copy()

    def set_box(self, box):
        self.box = np.asarray(box, float).reshape(3).copy()

    def get_positions(self, replica=0):
        return self.x.copy()

    def get_velocities(self, replica=0):
        return self.v.copy()

    def get_box(self):
        return self.box.copy()

    def get_forces(self, replica=0):
        ls, le = self._lambdas()
        return np.array(self.ff.energy_forces(self.x, self.box, ls, le)[1], float)

    def get_energy(self, potential=True, kinetic=True):
        ls, le = self._lambdas()
        ep = self.ff.energy(self.x, self.box, ls, le) if potential else 0.0
        ek = 0.5 * float(np.sum(self.mass[:, None] * self.v * self.v)) if kinetic else 0.0
        return np.array([ep]), np.array([ek])

    def get_energy_terms(self, replica=0):
        from blues_b200._native import ENERGY_TERMS
        ls, le = self._lambdas()
        return dict(zip(ENERGY_TERMS, np.asarray(self.ff.energy_forces(self.x, self.box, ls, le)[2], float).tolist()))

    def copy_state_from(self, other, positions=True, velocities=True, box=True):
        if box:
            self.box = other.box.copy()
        if positions:
            self.x = other.x.copy()
        if velocities:
            self.v = other.v.copy()

    def velocities_to_temperature(self, temperature):
        xi = orc.philox_normal3(self.seed, orc.STREAM_VELOCITY, 0, self.vel_counter, self.n_atoms)
        self.vel_counter += 1
        self.v = np.sqrt(orc.KB * temperature * self.invm)[:, None] * xi
        self.v = self.cons.apply_velocities(self.x, self.v)

    def minimize(self, max_iterations=0, tolerance=10.0):
        """Capped steepest descent with the constraints re-imposed (what bl_minimize does)."""
        ls, le = self._lambdas()
        x = self.cons.apply_positions(self.x, self.x, tol=1e-10)
        e, f = self.ff.energy_forces(x, self.box, ls, le)[:2]
        step = 1e-5
        for _ in range(int(max_iterations) or 200):
            d = step * f * self.mobile[:, None]
            n = np.linalg.norm(d, axis=1, keepdims=True)
            d *= np.minimum(1.0, 0.005 / np.maximum(n, 1e-30))
            xn = self.cons.apply_positions(x + d, x, tol=1e-10)
            en, fn = self.ff.energy_forces(xn, self.box, ls, le)[:2]
            if np.isfinite(en) and en < e:
                x, e, f, step = xn, en, fn, step * 1.3
            else:
                step *= 0.4
            if np.sqrt(np.max(np.sum((f * self.mobile[:, None]) ** 2, axis=1))) < tolerance:
                break
        self.x = x

    # -- globals ----------------------------------------------------------------------------------------------
    def get_global(self, name, replica=0):
        it = self.integ
        if name == 'lambda_sterics':
            return it.lam_s
        if name == 'lambda_electrostatics':
            return it.lam_e
        extra = {'kT': it.kT, 'nsteps': getattr(it, 'nsteps', 0), 'n_lambda_steps': getattr(it, 'n_lambda_steps', 0),
                 'nprop': getattr(it, 'nprop', 1), 'prop_lambda_min': getattr(it, 'prop_lambda_min', 2.0),
                 'prop_lambda_max': getattr(it, 'prop_lambda_max', -1.0), 'n_rebuilds': 0}
        if name in extra:
            return float(extra[name])
        return float(it.g[_GLOBAL_KEYS.get(name, name)])

    def set_global(self, name, value, replica=-1):
        it = self.integ
        if name in ('nprop', 'prop_lambda_min', 'prop_lambda_max'):
            setattr(it, name, int(value) if name == 'nprop' else float(value))
            return
        key = _GLOBAL_KEYS.get(name, name)
        if key not in it.g:
            from blues_b200._native import EngineError
            raise EngineError("global variable '%s' cannot be set (status -1)" % name)
        it.g[key] = type(it.g[key])(value)
        if key in ('lambda_', 'lambda_step'):
            it._update_alch()

    def reset_ncmc(self):
        self.integ.reset()

    # -- the hot path ---------------------------------------------------------------------------------------------
    def ncmc_run(self, n_steps, move=None):
        n_steps = int(n_steps)
        self._launches += 10 * n_steps
        if move is not None and 0 <= int(move['step']) < n_steps:
            k = int(move['step'])
            self._run(k)
            m = dict(move)
            m.pop('step')
            self.apply_move(m.pop('kind'), m.pop('atoms'), m.pop('masses', None), **m)
            self._run(n_steps - k)
        else:
            self._run(n_steps)

    def _run(self, n):
        if n <= 0:
            return
        self._push()
        self.integ.step(n)
        self._pull()

    def md_run(self, n_steps):
        self._launches += 10 * int(n_steps)
        self._run(int(n_steps))

    def apply_move(self, kind, atoms, masses=None, **water):
        from blues_b200 import _native
        atoms = [int(a) for a in atoms]
        if kind == _native.BL_MOVE_ROTATE:
            u = orc.philox_uniform4(self.seed, orc.STREAM_MOVE, 0, self.move_counter, [0])
            self.move_counter += 1
            R = orc.rotation_matrix_from_quaternion(orc.quaternion_from_uniforms(u[0][0], u[1][0], u[2][0]))
            self.x = orc.rotate_ligand(self.x, atoms, np.asarray(masses, float), R)
            return
        centre = orc.center_of_mass_f32(self.x, water['center_atoms'], water['center_masses'])
        radius = float(water['radius'])
        if kind == _native.BL_MOVE_WATER_SWAP:
            inside = orc.waters_in_sphere(self.x, self.box, [list(w) for w in np.asarray(water['waters']).tolist()], centre, radius)
            u = orc.philox_uniform4(self.seed, orc.STREAM_MOVE, 0, self.move_counter, [1])[0][0]
            self.move_counter += 1
            self._water = (centre.astype(np.float32).astype(float), bool(inside))
            if inside:
                chosen = inside[min(int(u * len(inside)), len(inside) - 1)]
                self.x, self.v = orc.water_swap(self.x, self.v, atoms, chosen)
        elif kind == _native.BL_MOVE_WATER_TRANSLATE:
            c0, go = self._water if self._water else (centre, False)
            u = orc.philox_uniform4(self.seed, orc.STREAM_MOVE, 0, self.move_counter, [2])
            self.move_counter += 1
            if go:
                self.x = orc.water_translate(self.x, self.box, atoms, c0, radius, u[0][0], u[1][0], u[2][0])
        elif kind == _native.BL_MOVE_WATER_CHECK:
            go = self._water[1] if self._water else False
            self.integ.g['protocol_work'] = orc.water_after_move(self.x, self.box, atoms, centre, radius, go,
                                                                 self.integ.g['protocol_work'])
        else:
            raise ValueError('unknown move kind %r' % kind)

    def accept_reject(self, correction=None):
        w = self.integ.log_acceptance_probability()
        lu = math.log(orc.philox_uniform4(self.seed, orc.STREAM_ACCEPT, 0, self.accept_counter, [0])[0][0])
        self.accept_counter += 1
        corr = 0.0 if correction is None else float(np.asarray(correction).reshape(-1)[0])
        if not math.isnan(w):
            w = w + corr
        return np.array([1 if w > lu else 0], np.int32), np.array([w]), np.array([lu])
