"""CPU ORACLE (test infrastructure, NOT product code) — float64 numpy restatement of the BLUES NCMC hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module.  The product (``blues_b200``) never does: it fails loudly without its CUDA library.

What is restated, and from where (``file:line`` under the reference checkout):

* the CustomIntegrator *program* BLUES builds — ``blues/integrators.py:159-231`` (reset block, external work
  ``perturbed_pe − unperturbed_pe``, splitting pass, extra-propagation window, ``H`` step with
  ``protocol_work += Enew − Eold``), executed literally by :class:`NCMCOracle` with a full energy
  evaluation wherever the program reads ``energy`` (reference semantics: ≥ 3 evaluations per step);
* ``getLogAcceptanceProbability`` / ``reset`` — ``blues/integrators.py:233-249``;
* ``_get_prop_lambda`` — ``blues/integrators.py:147-157``; ``calculateNCMCSteps`` — ``blues/utils.py:89-145``;
* the BLUES iteration (sync, NCMC leg with move at ``moveStep``, alchemical correction, Metropolis test,
  velocity redraw, MD leg) — ``blues/simulation.py:1028-1257``;
* ``RandomLigandRotationMove.move`` — ``blues/moves.py:278-310``; ``WaterTranslationMove`` —
  ``blues/moves.py:951-1083``.

The arithmetic itself lives in third-party dependencies that are NOT in the reference checkout
(openmmtools==0.15.0 pinned at ``devtools/conda-recipe/meta.yaml:42``; OpenMM 7.3/7.4; parmed; mdtraj).  Their
published algorithms are restated here: AMBER bonded terms, Lennard-Jones +
Ewald direct space, smooth PME (order-5 B-splines), exclusion/self/plasma/dispersion terms, the
openmmtools softcore alchemical forms, Langevin V/R/O substeps with SHAKE/RATTLE constraints, OpenMM's
``LangevinIntegrator`` for the MD leg, and the Shoemake quaternion rotation.

PARITY PINNING: integrator bookkeeping (``_n_lambda_steps``, prop window, NCMC step arithmetic, the
999999 sentinel) and the statistical known answer of ``tests/test_ethylene.py:140-163`` (populations
0.25/0.75) are pinned by the reference's own tests.  Force/energy arithmetic, the constraint solver,
thermostat noise and quaternion sampling have NO golden vectors in the reference: for those this oracle
is **parity unpinned** against OpenMM and is anchored by analytic checks only (finite-difference
forces, Ewald α-independence, energy conservation, constraint residuals) — see DESIGN.md.

Units: nm, ps, dalton, kJ/mol, e, K.  Random numbers: Philox4x32-10 + Box–Muller, bit-identical to the
device generator (blues_b200/csrc/philox.cuh) so noisy trajectories can be compared step for step.
"""
import math
import numpy as np

ONE_4PI_EPS0 = 138.935456
KB = 1.3806504e-23 * 6.02214179e23 / 1000.0  # kJ/mol/K (simtk.unit CODATA-2006 values)
PME_ORDER = 5

try:
    from scipy.special import erfc as _erfc, erf as _erf
except Exception:  # pragma: no cover
    _erfc = np.vectorize(math.erfc)
    _erf = np.vectorize(math.erf)


# =========================================================================================================
# counter-based RNG (Philox4x32-10)
# =========================================================================================================
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)

STREAM_LANGEVIN, STREAM_VELOCITY, STREAM_MOVE, STREAM_ACCEPT, STREAM_MD = 0, 1, 2, 3, 4


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs broadcastable uint32 arrays → four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _M0 * c0
            p1 = _M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
            n0 = hi1 ^ c1 ^ np.uint64(k0)
            n2 = hi0 ^ c3 ^ np.uint64(k1)
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _u01(x):
    return (x.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def philox_uniform4(seed, stream, replica, counter, index):
    """Four U(0,1) doubles per ``index`` entry for (seed, stream, replica, counter)."""
    index = np.atleast_1d(np.asarray(index, np.uint32))
    r = philox4x32(index, np.uint32(counter & 0xFFFFFFFF), np.uint32(replica), np.uint32(stream),
                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return tuple(_u01(x) for x in r)


def philox_normal3(seed, stream, replica, counter, n):
    """(n,3) standard normals: Box–Muller on the four uniforms of atom i."""
    u0, u1, u2, u3 = philox_uniform4(seed, stream, replica, counter, np.arange(n, dtype=np.uint32))
    r0 = np.sqrt(-2.0 * np.log(u0))
    r1 = np.sqrt(-2.0 * np.log(u2))
    return np.stack([r0 * np.cos(2 * np.pi * u1), r0 * np.sin(2 * np.pi * u1), r1 * np.cos(2 * np.pi * u3)], axis=1)


# =========================================================================================================
# Lepton-subset evaluator for alchemical_functions (blues/simulation.py:654-659)
# =========================================================================================================
def eval_lambda_function(expr, lam):
    env = {'lambda': None, 'min': min, 'max': max, 'abs': abs, 'step': lambda x: 1.0 if x >= 0 else 0.0,
           'sqrt': math.sqrt, 'exp': math.exp, 'log': math.log, 'sin': math.sin, 'cos': math.cos,
           'select': lambda c, a, b: a if c != 0 else b, 'delta': lambda x: 1.0 if x == 0 else 0.0}
    code = str(expr).replace('^', '**').replace('lambda', 'lambda_')
    env['lambda_'] = float(lam)
    del env['lambda']
    return float(eval(code, {'__builtins__': {}}, env))


def get_prop_lambda(prop_lambda):
    """blues/integrators.py:147-157"""
    pmax = round(prop_lambda + 0.5, 4)
    pmin = round(0.5 - prop_lambda, 4)
    if pmax - pmin <= 0.0:
        pmin, pmax = 2.0, -1.0
    return pmin, pmax


def calculate_ncmc_steps(nstepsNC=0, nprop=1, propLambda=0.3):
    """blues/utils.py:89-145 (integer bookkeeping only)."""
    if nstepsNC % 2 != 0:
        rounded = nstepsNC & ~1
        if not rounded:
            raise SystemExit(1)
        nstepsNC = rounded
    lam_steps = nstepsNC / (2 * (nprop * propLambda + 0.5 - propLambda))
    lam_steps = int(lam_steps) if int(lam_steps) % 2 == 0 else int(lam_steps) + 1
    in_prop = int(nprop * (2 * math.floor(propLambda * lam_steps)))
    out_prop = int(2 * math.ceil((0.5 - propLambda) * lam_steps))
    prop_steps = in_prop + out_prop
    if prop_steps != nstepsNC:
        nstepsNC = lam_steps
    return {'nstepsNC': nstepsNC, 'propSteps': prop_steps, 'moveStep': int(nstepsNC / 2), 'nprop': nprop,
            'propLambda': propLambda}


# =========================================================================================================
# energies and forces
# =========================================================================================================
def _min_image(d, box, periodic):
    if periodic:
        d = d - box * np.round(d / box)
    return d


def _pairs_within(x, box, rc, periodic):
    n = len(x)
    if not periodic or rc is None:
        i, j = np.triu_indices(n, 1)
        return i.astype(np.int64), j.astype(np.int64)
    from scipy.spatial import cKDTree
    xw = x - box * np.floor(x / box)
    xw = np.where(xw >= box, xw - box, xw)
    tree = cKDTree(xw, boxsize=box)
    p = tree.query_pairs(rc, output_type='ndarray')
    return p[:, 0].astype(np.int64), p[:, 1].astype(np.int64)


def bspline_weights(frac):
    """Order-5 cardinal B-spline values and derivatives: weight of grid point base+k is M5(frac + 4 − k)."""
    n = PME_ORDER
    frac = np.asarray(frac, float)
    w = np.zeros(frac.shape + (n,))
    dw = np.zeros(frac.shape + (n,))

    def M(order, uu):
        out = np.zeros_like(uu)
        for k in range(order + 1):
            out += (-1) ** k * math.comb(order, k) * np.maximum(uu - k, 0.0) ** (order - 1)
        return out / math.factorial(order - 1)

    for k in range(n):
        uu = frac + (n - 1 - k)
        w[..., k] = M(n, uu)
        dw[..., k] = M(n - 1, uu) - M(n - 1, uu - 1.0)
    return w, dw


def bspline_moduli(K):
    n = PME_ORDER
    frac0 = np.zeros(1)
    w, _ = bspline_weights(frac0)          # M5 at integer nodes: w[0,k] = M5(4-k)
    data = w[0][::-1]                        # M5(0..4) → nodes 1..4 non-zero at data[1:]
    m = np.arange(K)
    arg = 2 * np.pi * np.outer(m, np.arange(n)) / K
    sc = (data[None, :] * np.cos(arg)).sum(1)
    ss = (data[None, :] * np.sin(arg)).sum(1)
    mod = sc * sc + ss * ss
    for i in range(K):
        if mod[i] < 1e-7:
            mod[i] = 0.5 * (mod[(i - 1) % K] + mod[(i + 1) % K])
    return mod


def pme_reciprocal(x, q, box, alpha, grid, want_forces=True):
    """Smooth PME reciprocal energy/forces (Essmann et al. 1995), float64, numpy FFT."""
    K = np.asarray(grid, int)
    n = len(x)
    sel = np.nonzero(q != 0)[0]
    F = np.zeros((n, 3))
    if len(sel) == 0:
        return 0.0, F
    xs, qs = x[sel], q[sel]
    uu = (xs / box)
    uu = (uu - np.floor(uu)) * K
    base = np.floor(uu).astype(int)
    frac = uu - base
    base = base % K
    w, dw = bspline_weights(frac)        # (ns,3,5)
    Q = np.zeros(K)
    ks = np.arange(PME_ORDER)
    gx = (base[:, 0, None] + ks) % K[0]
    gy = (base[:, 1, None] + ks) % K[1]
    gz = (base[:, 2, None] + ks) % K[2]
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    val = qs[:, None, None, None] * wx[:, :, None, None] * wy[:, None, :, None] * wz[:, None, None, :]
    idx = (gx[:, :, None, None] * K[1] + gy[:, None, :, None]) * K[2] + gz[:, None, None, :]
    np.add.at(Q.reshape(-1), idx.reshape(-1), val.reshape(-1))
    S = np.fft.fftn(Q)
    mx = np.fft.fftfreq(K[0], 1.0 / K[0]) / box[0]
    my = np.fft.fftfreq(K[1], 1.0 / K[1]) / box[1]
    mz = np.fft.fftfreq(K[2], 1.0 / K[2]) / box[2]
    m2 = mx[:, None, None] ** 2 + my[None, :, None] ** 2 + mz[None, None, :] ** 2
    bmod = bspline_moduli(K[0])[:, None, None] * bspline_moduli(K[1])[None, :, None] * bspline_moduli(K[2])[None, None, :]
    V = float(np.prod(box))
    with np.errstate(divide='ignore', invalid='ignore'):
        G = np.exp(-np.pi ** 2 * m2 / alpha ** 2) / (m2 * bmod * np.pi * V)
    G[0, 0, 0] = 0.0
    E = 0.5 * ONE_4PI_EPS0 * float(np.sum(G * (S.real ** 2 + S.imag ** 2)))
    if want_forces:
        phi = np.real(np.fft.ifftn(G * S)) * Q.size * ONE_4PI_EPS0   # potential on the grid
        ph = phi.reshape(-1)[idx]                                    # (ns,5,5,5)
        dwx, dwy, dwz = dw[:, 0], dw[:, 1], dw[:, 2]
        fx = np.einsum('nijk,ni,nj,nk->n', ph, dwx, wy, wz) * K[0] / box[0]
        fy = np.einsum('nijk,ni,nj,nk->n', ph, wx, dwy, wz) * K[1] / box[1]
        fz = np.einsum('nijk,ni,nj,nk->n', ph, wx, wy, dwz) * K[2] / box[2]
        F[sel] = -qs[:, None] * np.stack([fx, fy, fz], axis=1)
    return E, F


def softcore_sterics(r, sigma, eps, lam, alpha, a, b, c):
    """openmmtools softcore LJ: U = λ^a 4ε x(x−1), x = (σ/r_eff)^6, r_eff = σ[α(1−λ)^b + (r/σ)^c]^{1/c}; returns (U, dU/dr)."""
    rs = r / sigma
    s = alpha * (1.0 - lam) ** b + rs ** c
    x = s ** (-6.0 / c)
    U = lam ** a * 4.0 * eps * x * (x - 1.0)
    dx_dr = (-6.0 / c) * s ** (-6.0 / c - 1.0) * c * rs ** (c - 1.0) / sigma
    dU = lam ** a * 4.0 * eps * (2.0 * x - 1.0) * dx_dr
    return U, dU


class ForceField(object):
    """Potential energy of a flattened system (``System.flatten()`` dictionary) in float64."""

    def __init__(self, topo):
        t = self.t = topo
        self.n = int(t['n_atoms'])
        self.periodic = t['nb_method'] in (2, 4)
        self.pme = t['nb_method'] == 4
        self.cutoff = float(t['cutoff']) if t['nb_method'] != 0 else None
        self.alpha = float(t['ewald_alpha'])
        n = self.n
        ex = np.asarray(t['excl_pairs'], np.int64).reshape(-1, 2)
        self.excl_code = np.sort(ex[:, 0] * n + ex[:, 1]) if len(ex) else np.zeros(0, np.int64)
        self.alch = np.asarray(t['alch_atoms'], np.int64)
        self.is_alch = np.zeros(n, bool)
        self.is_alch[self.alch] = True
        self.alch_q = np.zeros(n)
        self.alch_sig = np.zeros(n)
        self.alch_eps = np.zeros(n)
        if len(self.alch):
            self.alch_q[self.alch] = t['alch_charge']
            self.alch_sig[self.alch] = t['alch_sigma']
            self.alch_eps[self.alch] = t['alch_eps']

    # -- bonded ------------------------------------------------------------------------------
    def bonded(self, x, box):
        t = self.t
        F = np.zeros_like(x)
        comp = {}
        per = self.periodic
        b = t['bonds']
        E = 0.0
        if len(b):
            d = _min_image(x[b[:, 1]] - x[b[:, 0]], box, per)
            r = np.linalg.norm(d, axis=1)
            dr = r - t['bond_r0']
            comp['bond'] = float(np.sum(0.5 * t['bond_k'] * dr * dr))
            f = (t['bond_k'] * dr / r)[:, None] * d
            np.add.at(F, b[:, 0], f)
            np.add.at(F, b[:, 1], -f)
        a = t['angles']
        if len(a):
            v1 = _min_image(x[a[:, 0]] - x[a[:, 1]], box, per)
            v2 = _min_image(x[a[:, 2]] - x[a[:, 1]], box, per)
            r1 = np.linalg.norm(v1, axis=1)
            r2 = np.linalg.norm(v2, axis=1)
            cs = np.clip(np.einsum('ij,ij->i', v1, v2) / (r1 * r2), -1.0, 1.0)
            th = np.arccos(cs)
            dth = th - t['angle_t0']
            comp['angle'] = float(np.sum(0.5 * t['angle_k'] * dth * dth))
            dE = t['angle_k'] * dth
            sn = np.sqrt(np.maximum(1.0 - cs * cs, 1e-30))
            g1 = (v2 / (r1 * r2)[:, None] - (cs / (r1 * r1))[:, None] * v1)   # d cos / d x_i
            g2 = (v1 / (r1 * r2)[:, None] - (cs / (r2 * r2))[:, None] * v2)   # d cos / d x_k
            f1 = (dE / sn)[:, None] * g1        # -dE/dx_i = dE * (1/sin) * dcos/dx
            f3 = (dE / sn)[:, None] * g2
            np.add.at(F, a[:, 0], f1)
            np.add.at(F, a[:, 2], f3)
            np.add.at(F, a[:, 1], -(f1 + f3))
        tt = t['torsions']
        if len(tt):
            p0, p1, p2, p3 = (x[tt[:, k]] for k in range(4))
            b1 = _min_image(p1 - p0, box, per)
            b2 = _min_image(p2 - p1, box, per)
            b3 = _min_image(p3 - p2, box, per)
            n1 = np.cross(b1, b2)
            n2 = np.cross(b2, b3)
            b2n = np.linalg.norm(b2, axis=1)
            m1 = np.cross(n1, b2 / b2n[:, None])
            xx = np.einsum('ij,ij->i', n1, n2)
            yy = np.einsum('ij,ij->i', m1, n2)
            phi = np.arctan2(yy, xx)
            nn = t['torsion_n']
            comp['torsion'] = float(np.sum(t['torsion_k'] * (1.0 + np.cos(nn * phi - t['torsion_phase']))))
            dE = -t['torsion_k'] * nn * np.sin(nn * phi - t['torsion_phase'])   # dE/dphi
            # standard analytic torsion gradient (Blondel & Karplus): arctan2 convention above gives the
            # IUPAC sign flipped relative to b1xb2 · b2xb3 orientation, handled by the sign of phi itself
            n1s = np.einsum('ij,ij->i', n1, n1)
            n2s = np.einsum('ij,ij->i', n2, n2)
            dphi_d0 = -(b2n / n1s)[:, None] * n1
            dphi_d3 = (b2n / n2s)[:, None] * n2
            s12 = np.einsum('ij,ij->i', b1, b2) / (b2n * b2n)
            s32 = np.einsum('ij,ij->i', b3, b2) / (b2n * b2n)
            dphi_d1 = (-1.0 - s12)[:, None] * dphi_d0 + s32[:, None] * dphi_d3
            dphi_d2 = (-1.0 - s32)[:, None] * dphi_d3 + s12[:, None] * dphi_d0
            # phi defined with m1 = n1 x b2̂ has the opposite handedness to the formulas above
            sgn = -1.0
            for k, g in enumerate((dphi_d0, dphi_d1, dphi_d2, dphi_d3)):
                np.add.at(F, tt[:, k], -(dE * sgn)[:, None] * g)
        ra = t['restraint_atoms']
        if len(ra):
            d = _min_image(x[ra] - t['restraint_x0'], box, per)
            comp['restraint'] = float(np.sum(t['restraint_k'] * np.einsum('ij,ij->i', d, d)))
            np.add.at(F, ra, -2.0 * t['restraint_k'][:, None] * d)
        E = sum(comp.values())
        return E, F, comp

    # -- nonbonded -----------------------------------------------------------------------------
    def nonbonded(self, x, box, lam_s=1.0, lam_e=1.0, neighbor_pairs=None):
        t = self.t
        n = self.n
        F = np.zeros_like(x)
        comp = {}
        per, rc, alpha = self.periodic, self.cutoff, self.alpha
        q, sig, eps = t['charge'], t['sigma'], t['epsilon']
        i, j = _pairs_within(x, box, rc, per) if neighbor_pairs is None else neighbor_pairs
        lo, hi = np.minimum(i, j), np.maximum(i, j)
        if len(self.excl_code):
            keep = ~np.isin(lo * n + hi, self.excl_code)
            i, j = lo[keep], hi[keep]
        else:
            i, j = lo, hi
        d = _min_image(x[i] - x[j], box, per)
        r2 = np.einsum('ij,ij->i', d, d)
        if rc is not None:
            m = r2 < rc * rc
            i, j, d, r2 = i[m], j[m], d[m], r2[m]
        r = np.sqrt(r2)
        self.last_pairs = (i, j)
        # environment (NonbondedForce) part: alchemical atoms carry q = 0, eps = 0 there
        s_ij = 0.5 * (sig[i] + sig[j])
        e_ij = np.sqrt(eps[i] * eps[j])
        with np.errstate(divide='ignore', invalid='ignore'):
            x6 = np.where(e_ij > 0, (s_ij / r) ** 6, 0.0)
        U_lj = 4.0 * e_ij * x6 * (x6 - 1.0)
        dU = -4.0 * e_ij * (12.0 * x6 * x6 - 6.0 * x6) / r
        qq = ONE_4PI_EPS0 * q[i] * q[j]
        if self.pme:
            ar = alpha * r
            U_c = qq * _erfc(ar) / r
            dU = dU + qq * (-_erfc(ar) / r2 - 2.0 * alpha / math.sqrt(math.pi) * np.exp(-ar * ar) / r)
        elif rc is None:
            U_c = qq / r
            dU = dU - qq / r2
        else:
            # reaction field (OpenMM CutoffPeriodic, eps_rf = 78.3)
            krf = (1.0 / rc ** 3) * (78.3 - 1.0) / (2.0 * 78.3 + 1.0)
            crf = (1.0 / rc) * 3.0 * 78.3 / (2.0 * 78.3 + 1.0)
            U_c = qq * (1.0 / r + krf * r2 - crf)
            dU = dU + qq * (-1.0 / r2 + 2.0 * krf * r)
        comp['lj'] = float(U_lj.sum())
        comp['coulomb_direct'] = float(U_c.sum())
        # alchemical pairs (softcore sterics + lambda-scaled direct-space electrostatics)
        if len(self.alch):
            ai, aj = self.is_alch[i], self.is_alch[j]
            am = ai | aj
            if np.any(am):
                ii, jj, rr, rr2 = i[am], j[am], r[am], r2[am]
                both = ai[am] & aj[am]
                sg = np.where(self.is_alch[ii], self.alch_sig[ii], sig[ii]) + np.where(self.is_alch[jj], self.alch_sig[jj], sig[jj])
                sg = 0.5 * sg
                ep = np.sqrt(np.where(self.is_alch[ii], self.alch_eps[ii], eps[ii]) *
                             np.where(self.is_alch[jj], self.alch_eps[jj], eps[jj]))
                qa = np.where(self.is_alch[ii], self.alch_q[ii], q[ii]) * np.where(self.is_alch[jj], self.alch_q[jj], q[jj])
                ls = np.where(both & (not t['annihilate_sterics']), 1.0, lam_s)
                le = np.where(both & (not t['annihilate_electrostatics']), 1.0, lam_e)
                sgs = np.where(ep > 0, sg, 1.0)
                Us, dUs = softcore_sterics(rr, sgs, ep, ls, t['softcore_alpha'], t['softcore_a'], t['softcore_b'],
                                           t['softcore_c'])
                Us = np.where(ep > 0, Us, 0.0)
                dUs = np.where(ep > 0, dUs, 0.0)
                kq = ONE_4PI_EPS0 * qa * le ** t['softcore_d']
                if self.pme:
                    ar = alpha * rr
                    Ue = kq * _erfc(ar) / rr
                    dUe = kq * (-_erfc(ar) / rr2 - 2.0 * alpha / math.sqrt(math.pi) * np.exp(-ar * ar) / rr)
                elif rc is None:
                    Ue = kq / rr
                    dUe = -kq / rr2
                else:
                    krf = (1.0 / rc ** 3) * (78.3 - 1.0) / (2.0 * 78.3 + 1.0)
                    crf = (1.0 / rc) * 3.0 * 78.3 / (2.0 * 78.3 + 1.0)
                    Ue = kq * (1.0 / rr + krf * rr2 - crf)
                    dUe = kq * (-1.0 / rr2 + 2.0 * krf * rr)
                comp['alch_sterics'] = float(Us.sum())
                comp['alch_electrostatics'] = float(Ue.sum())
                dU = dU.copy()
                dU[am] += dUs + dUe
        f = (-dU / r)[:, None] * d
        np.add.at(F, i, f)
        np.add.at(F, j, -f)
        # exceptions (1-4) and Ewald exclusion corrections over the explicit exclusion list
        ex = t['excl_pairs']
        if len(ex):
            a, b = ex[:, 0], ex[:, 1]
            d = _min_image(x[a] - x[b], box, per)
            r2 = np.einsum('ij,ij->i', d, d)
            r = np.sqrt(r2)
            eqq = ONE_4PI_EPS0 * t['excl_qq']
            es, ee = t['excl_sigma'], t['excl_eps']
            x6 = np.where(ee > 0, (es / r) ** 6, 0.0)
            U14 = 4.0 * ee * x6 * (x6 - 1.0) + eqq / r
            dU = -4.0 * ee * (12.0 * x6 * x6 - 6.0 * x6) / r - eqq / r2
            comp['exceptions'] = float(U14.sum())
            if self.pme:
                kqq = ONE_4PI_EPS0 * q[a] * q[b]
                ar = alpha * r
                comp['ewald_exclusion'] = float(np.sum(-kqq * _erf(ar) / r))
                dU = dU - kqq * (2.0 * alpha / math.sqrt(math.pi) * np.exp(-ar * ar) / r - _erf(ar) / r2)
            f = (-dU / r)[:, None] * d
            np.add.at(F, a, f)
            np.add.at(F, b, -f)
        ax = t['alch_exc_pairs']
        if len(ax):
            a, b = ax[:, 0], ax[:, 1]
            both = self.is_alch[a] & self.is_alch[b]
            d = _min_image(x[a] - x[b], box, per)
            r = np.linalg.norm(d, axis=1)
            ls = np.where(both & (not t['annihilate_sterics']), 1.0, lam_s)
            le = np.where(both & (not t['annihilate_electrostatics']), 1.0, lam_e)
            ee = t['alch_exc_eps']
            sgs = np.where(ee > 0, t['alch_exc_sigma'], 1.0)
            Us, dUs = softcore_sterics(r, sgs, ee, ls, t['softcore_alpha'], t['softcore_a'], t['softcore_b'],
                                       t['softcore_c'])
            Us = np.where(ee > 0, Us, 0.0)
            dUs = np.where(ee > 0, dUs, 0.0)
            kq = ONE_4PI_EPS0 * t['alch_exc_qq'] * le ** t['softcore_d']
            comp['alch_exceptions'] = float(np.sum(Us + kq / r))
            dU = dUs - kq / (r * r)
            f = (-dU / r)[:, None] * d
            np.add.at(F, a, f)
            np.add.at(F, b, -f)
        if self.pme:
            V = float(np.prod(box))
            Er, Fr = pme_reciprocal(x, q, box, alpha, t['pme_grid'])
            comp['pme_reciprocal'] = Er
            F += Fr
            comp['ewald_self'] = float(-ONE_4PI_EPS0 * alpha / math.sqrt(math.pi) * np.sum(q * q))
            Q = float(np.sum(q))
            comp['ewald_plasma'] = float(-ONE_4PI_EPS0 * math.pi * Q * Q / (2.0 * alpha * alpha * V))
        if t['use_dispersion_correction']:
            comp['dispersion'] = float(t['dispersion_coeff'] / np.prod(box))
        return sum(comp.values()), F, comp

    def energy_forces(self, x, box, lam_s=1.0, lam_e=1.0):
        Eb, Fb, cb = self.bonded(x, box)
        En, Fn, cn = self.nonbonded(x, box, lam_s, lam_e)
        cb.update(cn)
        return Eb + En, Fb + Fn, cb

    def energy(self, x, box, lam_s=1.0, lam_e=1.0):
        return self.energy_forces(x, box, lam_s, lam_e)[0]

    def neighbor_pairs(self, x, box, rc=None):
        """Sorted (i<j) pairs within ``rc`` (default: cutoff) that are not excluded — the bit-exact neighbour test."""
        rc = self.cutoff if rc is None else rc
        i, j = _pairs_within(x, box, rc, self.periodic)
        if rc is not None:
            d = _min_image(x[i] - x[j], box, self.periodic)
            m = np.einsum('ij,ij->i', d, d) < rc * rc      # strict, like the force loop
            i, j = i[m], j[m]
        lo, hi = np.minimum(i, j), np.maximum(i, j)
        code = lo * self.n + hi
        if len(self.excl_code):
            code = code[~np.isin(code, self.excl_code)]
        return np.sort(code)


# =========================================================================================================
# constraints: cluster-wise Newton (matrix SHAKE) and exact RATTLE
# =========================================================================================================
class Constraints(object):
    def __init__(self, topo):
        cons = np.asarray(topo['constraints'], np.int64).reshape(-1, 2)
        self.cons = cons
        self.d = np.asarray(topo['constraint_d'], float)
        self.invm = np.where(topo['mass'] > 0, 1.0 / np.where(topo['mass'] > 0, topo['mass'], 1.0), 0.0)
        n = int(topo['n_atoms'])
        parent = list(range(n))

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        for a, b in cons:
            ra, rb = find(int(a)), find(int(b))
            if ra != rb:
                parent[ra] = rb
        groups = {}
        for k, (a, b) in enumerate(cons):
            groups.setdefault(find(int(a)), []).append(k)
        by_size = {}
        for ks in groups.values():
            by_size.setdefault(len(ks), []).append(ks)
        self.batches = []
        for m, lst in by_size.items():
            ck = np.asarray(lst, np.int64)                  # (nc, m) constraint ids
            ia, ja = cons[ck, 0], cons[ck, 1]                # (nc, m)
            coef = np.zeros((len(lst), m, m))
            for a in range(m):
                for b in range(m):
                    coef[:, a, b] = ((ia[:, a] == ia[:, b]).astype(float) - (ia[:, a] == ja[:, b])) * self.invm[ia[:, a]] \
                        - ((ja[:, a] == ia[:, b]).astype(float) - (ja[:, a] == ja[:, b])) * self.invm[ja[:, a]]
            self.batches.append((ck, ia, ja, coef))

    def apply_positions(self, x, xref, box=None, tol=1e-13, max_iter=50):
        """Displace x along the xref constraint directions (mass-weighted) until |d|² matches."""
        x = x.copy()
        for ck, ia, ja, coef in self.batches:
            rref = xref[ia] - xref[ja]                       # (nc,m,3)
            d2 = self.d[ck] ** 2
            for _ in range(max_iter):
                s = x[ia] - x[ja]
                diff = d2 - np.einsum('cmk,cmk->cm', s, s)
                if np.max(np.abs(diff) / d2) < tol:
                    break
                A = 2.0 * coef * np.einsum('cak,cbk->cab', s, rref)
                lam = np.linalg.solve(A, diff[..., None])[..., 0]
                corr = lam[..., None] * rref                 # (nc,m,3)
                np.add.at(x, ia, corr * self.invm[ia][..., None])
                np.add.at(x, ja, -corr * self.invm[ja][..., None])
        return x

    def apply_velocities(self, x, v):
        v = v.copy()
        for ck, ia, ja, coef in self.batches:
            s = x[ia] - x[ja]
            rhs = -np.einsum('cmk,cmk->cm', s, v[ia] - v[ja])
            A = coef * np.einsum('cak,cbk->cab', s, s)
            mu = np.linalg.solve(A, rhs[..., None])[..., 0]
            corr = mu[..., None] * s
            np.add.at(v, ia, corr * self.invm[ia][..., None])
            np.add.at(v, ja, -corr * self.invm[ja][..., None])
        return v


# =========================================================================================================
# the integrator program
# =========================================================================================================
class NCMCOracle(object):
    """Literal interpreter of the program built by ``AlchemicalExternalLangevinIntegrator``
    (blues/integrators.py:159-231) on top of :class:`ForceField`/:class:`Constraints`."""

    def __init__(self, topo, alchemical_functions=None, splitting='H V R O R V H', temperature=300.0,
                 collision_rate=1.0, timestep=0.002, nsteps_neq=100, nprop=1, prop_lambda=0.3, seed=0, replica=0):
        self.topo = topo
        self.ff = ForceField(topo)
        self.cons = Constraints(topo)
        self.funcs = alchemical_functions or {
            'lambda_sterics': 'min(1, (1/0.3)*abs(lambda-0.5))',
            'lambda_electrostatics': 'step(0.2-lambda) - 1/0.2*lambda*step(0.2-lambda) + 1/0.2*(lambda-0.8)*step(lambda-0.8)'}
        self.splitting = splitting.split()
        self.kT = KB * temperature
        self.gamma = collision_rate
        self.dt = timestep
        self.nsteps = int(nsteps_neq)
        self.n_H = sum(1 for s in self.splitting if s == 'H')
        self.n_lambda_steps = self.nsteps * self.n_H
        self.counts = {k: sum(1 for s in self.splitting if s == k) for k in 'ORV'}
        self.prop_lambda_min, self.prop_lambda_max = get_prop_lambda(prop_lambda)
        self.nprop = nprop
        self.seed, self.replica = int(seed), int(replica)
        self.mass = np.asarray(topo['mass'], float)
        self.mobile = self.mass > 0
        self.invm = np.where(self.mobile, 1.0 / np.where(self.mobile, self.mass, 1.0), 0.0)
        self.box = np.asarray(topo['box'], float).copy()
        self.x = np.zeros((self.ff.n, 3))
        self.v = np.zeros((self.ff.n, 3))
        self.noise_counter = 0
        self.n_energy_evals = 0
        self.g = dict(step=0, lambda_=0.0, lambda_step=0, protocol_work=0.0, shadow_work=0.0, heat=0.0, first_step=0,
                      perturbed_pe=0.0, unperturbed_pe=0.0, prop=1, Eold=0.0, Enew=0.0, debug=0)
        self._update_alch()

    # globals -------------------------------------------------------------------------------
    def _update_alch(self):
        lam = self.g['lambda_']
        self.lam_s = eval_lambda_function(self.funcs.get('lambda_sterics', '1'), lam)
        self.lam_e = eval_lambda_function(self.funcs.get('lambda_electrostatics', '1'), lam)

    def energy(self):
        self.n_energy_evals += 1
        return self.ff.energy(self.x, self.box, self.lam_s, self.lam_e)

    def forces(self):
        self.n_energy_evals += 1
        return self.ff.energy_forces(self.x, self.box, self.lam_s, self.lam_e)[1]

    def kinetic_energy(self):
        return 0.5 * float(np.sum(self.mass[:, None] * self.v * self.v))

    def reset(self):
        """blues/integrators.py:240-249"""
        self.g.update(step=0, lambda_=0.0, protocol_work=0.0, shadow_work=0.0, first_step=0, perturbed_pe=0.0,
                      unperturbed_pe=0.0, prop=1, lambda_step=0)
        self._update_alch()

    def log_acceptance_probability(self):
        """blues/integrators.py:233-238"""
        return -1.0 * (self.g['protocol_work'] + self.g['shadow_work']) / self.kT

    # substeps ------------------------------------------------------------------------------
    def _constrain_x(self, xref):
        self.x = self.cons.apply_positions(self.x, xref)

    def _constrain_v(self):
        self.v = self.cons.apply_velocities(self.x, self.v)

    def _remove_cm(self):
        if self.topo['remove_cm']:
            p = (self.mass[:, None] * self.v).sum(0) / self.mass.sum()
            self.v[self.mobile] -= p

    def _V(self):
        h = self.dt / self.counts['V']
        self.v += h * self.forces() * self.invm[:, None]
        self._constrain_v()

    def _R(self):
        h = self.dt / self.counts['R']
        x0 = self.x.copy()
        self.x = self.x + h * self.v * self.mobile[:, None]
        x1 = self.x.copy()
        self._constrain_x(x0)
        self.v += (self.x - x1) / h * self.mobile[:, None]
        self._constrain_v()

    def _O(self):
        h = self.dt / self.counts['O']
        a = math.exp(-self.gamma * h)
        b = math.sqrt(1.0 - math.exp(-2.0 * self.gamma * h))
        old_ke = self.kinetic_energy()
        xi = philox_normal3(self.seed, STREAM_LANGEVIN, self.replica, self.noise_counter, self.ff.n)
        self.noise_counter += 1
        sigma = np.sqrt(self.kT * self.invm)
        self.v = (a * self.v + b * sigma[:, None] * xi) * self.mobile[:, None]
        self._constrain_v()
        self.g['heat'] += self.kinetic_energy() - old_ke

    def _H(self):
        g = self.g
        if g['prop'] != 1:
            return
        g['debug'] += 1
        g['Eold'] = self.energy()
        g['lambda_'] = (g['lambda_step'] + 1) / self.n_lambda_steps
        g['lambda_step'] += 1
        self._update_alch()
        g['Enew'] = self.energy()
        g['protocol_work'] += g['Enew'] - g['Eold']

    def _pass(self):
        self._remove_cm()      # UpdateContextState → CMMotionRemover
        for s in self.splitting:
            getattr(self, '_' + s)()

    def step(self, n=1):
        g = self.g
        for _ in range(n):
            if g['step'] == 0:
                g['perturbed_pe'] = g['unperturbed_pe'] = self.energy()
                self._constrain_x(self.x.copy())
                self._constrain_v()
                g['protocol_work'] = 0.0
                g['lambda_'] = 0.0
                g['lambda_step'] = 0
                self._update_alch()
            if g['step'] < self.nsteps:
                g['perturbed_pe'] = self.energy()
                if g['first_step'] < 1:
                    g['first_step'] = 1
                    g['unperturbed_pe'] = g['perturbed_pe']
                g['protocol_work'] += g['perturbed_pe'] - g['unperturbed_pe']
                self._pass()
                if g['lambda_'] > self.prop_lambda_min and g['lambda_'] <= self.prop_lambda_max:
                    while g['prop'] < self.nprop:
                        g['prop'] += 1
                        self._pass()
                g['unperturbed_pe'] = self.energy()
                g['step'] += 1
                g['prop'] = 1

    # state helpers -----------------------------------------------------------------------------
    def set_velocities_to_temperature(self, temperature, counter):
        xi = philox_normal3(self.seed, STREAM_VELOCITY, self.replica, counter, self.ff.n)
        self.v = np.sqrt(KB * temperature * self.invm)[:, None] * xi
        self._constrain_v()


class LangevinMDOracle(object):
    """OpenMM ``LangevinIntegrator`` step (MD leg, blues/simulation.py:1189-1213)."""

    def __init__(self, topo, temperature=300.0, friction=1.0, timestep=0.002, seed=0, replica=0):
        self.topo = topo
        self.ff = ForceField(topo)
        self.cons = Constraints(topo)
        self.kT = KB * temperature
        self.gamma, self.dt = friction, timestep
        self.seed, self.replica = int(seed), int(replica)
        self.mass = np.asarray(topo['mass'], float)
        self.mobile = self.mass > 0
        self.invm = np.where(self.mobile, 1.0 / np.where(self.mobile, self.mass, 1.0), 0.0)
        self.box = np.asarray(topo['box'], float).copy()
        self.x = np.zeros((self.ff.n, 3))
        self.v = np.zeros((self.ff.n, 3))
        self.noise_counter = 0

    def step(self, n=1):
        vs = math.exp(-self.dt * self.gamma)
        fs = (1.0 - vs) / self.gamma if self.gamma > 0 else self.dt
        ns = math.sqrt(self.kT * (1.0 - vs * vs))
        for _ in range(n):
            if self.topo['remove_cm']:
                p = (self.mass[:, None] * self.v).sum(0) / self.mass.sum()
                self.v[self.mobile] -= p
            F = self.ff.energy_forces(self.x, self.box)[1]
            xi = philox_normal3(self.seed, STREAM_MD, self.replica, self.noise_counter, self.ff.n)
            self.noise_counter += 1
            self.v = (vs * self.v + fs * F * self.invm[:, None] + ns * np.sqrt(self.invm)[:, None] * xi) * self.mobile[:, None]
            x0 = self.x
            x1 = x0 + self.dt * self.v
            x1 = self.cons.apply_positions(x1, x0)
            self.v = (x1 - x0) / self.dt
            self.x = x1


# =========================================================================================================
# moves (blues/moves.py)
# =========================================================================================================
def quaternion_from_uniforms(u0, u1, u2):
    """Shoemake: Haar-uniform unit quaternion (w, x, y, z) from three U(0,1)."""
    s1, s2 = math.sqrt(1.0 - u0), math.sqrt(u0)
    return np.array([s1 * math.sin(2 * math.pi * u1), s1 * math.cos(2 * math.pi * u1),
                     s2 * math.sin(2 * math.pi * u2), s2 * math.cos(2 * math.pi * u2)])


def rotation_matrix_from_quaternion(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rotate_ligand(x, atom_indices, masses, R):
    """blues/moves.py:292-307: COM in float32 with the supplied masses, x' = (x − c)·R + c (row vector × R)."""
    pos = x[atom_indices]
    c32 = np.asarray(pos, np.float32)
    m32 = np.asarray(masses, np.float32).reshape(-1)
    com = ((c32 * m32[:, None]).sum(axis=0) / m32.sum()).astype(np.float64)
    out = x.copy()
    out[atom_indices] = np.dot(pos - com, R) + com
    return out


def random_sphere_point(radius, origin, u_r, u_phi, u_cos):
    """blues/moves.py:898-918 with explicit uniforms."""
    r = radius * u_r ** (1.0 / 3.0)
    phi = 2 * math.pi * u_phi
    ct = 2.0 * u_cos - 1.0
    th = math.acos(ct)
    return np.array([math.sin(th) * math.cos(phi), math.sin(th) * math.sin(phi), math.cos(th)]) * r + origin


# ---- WaterTranslationMove (blues/moves.py:846-1083) ------------------------------------------------------------
def center_of_mass_f32(x, atoms, masses):
    """blues/moves.py:921-949: float32 coordinates and masses, mass-weighted mean."""
    c32 = np.asarray(x, np.float32)[np.asarray(atoms)]
    m32 = np.asarray(masses, np.float32).reshape(-1)
    return ((c32 * m32[:, None]).sum(axis=0) / m32.sum()).astype(np.float64)


def periodic_distance_f32(p, c, box):
    """mdtraj.compute_distances(periodic=True) on an orthorhombic cell: float32 minimum-image distance."""
    d = np.asarray(p, np.float32) - np.asarray(c, np.float32)
    if box is not None:
        b = np.asarray(box, np.float32)
        d = d - b * np.round(d / b)
    return float(np.sqrt(np.sum(d * d)))


def waters_in_sphere(x, box, waters, center, radius):
    """Candidate waters of beforeMove (blues/moves.py:981-991): first atom within `radius` of the centre, residue order."""
    return [list(w) for w in waters if periodic_distance_f32(x[w[0]], center, box) <= radius]


def water_swap(x, v, alch, chosen):
    """blues/moves.py:993-999: the chosen water and the alchemical water trade positions and velocities."""
    x, v = x.copy(), v.copy()
    a, b = list(alch), list(chosen)
    x[a], x[b] = x[b].copy(), x[a].copy()
    v[a], v[b] = v[b].copy(), v[a].copy()
    return x, v


def water_translate(x, box, alch, center, radius, u_r, u_phi, u_cos):
    """blues/moves.py:1009-1053: nothing happens if the alchemical water's first atom is at or beyond the radius;
    otherwise the molecule is translated so that this atom sits on the random point of the sphere."""
    if periodic_distance_f32(x[alch[0]], center, box) >= radius:
        return x.copy()
    target = random_sphere_point(radius, np.asarray(center, float), u_r, u_phi, u_cos)
    out = x.copy()
    out[list(alch)] = x[list(alch)] - (x[alch[0]] - target)
    return out


def water_after_move(x, box, alch, center, radius, go, protocol_work):
    """blues/moves.py:1055-1083: outside the sphere after the protocol and the move was on -> work = 999999."""
    if periodic_distance_f32(x[alch[0]], center, box) > radius and go:
        return 999999.0
    return protocol_work


# ---- MonteCarloBarostat of the MD leg (blues/simulation.py:602-626 attaches it; algorithm: OpenMM
# MonteCarloBarostatImpl::updateContextState + ReferenceMonteCarloBarostat::applyBarostat) ---------------------------
def mc_barostat_trial(x, box, molecules, energy_fn, pressure_bar, temperature, volume_scale, u_vol, u_acc):
    """One isotropic volume move.  molecules: list of atom-index lists; energy_fn(x, box) -> kJ/mol.
    Returns (accepted, x, box, w)."""
    box = np.asarray(box, float)
    volume = box[0] * box[1] * box[2]
    d_volume = volume_scale * 2.0 * (u_vol - 0.5)
    new_volume = volume + d_volume
    scale = (new_volume / volume) ** (1.0 / 3.0)
    e0 = energy_fn(x, box)
    xn = x.copy()
    for atoms in molecules:
        centre = x[atoms].sum(axis=0) / len(atoms)
        wrapped = centre - np.floor(centre / box) * box           # into the first periodic box
        xn[atoms] = x[atoms] + (wrapped * scale - centre)
    new_box = box * scale
    e1 = energy_fn(xn, new_box)
    kT = 0.0083144626181532 * temperature
    pressure = pressure_bar * 6.02214076e23 * 1e-25                # bar -> kJ/mol/nm^3
    w = e1 - e0 + pressure * d_volume - len(molecules) * kT * math.log(new_volume / volume)
    if w > 0.0 and u_acc > math.exp(-w / kT):
        return False, x, box, w
    return True, xn, new_box, w


def alchemical_correction(e_ncmc0, e_md0, e_alch1, e_ncmc1, kT):
    """blues/simulation.py:1100-1119"""
    return (e_ncmc0 - e_md0 + e_alch1 - e_ncmc1) * (-1.0 / kT)


def metropolis_accept(log_accept, correction, log_u):
    """blues/simulation.py:1130-1140: NaN work is never accepted; correction is skipped for NaN work."""
    w = log_accept
    if not math.isnan(w):
        w = w + correction
    return bool(w > log_u)
