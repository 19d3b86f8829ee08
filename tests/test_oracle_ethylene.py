"""The reference's only known-answer test, reproduced with the ORACLE's integrator program (CPU).

blues/tests/test_ethylene.py:140-163: a charged ethylene pinned between two fixed sites, NCMC rotation moves (20 NCMC +
20 MD steps, 200 K, 1 fs); the populations of the two orientations must come out 0.25 / 0.75.  The system uses a
CustomNonbondedForce whose sigma / epsilon are scaled by lambda_sterics / lambda_electrostatics and a
CustomCentroidBondForce (tests/data/ethylene_system.xml), written out here as a small force-field class; everything else
— the H V R O R V H program with work accumulation, the rotation move, the alchemical correction and Metropolis rule,
the velocity redraw and the Langevin MD leg — is the oracle's code, which this test therefore pins to the reference.
"""
import json
import math
import os

import numpy as np

from oracle import ncmc_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ELEMENT_MASS = {6: 12.01078, 1: 1.007947, 7: 14.00672}


class EthyleneForceField(object):
    def __init__(self, fx):
        self.n = len(fx['mass'])
        self.bonds = np.array([b[:2] for b in fx['bonds']], int)
        self.bond_r0 = np.array([b[2] for b in fx['bonds']])
        self.bond_k = np.array([b[3] for b in fx['bonds']])
        self.angles = np.array([a[:3] for a in fx['angles']], int)
        self.angle_t0 = np.array([a[3] for a in fx['angles']])
        self.angle_k = np.array([a[4] for a in fx['angles']])
        self.tors = np.array([t[:4] for t in fx['torsions']], int)
        self.tors_n = np.array([t[4] for t in fx['torsions']])
        self.tors_phase = np.array([t[5] for t in fx['torsions']])
        self.tors_k = np.array([t[6] for t in fx['torsions']])
        cn = fx['custom_nonbonded']
        par = np.array(cn['params'])
        i, j = np.meshgrid(cn['set1'], cn['set2'], indexing='ij')
        self.pi, self.pj = i.ravel(), j.ravel()
        self.sig = 0.5 * (par[self.pi, 0] + par[self.pj, 0])
        self.eps = np.sqrt(par[self.pi, 1] * par[self.pj, 1])
        self.qq = par[self.pi, 2] * par[self.pj, 2]
        self.g1, self.g2 = fx['centroid_bond']['groups']
        self.kc = fx['centroid_bond']['k']
        m = np.array(fx['mass'])
        self.w1 = np.ones(len(self.g1)) / len(self.g1)                  # explicit weight="1" on both particles
        self.w2 = m[self.g2] / m[self.g2].sum()                          # default weights: masses
        # reuse the oracle's bonded code through a minimal flat topology
        self.t = {'bonds': self.bonds, 'bond_k': self.bond_k, 'bond_r0': self.bond_r0, 'angles': self.angles,
                  'angle_k': self.angle_k, 'angle_t0': self.angle_t0, 'torsions': self.tors, 'torsion_k': self.tors_k,
                  'torsion_n': self.tors_n, 'torsion_phase': self.tors_phase, 'restraint_atoms': np.zeros(0, int),
                  'n_atoms': self.n, 'nb_method': 0, 'cutoff': 1.0, 'ewald_alpha': 0.0, 'excl_pairs': np.zeros((0, 2), int),
                  'alch_atoms': np.zeros(0, int)}
        self._bonded = orc.ForceField(self.t)

    def energy_forces(self, x, box, lam_s=1.0, lam_e=1.0):
        E, F, _ = self._bonded.bonded(x, box)
        d = x[self.pi] - x[self.pj]
        r2 = np.einsum('ij,ij->i', d, d)
        r = np.sqrt(r2)
        sig, eps = self.sig * lam_s, self.eps * lam_e
        with np.errstate(divide='ignore', invalid='ignore'):
            s6 = np.where(sig > 0, (sig / r) ** 6, 0.0)
        E += float(np.sum(self.qq / r2 + 4 * eps * (s6 * s6 - s6)))
        dU = -2 * self.qq / (r2 * r) - 4 * eps * (12 * s6 * s6 - 6 * s6) / r
        f = (-dU / r)[:, None] * d
        np.add.at(F, self.pi, f)
        np.add.at(F, self.pj, -f)
        c = self.w1 @ x[self.g1] - self.w2 @ x[self.g2]
        E += 0.5 * self.kc * float(c @ c)
        for a, w in zip(self.g1, self.w1):
            F[a] -= self.kc * w * c
        for a, w in zip(self.g2, self.w2):
            F[a] += self.kc * w * c
        return E, F, {}

    def energy(self, x, box, lam_s=1.0, lam_e=1.0):
        return self.energy_forces(x, box, lam_s, lam_e)[0]


def _fixture():
    fx = json.load(open(os.path.join(GOLDEN, 'ethylene.json')))
    cons = np.array([c[:2] for c in fx['constraints']], int)
    topo = {'n_atoms': len(fx['mass']), 'mass': np.array(fx['mass']), 'constraints': cons,
            'constraint_d': np.array([c[2] for c in fx['constraints']]), 'box': np.array([2.0, 2.0, 2.0]),
            'remove_cm': 0, 'nb_method': 0, 'cutoff': 1.0, 'ewald_alpha': 0.0, 'excl_pairs': np.zeros((0, 2), int),
            'alch_atoms': np.zeros(0, int)}
    return fx, topo


def test_force_field_gradient():
    fx, topo = _fixture()
    ff = EthyleneForceField(fx)
    x = np.array(fx['positions_nm']) + 0.01 * np.random.RandomState(0).randn(8, 3)
    E, F, _ = ff.energy_forces(x, topo['box'], 0.7, 0.4)
    for a, k in ((2, 0), (4, 1), (7, 2)):
        xp, xm = x.copy(), x.copy()
        xp[a, k] += 1e-6
        xm[a, k] -= 1e-6
        fd = -(ff.energy(xp, topo['box'], 0.7, 0.4) - ff.energy(xm, topo['box'], 0.7, 0.4)) / 2e-6
        assert abs(fd - F[a, k]) < 1e-4 * max(1.0, abs(F[a, k]))


def run_blues(seed, n_iter=100, nsteps=20, T=200.0, dt=0.001):
    fx, topo = _fixture()
    ff = EthyleneForceField(fx)
    rng = np.random.RandomState(seed)
    lig = np.arange(2, 8)
    masses = np.array([ELEMENT_MASS[z] for z in np.array(fx['atomic_numbers'])[lig]])
    ncmc = orc.NCMCOracle(topo, None, 'H V R O R V H', T, 1.0, dt, nsteps, 1, 0.3, seed, 0)
    ncmc.ff = ff
    md = orc.LangevinMDOracle(topo, T, 1.0, dt, seed, 1)
    md.ff = ff
    md.x = np.array(fx['positions_nm'])
    xi = orc.philox_normal3(seed, orc.STREAM_VELOCITY, 1, 0, 8)
    md.v = md.cons.apply_velocities(md.x, np.sqrt(orc.KB * T * md.invm)[:, None] * xi)
    kT = orc.KB * T
    dists, accepted = [], 0
    for it in range(n_iter):
        e_md0 = ff.energy(md.x, topo['box'])
        ncmc.x, ncmc.v = md.x.copy(), md.v.copy()                        # _syncStatesMDtoNCMC
        e_nc0 = ncmc.energy()
        for step in range(nsteps):
            if step == nsteps // 2:                                      # move at moveStep
                q = orc.quaternion_from_uniforms(*rng.uniform(0, 1, 3))
                ncmc.x = orc.rotate_ligand(ncmc.x, lig, masses, orc.rotation_matrix_from_quaternion(q))
            ncmc.step(1)
        e_nc1 = ncmc.energy()
        corr = orc.alchemical_correction(e_nc0, e_md0, ff.energy(ncmc.x, topo['box']), e_nc1, kT)
        if orc.metropolis_accept(ncmc.log_acceptance_probability(), corr, math.log(rng.random_sample())):
            md.x = ncmc.x.copy()                                         # positions only
            accepted += 1
        ncmc.reset()
        xi = orc.philox_normal3(seed, orc.STREAM_VELOCITY, 1, it + 1, 8)  # setVelocitiesToTemperature
        md.v = md.cons.apply_velocities(md.x, np.sqrt(orc.KB * T * md.invm)[:, None] * xi)
        for blk in range(nsteps // 5):
            md.step(5)
            dists.append(np.linalg.norm(md.x[0] - md.x[2]))
    return np.array(dists), accepted


def test_two_state_populations_match_reference_known_answer():
    """blues/tests/test_ethylene.py:151-163: populations of dist(atom0, atom2) <= 0.49 nm vs > 0.49 nm = [0.25, 0.75]."""
    freqs, acc = [], 0
    for seed in (11, 23, 37, 41, 59):
        d, a = run_blues(seed)
        acc += a
        freqs.append([np.mean(d <= 0.49), np.mean(d > 0.49)])
    avg = np.mean(freqs, axis=0)
    err = np.std(freqs, axis=0) / math.sqrt(len(freqs))
    print('populations', avg, '+-', err, 'accepted', acc, 'of 500')
    assert acc > 50                                                       # NCMC rotations are accepted regularly
    assert abs(avg[0] - 0.25) < max(0.06, 2.5 * err[0])
    assert abs(avg[1] - 0.75) < max(0.06, 2.5 * err[1])
