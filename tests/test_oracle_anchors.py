"""Analytic / statistical anchors for the parts of the oracle that have no golden vectors upstream:
Lennard-Jones minimum, Haar-uniform rotations, unit-normal thermostat noise, Maxwell–Boltzmann draws, equipartition
under the V R O R V Langevin splitting with constraints."""
import math
import os

import numpy as np
import pytest
from scipy import stats

from blues_b200 import unit as u
from blues_b200.structure import Structure
from oracle import ncmc_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def test_lennard_jones_minimum_and_lorentz_berthelot():
    """Two neutral atoms: E(2^(1/6) sigma_ij) = -eps_ij with sigma_ij the arithmetic and eps_ij the geometric mean,
    zero force there, E(sigma_ij) = 0."""
    from blues_b200.system import System, NonbondedForce, NoCutoff
    system = System(2)
    system.masses[:] = 39.9
    nb = NonbondedForce(2)
    nb.sigma[:] = [0.34, 0.30]
    nb.epsilon[:] = [1.0, 0.25]
    nb.method = NoCutoff
    system.addForce(nb)
    topo = system.flatten()
    ff = orc.ForceField(topo)
    sij, eij = 0.32, 0.5
    for r, e_want in ((2 ** (1 / 6) * sij, -eij), (sij, 0.0)):
        x = np.array([[0.0, 0.0, 0.0], [r, 0.0, 0.0]])
        e, f = ff.energy_forces(x, np.asarray(topo['box'], float))[:2]
        assert e == pytest.approx(e_want, abs=1e-12)
        if e_want:
            assert np.max(np.abs(f)) < 1e-9


def test_rotations_are_haar_uniform():
    """Shoemake's quaternion from three uniforms (the move's Philox stream): the rotation angle of a Haar-distributed
    rotation has density (1 - cos t) / pi on [0, pi], and a rotated fixed vector is uniform on the sphere."""
    n = 4000
    angles, z = [], []
    for k in range(n):
        uu = orc.philox_uniform4(123, orc.STREAM_MOVE, 0, k, [0])
        R = orc.rotation_matrix_from_quaternion(orc.quaternion_from_uniforms(uu[0][0], uu[1][0], uu[2][0]))
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.linalg.det(R) == pytest.approx(1.0)
        angles.append(math.acos(max(-1.0, min(1.0, 0.5 * (np.trace(R) - 1.0)))))
        z.append((np.array([0.0, 0.0, 1.0]) @ R)[2])
    cdf = lambda t: (t - np.sin(t)) / np.pi
    assert stats.kstest(angles, cdf).pvalue > 0.01
    assert stats.kstest(z, stats.uniform(loc=-1.0, scale=2.0).cdf).pvalue > 0.01        # Archimedes: z is uniform


def test_thermostat_noise_and_velocity_draws_are_unit_normal():
    xi = orc.philox_normal3(2024, orc.STREAM_LANGEVIN, 3, 17, 20000).reshape(-1)
    assert stats.kstest(xi, 'norm').pvalue > 0.01
    assert abs(np.mean(xi)) < 0.02 and np.var(xi) == pytest.approx(1.0, abs=0.03)
    # consecutive counters / replicas give uncorrelated sets
    a = orc.philox_normal3(2024, orc.STREAM_LANGEVIN, 3, 18, 20000).reshape(-1)
    b = orc.philox_normal3(2024, orc.STREAM_LANGEVIN, 4, 17, 20000).reshape(-1)
    assert abs(np.corrcoef(xi, a)[0, 1]) < 0.02 and abs(np.corrcoef(xi, b)[0, 1]) < 0.02


def test_equipartition_under_the_langevin_splitting_with_constraints():
    """<KE> = (3N - N_constraints) kT / 2 for the oracle's C twin running the `V R O R V` program (no alchemical
    change: no H step in the splitting) on the 35-atom vacuum fixture, and the Maxwell-Boltzmann draw starts it there."""
    from blues_b200.alchemy import AbsoluteAlchemicalFactory, AlchemicalRegion
    from oracle.c_oracle import COracle
    s = Structure.load_npz(os.path.join(GOLDEN, 'vac_divaline.npz'))
    system = s.createSystem(nonbondedMethod='NoCutoff', constraints='HBonds')
    topo = system.flatten()
    nsteps = 40000
    T = 300.0
    c = COracle(topo, None, None, 'V R O R V', T, 5.0, 0.001, nsteps, 1, 2.0, -1.0, seed=12345)
    c.set_state(s.coordinates * 0.1)
    c.velocities_to_temperature(T)
    mass = np.asarray(topo['mass'], float)
    ndof = 3 * topo['n_atoms'] - len(topo['constraints']) - (3 if topo['remove_cm'] else 0)
    ke0 = 0.5 * np.sum(mass[:, None] * c.v ** 2)
    kT = orc.KB * T
    assert ke0 == pytest.approx(0.5 * ndof * kT, rel=0.5)             # one draw of a chi-square with ~80 dof
    kes = []
    for k in range(nsteps // 10):
        c.step(10)
        if k >= 100:
            kes.append(0.5 * np.sum(mass[:, None] * c.v ** 2))
    # ~3900 samples 10 fs apart with a 0.2 ps thermostat: ~200 independent ones, sigma of the mean ~1.1 % -> 6 % is > 5 sigma
    assert np.mean(kes) == pytest.approx(0.5 * ndof * kT, rel=0.06)
    # the constraints hold throughout
    cons = np.asarray(topo['constraints']).reshape(-1, 2)
    d = np.linalg.norm(c.x[cons[:, 0]] - c.x[cons[:, 1]], axis=1)
    assert np.max(np.abs(d - topo['constraint_d'])) < 1e-6
