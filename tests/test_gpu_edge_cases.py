"""Edge cases of the C ABI on the CUDA engine: the smallest systems, ragged sizes, empty tables."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _system(n, masses, box=None):
    from blues_b200.system import System, NonbondedForce, NoCutoff, PME, CMMotionRemover
    s = System(n)
    s.masses = np.asarray(masses, float)
    nb = NonbondedForce(n)
    s.addForce(nb)
    if box is not None:
        s.box = np.asarray(box, float)
        nb.method = PME
        nb.cutoff = 0.9
    else:
        nb.method = NoCutoff
    return s, nb


def test_single_free_particle_drifts_and_thermalises():
    from blues_b200 import _native
    s, nb = _system(1, [12.0])
    eng = _native.Engine(s.flatten(), n_replicas=3, seed=5)
    eng.set_langevin_integrator(300.0, 1.0, 0.002)
    eng.set_positions(np.zeros((1, 3)))
    eng.velocities_to_temperature(300.0)
    v = np.array([eng.get_velocities(r)[0] for r in range(3)])
    assert np.all(np.isfinite(v)) and len({tuple(np.round(x, 12)) for x in v}) == 3      # three walkers, three draws
    eng.md_run(50)
    x = np.array([eng.get_positions(r)[0] for r in range(3)])
    assert np.all(np.isfinite(x)) and np.all(np.abs(x) < 5.0) and np.any(np.abs(x) > 1e-4)
    ep, ek = eng.get_energy()
    assert np.allclose(ep, 0.0) and np.all(ek > 0)
    eng.close()


def test_two_ions_in_a_periodic_box_match_the_oracle():
    """N = 2 (one 32-atom block mostly padding, one atom per cell at most): PME energy and forces against the oracle."""
    from blues_b200 import _native
    from oracle.ncmc_oracle import ForceField
    s, nb = _system(2, [22.99, 35.45], box=[2.4, 2.5, 2.6])
    nb.charge[:] = [1.0, -1.0]
    nb.sigma[:] = [0.33, 0.44]
    nb.epsilon[:] = [0.0116, 0.4184]
    x = np.array([[0.3, 0.4, 0.5], [0.62, 0.55, 0.71]])
    topo = s.flatten()
    eng = _native.Engine(topo, n_replicas=1, seed=1)
    eng.set_langevin_integrator(300.0, 1.0, 0.002)
    eng.set_positions(x)
    E = eng.get_energy()[0][0]
    F = eng.get_forces(0)
    Eo, Fo, _ = ForceField(topo).energy_forces(x, topo['box'], 1.0, 1.0)
    assert abs(E - Eo) < 2e-4 * max(1.0, abs(Eo)), (E, Eo)
    assert np.max(np.abs(F - Fo)) < 2e-4 * np.max(np.abs(Fo)), (F, Fo)
    assert np.allclose(F[0], -F[1], rtol=1e-3, atol=1e-3)
    eng.md_run(20)                                        # lists, PME and the integrator on a 2-atom system
    assert np.all(np.isfinite(eng.get_positions(0)))
    eng.close()


def test_single_rigid_water_keeps_its_geometry():
    from blues_b200 import _native
    s, nb = _system(3, [15.999, 1.008, 1.008])
    nb.charge[:] = [-0.834, 0.417, 0.417]
    nb.sigma[:] = [0.315, 0.1, 0.1]
    nb.epsilon[:] = [0.636, 0.0, 0.0]
    for i, j in ((0, 1), (0, 2), (1, 2)):
        nb.addException(i, j, 0.0, 0.1, 0.0)
    doh, dhh = 0.09572, 0.15139
    s.addConstraint(0, 1, doh); s.addConstraint(0, 2, doh); s.addConstraint(1, 2, dhh)
    h = np.sqrt(doh ** 2 - (dhh / 2) ** 2)
    x = np.array([[0.0, 0.0, 0.0], [dhh / 2, h, 0.0], [-dhh / 2, h, 0.0]])
    eng = _native.Engine(s.flatten(), n_replicas=1, seed=2)
    eng.set_langevin_integrator(300.0, 1.0, 0.002)
    eng.set_positions(x)
    eng.velocities_to_temperature(300.0)
    eng.md_run(200)
    y = eng.get_positions(0)
    d = [np.linalg.norm(y[i] - y[j]) for i, j in ((0, 1), (0, 2), (1, 2))]
    assert np.allclose(d, [doh, doh, dhh], rtol=1e-5)        # OpenMM LangevinIntegrator default constraint tolerance 1e-5
    assert np.linalg.norm(y - x) > 1e-3                    # it moved (tumbled) as a rigid body
    eng.close()


def test_ragged_walker_counts_and_atom_counts_around_the_block_size():
    """31, 32, 33 atoms (the sorted mirrors are padded to blocks of 32) x 1, 2, 5 walkers: identical walkers stay identical."""
    from blues_b200 import _native
    rng = np.random.RandomState(0)
    for n in (31, 32, 33):
        s, nb = _system(n, np.full(n, 39.9), box=[2.2, 2.2, 2.2])
        nb.sigma[:] = 0.34
        nb.epsilon[:] = 0.996
        g = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing='ij'), -1).reshape(-1, 3)[:n]
        x = 0.5 * g + 0.1 + 0.01 * rng.rand(n, 3)
        for R in (1, 2, 5):
            eng = _native.Engine(s.flatten(), n_replicas=R, seed=3)
            eng.set_langevin_integrator(120.0, 1.0, 0.002)
            eng.set_positions(x)
            ep, _ = eng.get_energy()
            assert np.all(ep == ep[0]) and np.isfinite(ep[0]) and ep[0] < 0
            eng.velocities_to_temperature(120.0)
            eng.md_run(30)
            xs = [eng.get_positions(r) for r in range(R)]
            assert all(np.all(np.isfinite(v)) for v in xs)
            if R > 1:
                assert np.max(np.abs(xs[0] - xs[1])) > 1e-6        # different noise streams per walker
            eng.close()
