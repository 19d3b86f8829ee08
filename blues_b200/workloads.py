"""Named measurement systems: structure fixture + ``createSystem`` arguments + alchemical region.

Used by ``bench.py`` and by the tests; the data files are the committed fixtures under ``tests/golden`` (the upstream
``eqToluene.prmtop`` is missing, so the T4L entry is the labelled surrogate).
"""
import os

import numpy as np

from . import unit as u
from .structure import Structure
from .alchemy import AbsoluteAlchemicalFactory, AlchemicalRegion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

DEFAULT_FUNCS = {
    'lambda_sterics': 'min(1, (1/0.3)*abs(lambda-0.5))',
    'lambda_electrostatics': 'step(0.2-lambda) - 1/0.2*lambda*step(0.2-lambda) + 1/0.2*(lambda-0.8)*step(lambda-0.8)'}

CASES = {
    'tol_parm': dict(kw=dict(nonbondedMethod='PME', nonbondedCutoff=8.0 * u.angstroms, constraints='HBonds'),
                     alch=list(range(15))),
    'wat_divaline': dict(kw=dict(nonbondedMethod='PME', nonbondedCutoff=10.0 * u.angstroms, constraints='HBonds',
                                 ewaldErrorTolerance=0.005), alch=list(range(16, 35))),
    'vac_divaline': dict(kw=dict(nonbondedMethod='NoCutoff', constraints='HBonds'), alch=list(range(16, 35))),
    't4l_surrogate': dict(kw=dict(nonbondedMethod='PME', nonbondedCutoff=10.0 * u.angstroms, constraints='HBonds',
                                  hydrogenMass=3.024 * u.dalton, ewaldErrorTolerance=0.005),
                          alch=list(range(2634, 2649))),
}

# `t4l` (M2 = BASELINE configs[1]) is the bench line the driver reads; the others are the same engine on the other
# configurations, selected with `bench.py --workload`.
WORKLOADS = {
    't4l': dict(
        case='t4l_surrogate', nsteps_nc=5000, dt=0.004, move='rotate', replicas=1,
        p_in=4672867,         # non-excluded pairs within 1.0 nm at the fixture coordinates (oracle count)
        text='T4L-toluene geometry (22340 atoms, surrogate force field: eqToluene.prmtop is missing upstream), explicit '
             'TIP3P, PME rc 1.0 nm tol 5e-3 grid 24x25x28, HBonds + rigid water, HMR 3.024 Da, dt 4 fs, 300 K, '
             'nstepsNC=5000, RandomLigandRotationMove at moveStep'),
    't4l_frozen': dict(
        case='t4l_surrogate', nsteps_nc=5000, dt=0.004, move='rotate', replicas=1, p_in=None, freeze_radius_angstrom=5.0,
        text='T4L-toluene geometry (22340 atoms, surrogate force field), the example\'s default variant: freeze_radius '
             '5 A around :LIG (275 mobile atoms, 22065 frozen), PME rc 1.0 nm tol 5e-3, HBonds, HMR, dt 4 fs, '
             'nstepsNC=5000, RandomLigandRotationMove at moveStep'),
    't4l_tol5e4': dict(
        case='t4l_surrogate', nsteps_nc=5000, dt=0.004, move='rotate', replicas=1, p_in=4672867,
        kw=dict(ewaldErrorTolerance=0.0005),
        text='T4L-toluene geometry (22340 atoms, surrogate force field), OpenMM\'s default ewaldErrorTolerance 5e-4 '
             '(PME grid 45x48x54), rc 1.0 nm, HBonds, HMR, dt 4 fs, nstepsNC=5000, RandomLigandRotationMove'),
    'tolparm': dict(
        case='tol_parm', nsteps_nc=100, dt=0.002, move='rotate', replicas=1, p_in=None,
        text='M1 / BASELINE configs[0]: toluene in TIP3P (TOL-parm.prmtop, 975 atoms, cubic 2.1786 nm), PME rc 0.8 nm '
             'tol 5e-4 grid 24^3, HBonds, dt 2 fs, 300 K, nstepsNC=100, RandomLigandRotationMove at moveStep'),
    'water': dict(
        case='t4l_surrogate', nsteps_nc=1000, dt=0.002, move='water', replicas=1, p_in=4672867,
        alch=[2657, 2658, 2659], selection='(index 1656) or (index 1657)', radius_nm=0.9,
        text='M4 / BASELINE configs[3]: WaterTranslationMove on the T4L geometry (22340 atoms, surrogate force field), '
             'alchemical water = first HOH (atoms 2657-2659), sphere 0.9 nm around atoms 1656/1657, nstepsNC=1000, '
             'dt 2 fs, swap / translate / check hooks on the device'),
    'm5': dict(
        case='tol_parm', tile=(6, 6, 7), nsteps_nc=5000, dt=0.002, move='rotate', replicas=8, p_in=None,
        kw=dict(cutoff_angstrom=10.0, ewaldErrorTolerance=0.005),
        text='M5 / BASELINE configs[4]: TOL-parm tiled 6x6x7 = 245700 atoms, box 13.07x13.07x15.25 nm, PME rc 1.0 nm '
             'tol 5e-3, HBonds, dt 2 fs, 300 K, nstepsNC=5000, one alchemical toluene, 8 walkers per GPU'),
    'm5_t4l': dict(
        case='t4l_surrogate', tile=(2, 2, 3), nsteps_nc=5000, dt=0.004, move='rotate', replicas=8, p_in=None,
        text='M5 as SURVEY §8(d) states it: T4L geometry tiled 2x2x3 = 268080 atoms, PME rc 1.0 nm tol 5e-3, HBonds, HMR, '
             'dt 4 fs, nstepsNC=5000, one alchemical toluene, 8 walkers per GPU'),
}


def tile_structure(s, reps):
    """Replicate a periodic Structure reps = (nx, ny, nz) times along its box vectors (orthorhombic)."""
    d = s.to_arrays()
    n = len(d['atom_names'])
    nres = len(d['residue_names'])
    nx, ny, nz = reps
    ncopy = nx * ny * nz
    out = {}
    for k in ('atomic_numbers', 'masses', 'charges', 'lj_sigma', 'lj_epsilon', 'atom_names', 'atom_types'):
        out[k] = np.tile(d[k], ncopy)
    out['residue_names'] = np.tile(d['residue_names'], ncopy)
    rp = np.asarray(d['residue_pointers'])
    if len(rp) == nres + 1:
        out['residue_pointers'] = np.concatenate([rp[:-1] + c * n for c in range(ncopy)] + [[ncopy * n]])
    else:
        out['residue_pointers'] = np.concatenate([rp + c * n for c in range(ncopy)])
    out['atom_residue'] = np.concatenate([np.asarray(d['atom_residue']) + c * nres for c in range(ncopy)])
    for idx, extra in (('bonds', ('bond_k', 'bond_r0')), ('angles', ('angle_k', 'angle_t0')),
                       ('dihedrals', ('dihedral_k', 'dihedral_per', 'dihedral_phase', 'dihedral_scee', 'dihedral_scnb',
                                      'dihedral_ignore_end', 'dihedral_improper'))):
        a = np.asarray(d[idx])
        out[idx] = np.concatenate([a + c * n for c in range(ncopy)]) if len(a) else a
        for e in extra:
            out[e] = np.tile(d[e], ncopy)
    box = np.asarray(d['box'], float)
    shifts = [(i, j, k) for i in range(nx) for j in range(ny) for k in range(nz)]
    out['coordinates'] = np.concatenate([d['coordinates'] + np.asarray(sh) * box[:3] for sh in shifts])
    out['box'] = np.concatenate([box[:3] * np.asarray(reps), box[3:]])
    return Structure.from_arrays(out)


def lambda_tables(nsteps, n_H=2, funcs=None):
    """lambda_sterics / lambda_electrostatics at lambda_step / (nsteps * n_H), as the integrator object tabulates them."""
    from . import lepton
    funcs = funcs or DEFAULT_FUNCS
    n = nsteps * n_H
    return (np.asarray(lepton.tabulate(funcs.get('lambda_sterics', '1'), n), float),
            np.asarray(lepton.tabulate(funcs.get('lambda_electrostatics', '1'), n), float))


def load_workload(name='t4l'):
    """Structure, alchemical System, flat topology and start coordinates (nm) of a measurement configuration."""
    w = WORKLOADS[name]
    base = Structure.load_npz(os.path.join(GOLDEN, w['case'] + '.npz'))
    s = tile_structure(base, w['tile']) if w.get('tile') else base
    kw = dict(CASES[w['case']]['kw'])
    over = dict(w.get('kw', {}))
    if 'cutoff_angstrom' in over:
        kw['nonbondedCutoff'] = over.pop('cutoff_angstrom') * u.angstroms
    kw.update(over)
    system = s.createSystem(**kw)
    if w.get('freeze_radius_angstrom'):
        from .simulation import SystemFactory
        system = SystemFactory.freeze_radius(s, system, freeze_distance=w['freeze_radius_angstrom'] * u.angstroms,
                                             freeze_center=':LIG', freeze_solvent=':HOH,NA,CL,Cl-')
    alch = w.get('alch', CASES[w['case']]['alch'])
    system = AbsoluteAlchemicalFactory().create_alchemical_system(system, AlchemicalRegion(alchemical_atoms=alch))
    return dict(w, name=name, structure=s, base_structure=base, system=system, topo=system.flatten(),
                x=s.coordinates * 0.1, alch=alch)
