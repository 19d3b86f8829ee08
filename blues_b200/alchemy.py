"""Alchemical system construction (the ``openmmtools.alchemy`` surface used at ``blues/simulation.py:300-317``).

``AbsoluteAlchemicalFactory.create_alchemical_system`` returns a copy of the MD ``System`` in which
(settings of ``generateAlchSystem`` ``blues/simulation.py:221-317``):

* alchemical atoms get q = 0, ε = 0 in the ``NonbondedForce`` and every exception touching them gets
  chargeProd = 0, ε = 0 — so they have no reciprocal-space interaction at any λ
  (``alchemical_pme_treatment='direct-space'``);
* softcore sterics (``lambda_sterics``) and direct-space electrostatics (``lambda_electrostatics``) are
  re-added for alchemical×environment and alchemical×alchemical pairs, plus the modified 1-4 exceptions.

The added interactions are recorded as ``Custom*Force`` parameter records (for API parity with the
reference's tests, ``tests/test_simulation.py:170-186``) and, for the engine, in ``system.alchemical``.
"""
import copy
import numpy as np

from .system import CustomBondForce, CustomNonbondedForce, NonbondedForce

_STERICS = ('U_sterics = (lambda_sterics^softcore_a)*4*epsilon*x*(x-1.0); x = (sigma/reff_sterics)^6;'
            'reff_sterics = sigma*((softcore_alpha*(1.0-lambda_sterics)^softcore_b + (r/sigma)^softcore_c))^(1/softcore_c)')
_ELEC = ('U_electrostatics = (lambda_electrostatics^softcore_d)*ONE_4PI_EPS0*chargeprod*erfc(alpha_ewald*reff)/reff;'
         'reff = sigma*((softcore_beta*(1.0-lambda_electrostatics)^softcore_e + (r/sigma)^softcore_f))^(1/softcore_f)')


class AlchemicalRegion(object):
    def __init__(self, alchemical_atoms=None, annihilate_electrostatics=True, annihilate_sterics=False,
                 softcore_alpha=0.5, softcore_a=1, softcore_b=1, softcore_c=6, softcore_beta=0.0,
                 softcore_d=1, softcore_e=1, softcore_f=2, **kwargs):
        self.alchemical_atoms = sorted(int(a) for a in (alchemical_atoms or []))
        self.annihilate_electrostatics = bool(annihilate_electrostatics)
        self.annihilate_sterics = bool(annihilate_sterics)
        self.softcore = dict(softcore_alpha=float(softcore_alpha), softcore_a=float(softcore_a),
                             softcore_b=float(softcore_b), softcore_c=float(softcore_c),
                             softcore_beta=float(softcore_beta), softcore_d=float(softcore_d),
                             softcore_e=float(softcore_e), softcore_f=float(softcore_f))


class AbsoluteAlchemicalFactory(object):
    def __init__(self, consistent_exceptions=False, switch_width=None, alchemical_pme_treatment='direct-space',
                 alchemical_rf_treatment='switched', disable_alchemical_dispersion_correction=False, **kwargs):
        if alchemical_pme_treatment != 'direct-space':
            raise NotImplementedError("only alchemical_pme_treatment='direct-space' (the BLUES setting) is supported")
        self.disable_alchemical_dispersion_correction = disable_alchemical_dispersion_correction

    def create_alchemical_system(self, reference_system, alchemical_regions):
        region = alchemical_regions
        system = copy.deepcopy(reference_system)
        nb = system._force(NonbondedForce)
        if nb is None:
            # openmmtools copies every force it has no alchemical rule for: a system whose lambda dependence lives in
            # its own Custom*Force global parameters (blues/tests/test_ethylene.py:76) comes back unchanged
            return system
        atoms = np.asarray(region.alchemical_atoms, np.int32)
        if len(atoms) == 0:
            raise ValueError('alchemical region is empty')
        if region.softcore['softcore_beta'] != 0.0:
            raise NotImplementedError('softcore_beta != 0 is not supported by the native kernels')
        is_alch = np.zeros(system.getNumParticles(), bool)
        is_alch[atoms] = True
        al = dict(region.softcore)
        al['atoms'] = atoms
        al['charge'] = nb.charge[atoms].copy()
        al['sigma'] = nb.sigma[atoms].copy()
        al['epsilon'] = nb.epsilon[atoms].copy()
        al['annihilate_sterics'] = region.annihilate_sterics
        al['annihilate_electrostatics'] = region.annihilate_electrostatics
        # exceptions that touch the region move out of the NonbondedForce
        touch = is_alch[nb.exc_idx[:, 0]] | is_alch[nb.exc_idx[:, 1]]
        live = touch & ((nb.exc_qq != 0) | (nb.exc_eps != 0))
        al['exc_pairs'] = nb.exc_idx[live].copy()
        al['exc_qq'] = nb.exc_qq[live].copy()
        al['exc_sigma'] = nb.exc_sigma[live].copy()
        al['exc_eps'] = nb.exc_eps[live].copy()
        nb.exc_qq[touch] = 0.0
        nb.exc_eps[touch] = 0.0
        nb.charge[atoms] = 0.0
        nb.epsilon[atoms] = 0.0
        system.alchemical = al
        # parameter records, one per interaction class openmmtools would add
        system.addForce(CustomNonbondedForce(_STERICS, 'alchemically modified NonbondedForce for non-alchemical/alchemical sterics'))
        system.addForce(CustomNonbondedForce(_STERICS, 'alchemically modified NonbondedForce for alchemical/alchemical sterics'))
        system.addForce(CustomNonbondedForce(_ELEC, 'alchemically modified NonbondedForce for non-alchemical/alchemical electrostatics'))
        system.addForce(CustomNonbondedForce(_ELEC, 'alchemically modified NonbondedForce for alchemical/alchemical electrostatics'))
        if len(al['exc_pairs']):
            system.addForce(CustomBondForce(_STERICS, 'alchemically modified BondForce for alchemical/alchemical sterics exceptions'))
            system.addForce(CustomBondForce('U = lambda_electrostatics*ONE_4PI_EPS0*chargeprod/r',
                                            'alchemically modified BondForce for alchemical/alchemical electrostatics exceptions'))
        return system
