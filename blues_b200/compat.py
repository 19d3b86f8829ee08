"""Import shim: makes code written against BLUES and its dependency stack resolve to this package.

``blues_b200.compat.install()`` registers module aliases in ``sys.modules`` (only for names that are not importable
already), limited to the surface the reference's NCMC path touches:

    blues, blues.{simulation,moves,integrators,reporters,settings,utils}   -> blues_b200.*
    simtk.unit                                                             -> blues_b200.unit
    simtk.openmm (System, Context, Platform, LangevinIntegrator, the Force classes …), simtk.openmm.app
    parmed (load_file, Structure, amber.AmberMask / Rst7, geometry.center_of_mass)
    openmmtools.alchemy (AbsoluteAlchemicalFactory, AlchemicalRegion)
    mdtraj (load, load_netcdf, compute_distances, compute_dihedrals, utils.uniform_quaternion …) -> blues_b200.trajectory

so that a user script or the reference's own test files (``blues/tests/test_simulation.py`` …) import unchanged.
With ``data_root`` given, ``blues.utils.get_data_filename('blues', 'tests/data/…')`` resolves inside that checkout
(the reference keeps its fixtures under ``blues/tests/data``).  Nothing here adds behaviour: every name is an alias.
"""
import importlib.util
import os
import sys
import types

_INSTALLED = {}


def _missing(name):
    if name in sys.modules:
        return name in _INSTALLED
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__blues_b200_alias__ = True
    return m


def install(data_root=None, force=False):
    """Register the aliases; returns the list of module names that were installed."""
    from . import unit, mm, system, structure, alchemy, utils, simulation, moves, integrators, reporters, settings
    from . import trajectory
    import blues_b200

    mods = {}
    # --- simtk ----------------------------------------------------------------------------------------------------
    openmm_attrs = {k: getattr(system, k) for k in dir(system) if k.endswith('Force') or k in ('System', 'CMMotionRemover',
                                                                                            'MonteCarloBarostat', 'XmlSerializer')}
    for k in ('Context', 'State', 'Platform', 'LangevinIntegrator', 'Vec3', 'OpenMMException'):
        openmm_attrs[k] = getattr(mm, k)
    from . import lepton
    openmm_attrs.update(Discrete1DFunction=lepton.Discrete1DFunction, Continuous1DFunction=lepton.Continuous1DFunction)
    app = _module('simtk.openmm.app', Simulation=mm.Simulation, StateDataReporter=reporters.StateDataReporter,
                  NoCutoff=system.NoCutoff, CutoffNonPeriodic=system.CutoffNonPeriodic,
                  CutoffPeriodic=system.CutoffPeriodic, Ewald=system.Ewald, PME=system.PME,
                  HBonds=system.HBonds, AllBonds=system.AllBonds, HAngles=system.HAngles)
    openmm = _module('simtk.openmm', app=app, unit=unit, **openmm_attrs)
    openmm_inner = _module('simtk.openmm.openmm', **openmm_attrs)       # `from simtk.openmm.openmm import Discrete1DFunction`
    openmm.openmm = openmm_inner
    simtk = _module('simtk', unit=unit, openmm=openmm)
    simtk.__path__ = []
    openmm.__path__ = []
    mods.update({'simtk': simtk, 'simtk.unit': unit, 'simtk.openmm': openmm, 'simtk.openmm.app': app,
                 'simtk.openmm.openmm': openmm_inner})
    # --- parmed -----------------------------------------------------------------------------------------------------
    amber = _module('parmed.amber', AmberMask=structure.AmberMask, Rst7=structure.Rst7)
    geometry = _module('parmed.geometry', center_of_mass=structure.geometry.center_of_mass)
    parmed = _module('parmed', load_file=structure.load_file, Structure=structure.Structure, amber=amber,
                     geometry=geometry, unit=unit)
    parmed.__path__ = []
    mods.update({'parmed': parmed, 'parmed.amber': amber, 'parmed.geometry': geometry})
    # --- openmmtools ------------------------------------------------------------------------------------------------
    omt_alchemy = _module('openmmtools.alchemy', AbsoluteAlchemicalFactory=alchemy.AbsoluteAlchemicalFactory,
                          AlchemicalRegion=alchemy.AlchemicalRegion)
    omt = _module('openmmtools', alchemy=omt_alchemy)
    omt.__path__ = []
    mods.update({'openmmtools': omt, 'openmmtools.alchemy': omt_alchemy})
    # --- mdtraj: reading back the trajectories the reporters write (blues/tests/test_ethylene.py:113-163) -------------
    md_utils = _module('mdtraj.utils', uniform_quaternion=lambda size=None, random_state=None:
                       moves.uniform_quaternion(random_state),
                       rotation_matrix_from_quaternion=moves.rotation_matrix_from_quaternion)
    mdtraj = _module('mdtraj', load=trajectory.load, load_netcdf=trajectory.load_netcdf, Trajectory=trajectory.Trajectory,
                     compute_distances=trajectory.compute_distances, compute_dihedrals=trajectory.compute_dihedrals,
                     utils=md_utils)
    mdtraj.__path__ = []
    mods.update({'mdtraj': mdtraj, 'mdtraj.utils': md_utils})
    # --- blues --------------------------------------------------------------------------------------------------------
    blues_utils = utils
    if data_root is not None:
        # the reference's fixtures live inside its package directory: <checkout>/blues/tests/data
        blues_utils = _module('blues.utils', **{k: v for k, v in vars(utils).items() if not k.startswith('__')})
        root = os.path.abspath(data_root)

        def get_data_filename(package_root, relative_path):
            fn = os.path.join(root, package_root, relative_path)
            if not os.path.exists(fn):
                raise ValueError("Sorry! %s does not exist. If you just added it, you'll have to re-install" % fn)
            return fn
        blues_utils.get_data_filename = get_data_filename
    blues = _module('blues', utils=blues_utils, simulation=simulation, moves=moves, integrators=integrators,
                    reporters=reporters, settings=settings, __version__=getattr(blues_b200, '__version__', '0.1.0'))
    blues.__path__ = []
    mods.update({'blues': blues, 'blues.utils': blues_utils, 'blues.simulation': simulation, 'blues.moves': moves,
                 'blues.integrators': integrators, 'blues.reporters': reporters, 'blues.settings': settings})

    # code written for BLUES configures logging by the reference's module names ("blues.simulation", …)
    import logging
    for short, mod in (('simulation', simulation), ('moves', moves), ('reporters', reporters), ('settings', settings),
                       ('utils', utils), ('integrators', integrators)):
        if isinstance(getattr(mod, 'logger', None), logging.Logger):
            mod.logger = logging.getLogger('blues.' + short)

    done = []
    for name, m in mods.items():
        top = name.split('.')[0]
        if force or _missing(top) or top in _INSTALLED or name in _INSTALLED:
            sys.modules[name] = m
            _INSTALLED[name] = m
            done.append(name)
    return done


def uninstall():
    for name in list(_INSTALLED):
        if sys.modules.get(name) is _INSTALLED[name]:
            del sys.modules[name]
        del _INSTALLED[name]
