"""Utility functions with the signatures of ``blues/utils.py``.

``calculateNCMCSteps`` (``blues/utils.py:89-145``), ``zero_masses`` (``:202-221``), ``atomIndexfromTop``
(``:224-245``), ``parse_unit_quantity`` (``:180-199``), ``check_amber_selection`` (``:148-177``),
``saveSimulationFrame`` (``:20-61``), ``print_host_info`` (``:64-86``), ``get_data_filename`` (``:248-273``).
"""
import logging
import math
import os
import sys
from platform import uname

from . import unit
from .structure import AmberMask, Structure

logger = logging.getLogger(__name__)


def saveSimulationFrame(simulation, outfname):
    """Write the current frame of ``simulation`` (format from the extension: .pdb, .rst7/.inpcrd/.restrt)."""
    state = simulation.context.getState(getPositions=True, getVelocities=True, enforcePeriodicBox=True)
    top = simulation.topology
    base = getattr(top, '_s', None)
    if base is None:
        raise ValueError('simulation.topology does not come from a blues_b200 Structure')
    s = Structure.from_arrays(base.to_arrays())
    s.positions = state.getPositions(asNumpy=True)
    s.velocities = state.getVelocities(asNumpy=True)
    box = state.getPeriodicBoxVectors(asNumpy=True).value_in_unit(unit.angstroms)
    if base.box is not None:
        s.box = [box[0][0], box[1][1], box[2][2], 90.0, 90.0, 90.0]
    s.save(outfname, overwrite=True)
    logger.info('\tSaving Frame to: %s' % outfname)


def print_host_info(simulation):
    """Log platform, host and device properties of the simulation."""
    plat = simulation.context.getPlatform()
    msg = 'blues_b200 simulation generated for {} platform\n'.format(plat.getName())
    for k, v in uname()._asdict().items():
        msg += '{} = {} \n'.format(k, v)
    for prop in plat.getPropertyNames():
        msg += '{} = {} \n'.format(prop, plat.getPropertyValue(simulation.context, prop))
    logger.info(msg)


def calculateNCMCSteps(nstepsNC=0, nprop=1, propLambda=0.3, **kwargs):
    """Number of lambda-switching steps, total propagation steps and the step at which the move is applied.

    Same arithmetic as ``blues/utils.py:89-145``: ``nstepsNC`` is forced even; with extra propagation
    (``nprop`` > 1 inside ``0.5 ± propLambda``) the switching-step count is re-derived so the protocol stays
    symmetric; ``moveStep = nstepsNC / 2``.
    """
    if nstepsNC % 2:
        even = nstepsNC & ~1
        msg = 'nstepsNC=%i must be even for symmetric protocol.' % nstepsNC
        if not even:
            logger.error(msg)
            sys.exit(1)
        logger.warning(msg + ' Setting to nstepsNC=%i' % even)
        nstepsNC = even
    lambda_steps = int(nstepsNC / (2 * (nprop * propLambda + 0.5 - propLambda)))
    if lambda_steps % 2:
        lambda_steps += 1
    inside = int(nprop * (2 * math.floor(propLambda * lambda_steps)))
    outside = int(2 * math.ceil((0.5 - propLambda) * lambda_steps))
    prop_steps = inside + outside
    if prop_steps != nstepsNC:
        logger.warning('nstepsNC=%s is incompatible with prop_lambda=%s and nprop=%s.' % (nstepsNC, propLambda, nprop))
        logger.warning('Changing NCMC protocol to %s lambda switching within %s total propagation steps.' %
                       (lambda_steps, prop_steps))
        nstepsNC = lambda_steps
    return {'nstepsNC': nstepsNC, 'propSteps': prop_steps, 'moveStep': int(nstepsNC / 2), 'nprop': nprop,
            'propLambda': propLambda}


def check_amber_selection(structure, selection):
    """Exit with an error if the Amber mask selects nothing (``blues/utils.py:148-177``)."""
    try:
        mask = AmberMask(structure, str(selection))
        idx = [i for i in mask.Selected()]
    except Exception:
        idx = []
    if not idx:
        if ':' in selection:
            names = sorted(set(structure.residue_names))
        else:
            names = sorted(set(structure.atom_names))
        logger.error("'%s' was not a valid Amber selection. \n\tValid names: %s" % (selection, names))
        sys.exit(1)
    return True


def parse_unit_quantity(unit_quantity_str):
    """``'3.024*daltons'`` / ``'1 * 1/picoseconds'`` → Quantity (``blues/utils.py:180-199``)."""
    value, _, uname_ = unit_quantity_str.replace(' ', '').partition('*')
    if '/' in uname_:
        num, den = uname_.split('/', 1)
        top = unit.dimensionless if num in ('1', '1.0') else getattr(unit, num)
        return unit.Quantity(float(value), top / getattr(unit, den))
    return unit.Quantity(float(value), getattr(unit, uname_))


def zero_masses(system, atomList=None):
    """Freeze atoms by zeroing their masses (``blues/utils.py:202-221``)."""
    for index in atomList:
        system.setParticleMass(int(index), 0 * unit.daltons)
    return system


def atomIndexfromTop(resname, topology):
    """Atom indices of every residue whose name is ``resname`` (``blues/utils.py:224-245``)."""
    return [atom.index for atom in topology.atoms() if str(resname) == str(atom.residue.name)]


def get_data_filename(package_root, relative_path):
    """Path of a data file shipped with ``package_root`` (``blues/utils.py:248-273``)."""
    import importlib
    try:
        root = os.path.dirname(importlib.import_module(package_root).__file__)
    except ImportError:
        root = package_root
    fn = os.path.join(root, relative_path)
    if not os.path.exists(fn):
        alt = os.path.join(os.path.dirname(root), relative_path)
        if os.path.exists(alt):
            return alt
        raise ValueError("Sorry! %s does not exist. If you just added it, you'll have to re-install" % fn)
    return fn


def spreadLambdaProtocol(switching_values, steps, switching_types='auto', kind='cubic', return_tab_function=True):
    """Stretch a one-way lambda schedule (1 -> 0, e.g. the windows of a free-energy protocol) over a symmetric NCMC
    protocol of ``steps`` steps: off at the midpoint, back on at the end (``blues/utils.py:276-369``).

    The schedule is mirrored about its last point, interpolated (``scipy.interpolate.interp1d(kind=kind)``) on the
    ``steps + 1`` points k / steps, and the plateaus the interpolant may overshoot are restored: a ``'sterics'`` schedule
    stays at exactly 1 before its last leading 1 and after the mirrored one, an ``'electrostatics'`` schedule at exactly 0
    between its first 0 and the mirrored one (``'auto'``: sterics if the second value is still 1).  The second half is the
    mirror image of the first.  Returns a ``Discrete1DFunction`` for ``integrator.addTabulatedFunction`` (use it as
    ``'name(lambda*steps)'`` in ``alchemical_functions``), or the plain list with ``return_tab_function=False``.
    """
    import numpy as np
    from scipy.interpolate import interp1d
    values = [float(v) for v in switching_values]
    n_ones = values.count(1.0)
    first_zero = values.index(0.0)
    both = values + values[-2::-1]                         # off state in the middle
    x = np.arange(len(both)) / float(len(both) - 1)
    xs = np.arange(0.0, 1.0 + 1.0 / float(steps), 1.0 / float(steps))
    ys = interp1d(x, both, kind=kind)(xs)
    if switching_types == 'auto':
        switching_types = 'sterics' if both[1] == 1.0 else 'electrostatics'
    if switching_types == 'sterics':
        lo, hi = x[n_ones - 1], x[-n_ones]
        tab = [1.0 if (t < lo or t > hi) else float(y) for t, y in zip(xs, ys)]
    elif switching_types == 'electrostatics':
        lo, hi = x[first_zero], x[-(first_zero + 1)]
        tab = [0.0 if (lo < t < hi) else float(y) for t, y in zip(xs, ys)]
    else:
        raise ValueError('`switching_types` should be either sterics or electrostatics, currently ' + str(switching_types))
    half = math.floor(len(tab) / 2.0)
    tab = [v if i <= half else tab[-i - 1] for i, v in enumerate(tab)]
    for i, v in enumerate(tab):
        if v < 0.0 or v > 1.0:
            raise ValueError('interpolated lambda %f at index %i is outside [0, 1]: check switching_types / kind' % (v, i))
    if return_tab_function:
        from .lepton import Discrete1DFunction
        return Discrete1DFunction(tab)
    return tab
