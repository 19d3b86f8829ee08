"""Move plugin API and the two built-in moves (``blues/moves.py:39-410, 846-1083``).

``Move`` (``initializeSystem`` / ``beforeMove`` / ``move`` / ``afterMove`` / ``_error``), ``MoveEngine``
(``selectMove`` / ``runEngine``), ``RandomLigandRotationMove`` and ``WaterTranslationMove`` keep the reference's
names, arguments and context-in / context-out contract, so user ``Move`` subclasses written for BLUES keep
working (they cost one host round-trip at ``moveStep``).  A move may additionally implement
``device_move()`` returning a descriptor executed inside ``bl_ncmc_run`` with no host round-trip;
``RandomLigandRotationMove`` does whenever the caller has not pinned a numpy random state.

``SmartDartMove`` (centre-of-mass darting) and ``CombinationMove`` are host-path moves.  Not provided:
``SideChainMove`` (OpenEye-licensed, marked untested upstream ``blues/moves.py:413-415``; outside the NCMC hot
path — DESIGN.md).
"""
import copy
import re
import sys
import traceback

import numpy

from . import unit
from . import _native
from .structure import geometry


class Move(object):
    """Base class: hooks called by ``BLUESSimulation._stepNCMC`` around the NCMC protocol."""

    def __init__(self):
        pass

    def initializeSystem(self, system, integrator):
        """Modify the alchemical system / NCMC integrator once at set-up; returns both."""
        return system, integrator

    def beforeMove(self, context):
        """Called before the first NCMC step of an iteration."""
        return context

    def afterMove(self, context):
        """Called after the last NCMC step of an iteration."""
        return context

    def _error(self, context):
        """Called when a step of the NCMC protocol raised."""
        return context

    def move(self, context):
        """Perturb the coordinates held by ``context`` at the protocol midpoint; returns the context."""
        return context


def _random_state(random_state):
    """mdtraj / scikit-learn ``check_random_state``: None → global numpy RNG, int → fresh RandomState(seed)."""
    if random_state is None or random_state is numpy.random:
        return numpy.random.mtrand._rand
    if isinstance(random_state, (int, numpy.integer)):
        return numpy.random.RandomState(random_state)
    return random_state


def uniform_quaternion(random_state=None):
    """Haar-uniform unit quaternion (Shoemake, Graphics Gems III) — what ``mdtraj.utils.uniform_quaternion`` draws."""
    u0, u1, u2 = _random_state(random_state).uniform(0, 1, size=3)
    s1, s2 = numpy.sqrt(1 - u0), numpy.sqrt(u0)
    return numpy.array([s1 * numpy.sin(2 * numpy.pi * u1), s1 * numpy.cos(2 * numpy.pi * u1),
                        s2 * numpy.sin(2 * numpy.pi * u2), s2 * numpy.cos(2 * numpy.pi * u2)])


def rotation_matrix_from_quaternion(q):
    """3x3 rotation matrix of the unit quaternion (w, x, y, z)."""
    w, x, y, z = q
    return numpy.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class RandomLigandRotationMove(Move):
    """Random rigid rotation of a ligand about its centre of mass (``blues/moves.py:148-310``).

    Parameters
    ----------
    structure : Structure of the whole system
    resname : residue name (substring match, as in the reference) of the ligand
    random_state : None (device RNG / global numpy RNG), int seed or ``numpy.random.RandomState``
    """

    def __init__(self, structure, resname='LIG', random_state=None):
        self.structure = structure
        self.resname = resname
        self.random_state = random_state
        self.atom_indices = self.getAtomIndices(structure, self.resname)
        sub = structure[self.atom_indices]
        self.topology = sub.topology
        self.totalmass = 0
        self.masses = []
        self.center_of_mass = None
        self.positions = sub.positions
        self._calculateProperties()

    def getAtomIndices(self, structure, resname):
        return [atom.index for atom in structure.topology.atoms() if str(resname) in atom.residue.name]

    def getMasses(self, topology):
        """Element masses (float32 column, dalton) and their sum — not the (possibly repartitioned) system masses."""
        n = int(topology.getNumAtoms())
        masses = unit.Quantity(numpy.zeros([n, 1], numpy.float32), unit.dalton)
        for idx, atom in enumerate(topology.atoms()):
            masses[idx] = atom.element._mass
        return masses, masses.sum()

    def getCenterOfMass(self, positions, masses):
        coordinates = numpy.asarray(positions._value, numpy.float32)
        return geometry.center_of_mass(coordinates, masses) * positions.unit

    def _calculateProperties(self):
        self.masses, self.totalmass = self.getMasses(self.topology)
        self.center_of_mass = self.getCenterOfMass(self.positions, self.masses)

    def device_move(self):
        """Descriptor for the on-device rotation, or None when numpy-RNG semantics were requested / overridden."""
        if self.random_state is not None or type(self).move is not RandomLigandRotationMove.move:
            return None
        return dict(kind=_native.BL_MOVE_ROTATE, atoms=list(self.atom_indices),
                    masses=numpy.asarray(self.masses._value, float).reshape(-1))

    def _after_device_move(self, context):
        positions = context.getState(getPositions=True).getPositions(asNumpy=True)
        self.positions = positions[self.atom_indices]
        self.center_of_mass = self.getCenterOfMass(self.positions, self.masses)

    def move(self, context):
        """Host path (one state round-trip): x' = (x − c)·R + c with R from a uniform quaternion."""
        positions = context.getState(getPositions=True).getPositions(asNumpy=True)
        self.positions = positions[self.atom_indices]
        self.center_of_mass = self.getCenterOfMass(self.positions, self.masses)
        reduced = self.positions - self.center_of_mass
        rot = rotation_matrix_from_quaternion(uniform_quaternion(self.random_state))
        moved = numpy.dot(reduced._value, rot) * positions.unit + self.center_of_mass
        for k, atomidx in enumerate(self.atom_indices):
            positions[atomidx] = moved[k]
        context.setPositions(positions)
        self.positions = context.getState(getPositions=True).getPositions(asNumpy=True)[self.atom_indices]
        return context


class MoveEngine(object):
    """Chooses among moves with given probabilities and runs the chosen one (``blues/moves.py:313-410``)."""

    def __init__(self, moves, probabilities=None):
        self.moves = moves if isinstance(moves, list) else [moves]
        if probabilities is None:
            self.probabilities = [1.0 / len(self.moves)] * len(self.moves)
        else:
            total = float(sum(probabilities))
            self.probabilities = [p / total for p in probabilities]
        if len(self.moves) != len(self.probabilities):
            print('moves and probability list lengths need to match')
            raise IndexError
        self.selected_move = None
        self.move_name = None

    def selectMove(self):
        k = numpy.random.choice(len(self.probabilities), p=self.probabilities)
        self.selected_move = self.moves[k]
        self.move_name = self.selected_move.__class__.__name__

    def runEngine(self, context):
        try:
            return self.selected_move.move(context)
        except Exception as e:
            print('Error: move not implemented correctly, printing traceback:')
            traceback.print_tb(sys.exc_info()[2])
            print(e)
            raise SystemExit


class SmartDartMove(RandomLigandRotationMove):
    """Centre-of-mass smart darting between pre-defined ligand positions (``blues/moves.py:1086-1514``; Andricioaei,
    Straub and Voter, J. Chem. Phys. 114, 6994 (2001)).

    Every file of ``coord_files`` (with ``topology`` when the files carry coordinates only) contributes one dart:
    the ligand's centre of mass expressed in the local frame of three ``basis_particles`` (origin p1, axes p2 − p1,
    p3 − p1 and their cross product), so that darts follow the protein.  ``move`` rebuilds the darts from the
    current basis particles; if the ligand's centre of mass lies within ``dart_radius`` of exactly one dart, the
    ligand is translated to another dart chosen uniformly (the same one allowed with ``self_dart``), keeping its
    offset from the dart centre.  Overlapping darts raise, as upstream.  Marked untested upstream; host path only
    (one state round-trip at ``moveStep``).
    """

    def __init__(self, structure, basis_particles, coord_files, topology=None, dart_radius=0.2 * unit.nanometers,
                 self_dart=False, resname='LIG'):
        super(SmartDartMove, self).__init__(structure, resname=resname)
        if len(coord_files) < 2:
            raise ValueError('You should include at least two files in coord_files ' +
                             'in order to benefit from smart darting')
        self.dartboard = []
        self.n_dartboard = []
        self.particle_pairs = []
        self.particle_weights = []
        self.basis_particles = list(basis_particles)
        self.dart_radius = dart_radius
        self.self_dart = self_dart
        self.dartsFromParmEd(coord_files, topology)

    def device_move(self):
        return None

    # ---- frame algebra (plain arrays in nanometers) ---------------------------------------------------------
    @staticmethod
    def _nm(v):
        return numpy.asarray(v.value_in_unit(unit.nanometers) if unit.is_quantity(v) else v, float)

    def _normalize(self, vector):
        v = numpy.asarray(vector, float)
        return v / numpy.sqrt(numpy.sum(v * v))

    def _localCoord(self, particle1, particle2, particle3):
        p1, p2, p3 = self._nm(particle1), self._nm(particle2), self._nm(particle3)
        v1, v2 = p2 - p1, p3 - p1
        return v1, v2, numpy.cross(v1, v2)

    def _changeBasis(self, a, b):
        """Coordinates of the vector ``b`` in the basis whose vectors are the rows of ``a``."""
        return numpy.linalg.solve(numpy.asarray(a, float).T, numpy.asarray(b, float))

    def _undoBasis(self, a, b):
        """Cartesian vector with coordinates ``b`` in the basis whose vectors are the rows of ``a``."""
        return numpy.dot(numpy.asarray(a, float).T, numpy.asarray(b, float))

    def _findNewCoord(self, particle1, particle2, particle3, center):
        basis = numpy.array(self._localCoord(particle1, particle2, particle3))
        return self._changeBasis(basis, self._nm(center) - self._nm(particle1)) * unit.nanometers

    def _findOldCoord(self, particle1, particle2, particle3, center):
        basis = numpy.array(self._localCoord(particle1, particle2, particle3))
        return (self._undoBasis(basis, self._nm(center)) + self._nm(particle1)) * unit.nanometers

    # ---- darts -------------------------------------------------------------------------------------------------
    def dartsFromParmEd(self, coord_files, topology=None):
        from .structure import load_file
        n_dartboard, dartboard = [], []
        for coord_file in coord_files:
            if not isinstance(coord_file, str):
                temp = coord_file                                 # an already loaded Structure
            elif topology is None:
                temp = load_file(coord_file)
            else:
                temp = load_file(topology, xyz=coord_file)
            pos = self._nm(temp.positions)
            lig = pos[self.atom_indices] * unit.nanometers
            p = pos[self.basis_particles]
            com = self.getCenterOfMass(lig, self.masses)
            new_coord = self._findNewCoord(p[0], p[1], p[2], com)
            old_coord = self._findOldCoord(p[0], p[1], p[2], new_coord)
            numpy.testing.assert_almost_equal(self._nm(old_coord), self._nm(com).reshape(3), decimal=1)
            n_dartboard.append(new_coord)
            dartboard.append(old_coord)
        self.n_dartboard = n_dartboard
        self.dartboard = dartboard

    def _findDart(self, context):
        pos = self._nm(context.getState(getPositions=True).getPositions(asNumpy=True))
        p = pos[self.basis_particles]
        self.dartboard = [self._findOldCoord(p[0], p[1], p[2], dart) for dart in self.n_dartboard]
        return self.dartboard[:]

    def _calc_from_center(self, com):
        c = self._nm(com).reshape(3)
        radius = self._nm(self.dart_radius)
        diffs = [c - self._nm(dart).reshape(3) for dart in self.dartboard]
        inside = [k for k, d in enumerate(diffs) if numpy.sqrt(numpy.sum(d * d)) <= radius]
        if len(inside) == 1:
            return inside[0], diffs[inside[0]] * unit.nanometers
        if len(inside) == 0:
            return None, diffs[-1] * unit.nanometers
        raise ValueError(' The spheres defining two darting regions have overlapped, ' +
                         'which results in potential problems with detailed balance. ' +
                         'We are terminating the simulation. Please check the size and ' +
                         'identity of your darting regions defined by dart_radius.')

    def _reDart(self, selected_dart, changevec):
        dartindex = list(range(len(self.dartboard)))
        if self.self_dart is False:
            dartindex.pop(selected_dart)
        choice = numpy.random.choice(dartindex)
        return (self._nm(self.dartboard[choice]).reshape(3) + self._nm(changevec).reshape(3)) * unit.nanometers

    def move(self, context):
        if len(self.n_dartboard) == 0:
            raise ValueError('No darts are specified. Make sure you use ' +
                             'SmartDartMove.dartsFromParmed() before using the move() function')
        positions = context.getState(getPositions=True).getPositions(asNumpy=True)
        xyz = self._nm(positions)
        self._findDart(context)
        center = self.getCenterOfMass(xyz[self.atom_indices] * unit.nanometers, self.masses)
        selected_dart, changevec = self._calc_from_center(com=center)
        if selected_dart is not None:
            shift = self._nm(self._reDart(selected_dart, changevec)) - self._nm(center).reshape(3)
            new = xyz.copy()
            new[self.atom_indices] = new[self.atom_indices] + shift
            context.setPositions(new * unit.nanometers)
        return context                     # (upstream returns None when no dart is hit; the Move contract wants the context)


class SideChainMove(Move):
    """Placeholder for ``blues/moves.py:413-843``: upstream needs the OpenEye toolkits (a licensed dependency) for its
    rotor perception and prints "SideChainMove class will be unavailable" without them; here the class exists so that
    ``from blues.moves import SideChainMove`` resolves, and constructing it says why it cannot run."""

    def __init__(self, *args, **kwargs):
        raise ImportError('SideChainMove needs the OpenEye toolkits (openeye.oechem), which are not available; '
                          'it is outside the NCMC hot path this package covers (DESIGN.md)')


class CombinationMove(Move):
    """Several moves applied as one, in listed or in reverse order with equal probability (detailed balance) —
    ``blues/moves.py:1517-1560``.  The upstream class is marked untested and cannot run as written (it reads
    ``self.move_list`` and calls an undefined ``reverse``, and returns nothing); this one keeps its constructor and
    intent.  ``atom_indices`` is the union of the members' (``SimulationFactory`` builds the alchemical region from
    the first move's ``atom_indices``); the ``beforeMove`` / ``afterMove`` / ``_error`` hooks fan out in the same order.
    """

    def __init__(self, moves):
        self.moves = list(moves)
        self.move_list = self.moves
        seen = []
        for m in self.moves:
            for a in getattr(m, 'atom_indices', []):
                if a not in seen:
                    seen.append(a)
        self.atom_indices = seen

    def initializeSystem(self, system, integrator):
        for m in self.moves:
            system, integrator = m.initializeSystem(system, integrator)
        return system, integrator

    def beforeMove(self, context):
        for m in self.moves:
            context = m.beforeMove(context)
        return context

    def afterMove(self, context):
        for m in self.moves:
            context = m.afterMove(context)
        return context

    def _error(self, context):
        for m in self.moves:
            context = m._error(context)
        return context

    def move(self, context):
        order = self.moves if numpy.random.random() > 0.5 else list(reversed(self.moves))
        for single_move in order:
            context = single_move.move(context)
        return context


# ---------------------------------------------------------------------------------------------------------
# atom-selection mini language (the mdtraj DSL forms the reference passes: 'protein',
# '(index 1656) or (index 1657)', 'resname LIG', 'name CA', 'resid 10 to 20')
# ---------------------------------------------------------------------------------------------------------
_PROTEIN = {'ALA', 'ARG', 'ASN', 'ASP', 'ASH', 'CYS', 'CYX', 'CYM', 'GLN', 'GLU', 'GLH', 'GLY', 'HIS', 'HID', 'HIE',
            'HIP', 'ILE', 'LEU', 'LYS', 'LYN', 'MET', 'PHE', 'PRO', 'SER', 'THR', 'TRP', 'TYR', 'VAL', 'ACE', 'NME',
            'NHE'}
_WATER = {'HOH', 'WAT', 'TIP3', 'TIP4', 'SPC', 'H2O'}


def select_atoms(structure, expression):
    s = structure
    toks = re.findall(r'\(|\)|[^\s()]+', expression)
    pos = [0]

    def peek():
        return toks[pos[0]] if pos[0] < len(toks) else None

    def take():
        pos[0] += 1
        return toks[pos[0] - 1]

    def numbers():
        vals = []
        while peek() is not None and re.fullmatch(r'-?\d+', peek()):
            a = int(take())
            if peek() == 'to':
                take()
                vals.extend(range(a, int(take()) + 1))
            else:
                vals.append(a)
        return vals

    def words():
        vals = []
        while peek() is not None and peek() not in ('and', 'or', ')', 'not'):
            vals.append(take())
        return vals

    def primary():
        t = take()
        n = s.n_atoms
        if t == '(':
            v = expr()
            if take() != ')':
                raise ValueError('unbalanced parentheses in %r' % expression)
            return v
        if t == 'not':
            return ~primary()
        if t == 'all':
            return numpy.ones(n, bool)
        if t == 'protein':
            return numpy.asarray([s.residue_names[r] in _PROTEIN for r in s.atom_residue], bool)
        if t in ('water', 'waters'):
            return numpy.asarray([s.residue_names[r] in _WATER for r in s.atom_residue], bool)
        if t == 'index':
            m = numpy.zeros(n, bool)
            m[numbers()] = True
            return m
        if t in ('resid', 'residue', 'resSeq'):
            return numpy.isin(s.atom_residue, numbers())
        if t == 'resname':
            w = set(words())
            return numpy.asarray([s.residue_names[r] in w for r in s.atom_residue], bool)
        if t == 'name':
            w = set(words())
            return numpy.asarray([nm in w for nm in s.atom_names], bool)
        if t in ('element', 'symbol'):
            from .structure import _SYMBOLS
            w = set(words())
            return numpy.asarray([_SYMBOLS[z] in w for z in s.atomic_numbers], bool)
        raise ValueError('unsupported selection keyword %r in %r' % (t, expression))

    def conj():
        v = primary()
        while peek() == 'and':
            take()
            v = v & primary()
        return v

    def expr():
        v = conj()
        while peek() == 'or':
            take()
            v = v | conj()
        return v

    mask = expr()
    if pos[0] != len(toks):
        raise ValueError('could not parse selection %r' % expression)
    return numpy.nonzero(mask)[0]


class WaterTranslationMove(Move):
    """Swap a random water inside a sphere around the protein selection's centre of mass with the alchemical
    water, translate it to a uniform random point of the sphere at the protocol midpoint and force rejection
    (``protocol_work = 999999``) if it ends outside (``blues/moves.py:846-1083``).

    ``on_device=True`` (default) runs the three hooks as kernels behind ``bl_apply_move`` / ``bl_ncmc_run``
    (``BL_MOVE_WATER_SWAP`` / ``_TRANSLATE`` / ``_CHECK``): no state round-trip, one independent choice per walker,
    random numbers from the engine's Philox stream.  ``on_device=False`` is the host path with the reference's use
    of the global numpy RNG (one full-state round-trip per hook)."""

    def __init__(self, structure, water_name=['WAT', 'HOH'], protein_selection='protein', radius=2.3 * unit.nanometers,
                 on_device=True):
        self.on_device = on_device
        self.radius = radius
        self.water_name = water_name
        self.water_residues = []
        self.before_ncmc_check = True
        self.structure = structure
        for res in structure.topology.residues():
            if res.name in self.water_name:
                self.water_residues.append([atom.index for atom in res.atoms()])
        self.atom_indices = self.water_residues[0]
        self.protein_atoms = select_atoms(structure, protein_selection)
        self.protein_masses = self._getMasses(structure.topology)[self.protein_atoms]
        self.go = True

    def _random_sphere_point(self, radius, origin):
        r = radius * (numpy.random.random() ** (1. / 3.))
        phi = numpy.random.uniform(0, 2 * numpy.pi)
        costheta = numpy.random.uniform(-1, 1)
        theta = numpy.arccos(costheta)
        direction = numpy.array([numpy.sin(theta) * numpy.cos(phi), numpy.sin(theta) * numpy.sin(phi), numpy.cos(theta)])
        return direction * r + origin

    def _getMasses(self, topology):
        masses = unit.Quantity(numpy.zeros([int(topology.getNumAtoms()), 1], numpy.float32), unit.dalton)
        for idx, atom in enumerate(topology.atoms()):
            masses[idx] = atom.element._mass
        return masses

    def _getCenterOfMass(self, positions, masses):
        if unit.is_quantity(positions):
            xyz = numpy.asarray(positions._value, numpy.float32)
            return geometry.center_of_mass(xyz, masses) * positions.unit
        return geometry.center_of_mass(numpy.asarray(positions, numpy.float32), masses)

    # periodic distance between atom `index` and a point (nm), float32 like mdtraj
    def _distance(self, xyz_nm, box_nm, index, point_nm):
        d = numpy.asarray(xyz_nm[index], numpy.float32) - numpy.asarray(point_nm, numpy.float32)
        L = numpy.asarray(box_nm, numpy.float32)
        d = d - L * numpy.round(d / L)
        return float(numpy.sqrt(numpy.sum(d * d)))

    def _frame(self, context):
        state = context.getState(getPositions=True, getVelocities=True)
        pos = state.getPositions(asNumpy=True)
        vel = state.getVelocities(asNumpy=True)
        box = numpy.diag(state.getPeriodicBoxVectors(asNumpy=True).value_in_unit(unit.nanometers))
        xyz = pos.value_in_unit(unit.nanometers)
        com = self._getCenterOfMass(numpy.asarray(xyz, numpy.float32)[self.protein_atoms], self.protein_masses)
        return pos, vel, xyz, box, numpy.asarray(com, float)

    # ---- device path ------------------------------------------------------------------------------------
    def _hooks_on_device(self):
        # the three hooks share per-walker device state (sphere centre, go flag): all of them run there or none does
        return (self.on_device and type(self).move is WaterTranslationMove.move
                and type(self).beforeMove is WaterTranslationMove.beforeMove
                and type(self).afterMove is WaterTranslationMove.afterMove)

    def _device(self, context):
        return self._hooks_on_device() and hasattr(context, '_engine')

    def _descriptor(self, kind, with_waters=False):
        d = dict(kind=kind, step=0, atoms=list(self.atom_indices),
                 center_atoms=numpy.asarray(self.protein_atoms, numpy.int32),
                 center_masses=numpy.asarray(self.protein_masses._value, float).reshape(-1),
                 radius=self.radius.value_in_unit(unit.nanometers))
        if with_waters:
            n = len(self.atom_indices)
            d['waters'] = numpy.asarray([w for w in self.water_residues if len(w) == n], numpy.int32)
        return d

    def device_move(self):
        """Descriptor of the on-device translation for ``bl_ncmc_run`` (None on the host path)."""
        if not self._hooks_on_device():
            return None
        return self._descriptor(_native.BL_MOVE_WATER_TRANSLATE)

    def _apply(self, context, kind, with_waters=False):
        d = self._descriptor(kind, with_waters)
        d.pop('step')
        context._engine.apply_move(d.pop('kind'), d.pop('atoms'), None, **d)
        return context

    # ---- hooks ---------------------------------------------------------------------------------------------
    def beforeMove(self, context):
        if self._device(context):
            self.go = True                     # the per-walker flag lives on the device
            return self._apply(context, _native.BL_MOVE_WATER_SWAP, with_waters=True)
        pos, vel, xyz, box, com = self._frame(context)
        self._com = com
        radius = self.radius.value_in_unit(unit.nanometers)
        waters = copy.deepcopy(self.water_residues)
        numpy.random.shuffle(waters)
        chosen = None
        for w in waters:
            if self._distance(xyz, box, w[0], com) <= radius:
                chosen = w
                break
        if chosen is None:
            self.go = False
            return context
        new_pos = numpy.copy(pos._value)
        new_vel = numpy.copy(vel._value)
        new_pos[self.atom_indices], new_pos[chosen] = pos._value[chosen], pos._value[self.atom_indices]
        new_vel[self.atom_indices], new_vel[chosen] = vel._value[chosen], vel._value[self.atom_indices]
        context.setPositions(new_pos * pos.unit)
        context.setVelocities(new_vel * vel.unit)
        self.go = True
        return context

    def move(self, context):
        if self._device(context):
            return self._apply(context, _native.BL_MOVE_WATER_TRANSLATE)
        if self.go is False:
            return context
        pos, vel, xyz, box, _ = self._frame(context)
        # the reference reuses the centre of mass computed from the frame cached by beforeMove (moves.py:1021)
        com = getattr(self, '_com', None)
        if com is None:
            com = self._frame(context)[4]
        radius = self.radius.value_in_unit(unit.nanometers)
        target = self._random_sphere_point(radius, com)
        if self._distance(xyz, box, self.atom_indices[0], com) >= radius:
            return context
        new_pos = numpy.copy(xyz)
        displacement = new_pos[self.atom_indices[0]] - target
        new_pos[self.atom_indices] = new_pos[self.atom_indices] - displacement
        context.setPositions(new_pos * unit.nanometers)
        return context

    def afterMove(self, context):
        if self._device(context):
            return self._apply(context, _native.BL_MOVE_WATER_CHECK)
        pos, vel, xyz, box, com = self._frame(context)
        if self._distance(xyz, box, self.atom_indices[0], com) > self.radius.value_in_unit(unit.nanometers) and self.go:
            context._integrator.setGlobalVariableByName("protocol_work", 999999)
        return context
