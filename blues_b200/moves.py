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
    """Rotation of a side chain about one of its rotatable heavy-atom bonds (``blues/moves.py:418-843``).

    Upstream needs the OpenEye toolkits for three graph queries — backbone atoms (``OEIsBackboneAtom``), ring membership
    (``OEFindRingAtomsAndBonds``) and ``bond.IsRotor()`` — and is unavailable without a licence.  The same queries are
    answered here from the structure's own bond graph: backbone = atoms named N, CA, C, O; a rotor is a bond between two
    heavy atoms that is not in a ring, whose ends both carry another heavy neighbour (terminal groups such as methyls,
    hydroxyls, carbonyl oxygens do not define a torsion) and that does not join two three-coordinate C/N atoms (amide,
    guanidinium and other conjugated bonds, which OpenEye perceives as non-single from the geometry).  Everything else —
    the dictionaries ``rot_atoms`` / ``rot_bonds`` / ``qry_atoms``, the atom walk of ``getRotAtoms``, the random choice of
    bond and angle with Python's ``random`` module, the rotation matrix and ``move`` — follows the reference, whose test
    (``blues/tests/test_sidechain.py:64-68``: one rotor, 11 listed atoms for valine) this class reproduces.

    Bonds are identified by ``(index of first atom, index of second atom)`` tuples instead of OpenEye bond pointers.
    Like upstream, ``atom_indices`` is the ``rot_atoms`` dictionary itself (``blues/moves.py:473``).
    """
    BACKBONE_NAMES = ('N', 'CA', 'C', 'O')

    def __init__(self, structure, residue_list, verbose=False, write_move=False):
        self.structure = structure
        self.molecule = self._bondGraph()
        self.residue_list = residue_list
        self.all_atoms = [atom.index for atom in self.structure.topology.atoms()]
        self.rot_atoms, self.rot_bonds, self.qry_atoms = self.getRotBondAtoms()
        self.atom_indices = self.rot_atoms
        self.verbose = verbose
        self.write_move = write_move

    # -- graph perception (the OpenEye part of the reference) ------------------------------------------------------
    def _bondGraph(self):
        """Adjacency lists, atomic numbers, residue numbers / names and ring-bond flags of the structure."""
        s = self.structure
        n = len(s.atoms)
        adj = [[] for _ in range(n)]
        for i, j in numpy.asarray(s.bonds, int).reshape(-1, 2):
            adj[int(i)].append(int(j))
            adj[int(j)].append(int(i))
        z = [int(a.atomic_number) for a in s.atoms]
        resnum, resname = [0] * n, [''] * n
        for res in s.topology.residues():
            for a in res.atoms():
                resnum[a.index] = int(res.id)
                resname[a.index] = res.name
        names = [a.name for a in s.atoms]
        return {'adj': adj, 'z': z, 'resnum': resnum, 'resname': resname, 'names': names}

    def _bondInRing(self, a, b):
        """True if a path from a to b exists that does not use the bond a-b (depth-first search)."""
        adj = self.molecule['adj']
        seen, stack = {a}, [x for x in adj[a] if x != b]
        while stack:
            x = stack.pop()
            if x == b:
                return True
            if x not in seen:
                seen.add(x)
                stack.extend(y for y in adj[x] if y not in seen)
        return False

    def _isRotor(self, a, b):
        m = self.molecule
        if m['z'][a] <= 1 or m['z'][b] <= 1:
            return False
        heavy = lambda x: sum(1 for y in m['adj'][x] if m['z'][y] > 1)
        if heavy(a) < 2 or heavy(b) < 2:
            return False
        planar = lambda x: m['z'][x] in (6, 7) and len(m['adj'][x]) == 3
        if planar(a) and planar(b):
            return False
        return not self._bondInRing(a, b)

    def getBackboneAtoms(self, molecule):
        """Indices of the backbone atoms (``blues/moves.py:485-508``)."""
        return [i for i, nm in enumerate(molecule['names']) if nm in self.BACKBONE_NAMES]

    def getTargetAtoms(self, molecule, backbone_atoms, residue_list):
        """Non-backbone atoms of the target residues as ``{atom index: atom index}`` (``blues/moves.py:510-550``)."""
        bb = set(backbone_atoms)
        qry_atoms = {}
        for i in range(len(molecule['names'])):
            if i not in bb and molecule['resnum'][i] in residue_list and molecule['resname'][i] != 'HOH':
                qry_atoms[i] = i
        return qry_atoms, backbone_atoms

    def findHeavyRotBonds(self, pdb_OEMol, qry_atoms):
        """``{(a, b): residue number}`` of the rotatable heavy-atom bonds of the query atoms (``blues/moves.py:552-586``)."""
        rot_bonds = {}
        for atom in qry_atoms.keys():
            for nb in pdb_OEMol['adj'][atom]:
                bond = (min(atom, nb), max(atom, nb))
                if bond not in rot_bonds and self._isRotor(atom, nb):
                    rot_bonds[bond] = pdb_OEMol['resnum'][atom]
        return rot_bonds

    def getRotAtoms(self, rotbonds, molecule, backbone_atoms):
        """``{residue: {bond: [axis1, axis2, atoms downstream of the bond]}}`` — the walk of ``blues/moves.py:588-651``."""
        backbone = set(backbone_atoms)
        adj, z = molecule['adj'], molecule['z']
        rot_atom_dict = {}
        for bond, resnum in rotbonds.items():
            ax1, ax2 = bond
            rot_atom_dict.setdefault(resnum, {})[bond] = []
            idx_list = [ax1, ax2]
            query_list = []
            if ax1 not in backbone:
                query_list.append(ax1)
            if ax2 not in query_list and ax2 not in backbone:
                query_list.append(ax2)
            for atom in query_list:                       # grows while it is walked, as upstream
                for candidate in adj[atom]:
                    if candidate not in query_list and candidate not in backbone and candidate != ax2:
                        query_list.append(candidate)
                        if z[candidate] > 1:
                            for can_nbor in adj[candidate]:
                                if can_nbor not in query_list and candidate not in backbone and candidate != ax2:
                                    query_list.append(can_nbor)
            for y in query_list:
                if y not in idx_list:
                    idx_list.append(y)
            rot_atom_dict[resnum][bond] = list(idx_list)
        return rot_atom_dict

    def getRotBondAtoms(self):
        backbone_atoms = self.getBackboneAtoms(self.molecule)
        qry_atoms, backbone_atoms = self.getTargetAtoms(self.molecule, backbone_atoms, self.residue_list)
        rot_bonds = self.findHeavyRotBonds(self.molecule, qry_atoms)
        rot_atoms = self.getRotAtoms(rot_bonds, self.molecule, backbone_atoms)
        return rot_atoms, rot_bonds, qry_atoms

    # -- the move ----------------------------------------------------------------------------------------------------
    def chooseBondandTheta(self):
        """Random residue, bond and angle in [0, 2 pi) from Python's ``random`` module (``blues/moves.py:683-708``)."""
        import math
        import random
        res_choice = random.choice(list(self.rot_atoms.keys()))
        bond_choice = random.choice(list(self.rot_atoms[res_choice].keys()))
        targetatoms = self.rot_atoms[res_choice][bond_choice]
        theta_ran = random.random() * 2 * math.pi
        return theta_ran, targetatoms, res_choice, bond_choice

    def rotation_matrix(self, axis, theta):
        """Counter-clockwise rotation about ``axis`` by ``theta`` radians (``blues/moves.py:710-730``)."""
        import math
        axis = numpy.asarray(axis, float)
        axis = axis / math.sqrt(numpy.dot(axis, axis))
        a = math.cos(theta / 2.0)
        b, c, d = -axis * math.sin(theta / 2.0)
        aa, bb, cc, dd = a * a, b * b, c * c, d * d
        bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
        return numpy.array([[aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)],
                            [2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)],
                            [2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]])

    def move(self, context, verbose=False):
        """Rotate the atoms listed for a randomly chosen bond about it (``blues/moves.py:732-843``): the two axis atoms
        stay where they are, the structure's coordinates follow the context's."""
        theta, target_atoms, res, bond = self.chooseBondandTheta()
        print('Rotating bond: %s in resnum: %s by %.2f radians' % (bond, res, theta))
        state = context.getState(getPositions=True)
        nc_positions = copy.deepcopy(state.getPositions(asNumpy=True))
        x = numpy.array(nc_positions.value_in_unit(unit.nanometers), float)
        axis1, axis2 = target_atoms[0], target_atoms[1]
        rot_matrix = self.rotation_matrix(x[axis1] - x[axis2], theta)
        for atom in target_atoms:
            before = x[atom].copy()
            x[atom] = numpy.dot(rot_matrix, x[atom] - x[axis2]) + x[axis2]
            if self.verbose or verbose:
                print('atom %d: %s -> %s' % (atom, before, x[atom]))
        context.setPositions(x * unit.nanometers)
        self.structure.positions = x * unit.nanometers
        if self.write_move:
            self.structure.save('sc_move_%s_%s_%s.pdb' % (res, axis1, axis2), overwrite=True)
        return context


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
