"""``SystemFactory`` / ``SimulationFactory`` / ``BLUESSimulation`` / ``MonteCarloSimulation`` on the native engine.

Drop-in counterparts of ``blues/simulation.py`` (``SystemFactory`` :31-480, ``SimulationFactory`` :483-809,
``BLUESSimulation`` :812-1257, ``MonteCarloSimulation`` :1260-1335): same class and method names, argument
meaning, state-table bookkeeping and error behaviour.  What differs is below the API: the three Simulations
are ``blues_b200.mm.Simulation`` objects that own native handles, the NCMC leg runs as device-resident chunks
(``bl_ncmc_run``) instead of a per-step Python loop, and built-in moves execute on the device.
"""
import logging
import math
import sys

import numpy as np

from . import mm as openmm
from . import unit, utils
from . import alchemy
from .integrators import AlchemicalExternalLangevinIntegrator
from .structure import AmberMask
from .system import CustomExternalForce, MonteCarloBarostat

logger = logging.getLogger(__name__)
finfo = np.finfo(np.float32)
rtol = finfo.precision


class SystemFactory(object):
    """Generates the MD ``System`` and the alchemical ``System`` for a structure (``blues/simulation.py:31-86``).

    >>> systems = SystemFactory(structure, ligand.atom_indices, config['system'])
    >>> systems.md, systems.alch
    """

    def __init__(self, structure, atom_indices, config=None):
        self.structure = structure
        self.atom_indices = atom_indices
        self._config = config
        if self._config:
            self.alch_config = self._config.pop('alchemical') if 'alchemical' in self._config else {}
            self.md = SystemFactory.generateSystem(self.structure, **self._config)
            self.alch = SystemFactory.generateAlchSystem(self.md, self.atom_indices, **self.alch_config)

    @staticmethod
    def amber_selection_to_atomidx(structure, selection):
        """Amber-mask string → list of atom indices (``blues/simulation.py:88-111``)."""
        return [i for i in AmberMask(structure, str(selection)).Selected()]

    @staticmethod
    def atomidx_to_atomlist(structure, mask_idx):
        """Atom objects for the given indices (``blues/simulation.py:113-137``)."""
        wanted = set(int(i) for i in mask_idx)
        atom_list = [structure.atoms[i] for i in sorted(wanted)]
        logger.debug('\nFreezing {}'.format(atom_list))
        return atom_list

    @classmethod
    def generateSystem(cls, structure, **kwargs):
        """``structure.createSystem(**kwargs)`` (``blues/simulation.py:139-219``)."""
        return structure.createSystem(**kwargs)

    @classmethod
    def generateAlchSystem(cls, system, atom_indices, softcore_alpha=0.5, softcore_a=1, softcore_b=1, softcore_c=6,
                           softcore_beta=0.0, softcore_d=1, softcore_e=1, softcore_f=2,
                           annihilate_electrostatics=True, annihilate_sterics=False,
                           disable_alchemical_dispersion_correction=True, alchemical_pme_treatment='direct-space',
                           suppress_warnings=True, **kwargs):
        """Alchemical system for the NCMC simulation (``blues/simulation.py:221-317``)."""
        factory = alchemy.AbsoluteAlchemicalFactory(
            disable_alchemical_dispersion_correction=disable_alchemical_dispersion_correction,
            alchemical_pme_treatment=alchemical_pme_treatment)
        region = alchemy.AlchemicalRegion(
            alchemical_atoms=atom_indices, softcore_alpha=softcore_alpha, softcore_a=softcore_a,
            softcore_b=softcore_b, softcore_c=softcore_c, softcore_beta=softcore_beta, softcore_d=softcore_d,
            softcore_e=softcore_e, softcore_f=softcore_f, annihilate_electrostatics=annihilate_electrostatics,
            annihilate_sterics=annihilate_sterics)
        return factory.create_alchemical_system(system, region)

    @classmethod
    def restrain_positions(cls, structure, system, selection="(@CA,C,N)", weight=5.0, **kwargs):
        """Harmonic positional restraints on the selection (``blues/simulation.py:319-362``)."""
        mask_idx = cls.amber_selection_to_atomidx(structure, selection)
        logger.info("{} positional restraints applied to selection: '{}' ({} atoms) on {}".format(
            weight, selection, len(mask_idx), system))
        force = CustomExternalForce('k_restr*periodicdistance(x, y, z, x0, y0, z0)^2')
        force.addGlobalParameter("k_restr", weight)
        for name in ('x0', 'y0', 'z0'):
            force.addPerParticleParameter(name)
        xyz_nm = np.asarray(structure.coordinates, float) * 0.1
        for i in mask_idx:
            force.addParticle(i, xyz_nm[i])
        system.addForce(force)
        return system

    @classmethod
    def freeze_atoms(cls, structure, system, freeze_selection=":LIG", **kwargs):
        """Zero the masses of the selection (``blues/simulation.py:364-392``)."""
        mask_idx = cls.amber_selection_to_atomidx(structure, freeze_selection)
        logger.info("Freezing selection '{}' ({} atoms) on {}".format(freeze_selection, len(mask_idx), system))
        cls.atomidx_to_atomlist(structure, mask_idx)
        return utils.zero_masses(system, mask_idx)

    @classmethod
    def freeze_radius(cls, structure, system, freeze_distance=5.0 * unit.angstrom, freeze_center=':LIG',
                      freeze_solvent=':HOH,NA,CL', **kwargs):
        """Freeze everything except non-solvent residues within ``freeze_distance`` of ``freeze_center``
        (``blues/simulation.py:394-480``), with the reference's sanity exits."""
        n_atoms = system.getNumParticles()
        if hasattr(freeze_distance, '_value'):
            freeze_distance = freeze_distance._value
        selection = "(%s<:%f)&!(%s)" % (freeze_center, freeze_distance, freeze_solvent)
        logger.info('Inverting parmed selection for freezing: %s' % selection)
        site_idx = cls.amber_selection_to_atomidx(structure, selection)
        freeze_idx = set(range(n_atoms)) - set(site_idx)
        center_idx = cls.amber_selection_to_atomidx(structure, freeze_center)
        if len(freeze_idx) == n_atoms:
            logger.error('All %i atoms appear to be selected for freezing. Check your atom selection.' % len(freeze_idx))
            sys.exit(1)
        if len(freeze_idx) / n_atoms == 0.98:
            logger.error('98% of your system appears to be selected for freezing. Check your atom selection')
            sys.exit(1)
        if len(site_idx) <= len(center_idx):
            logger.error("%i unfrozen atoms is less than (or equal to) the number of atoms used as the selection "
                         "center '%s' (%i atoms). Check your atom selection." % (len(site_idx), freeze_center, len(center_idx)))
            sys.exit(1)
        if len(freeze_idx) / n_atoms == 0.80:
            logger.warning('80% of your system appears to be selected for freezing. This may cause unexpected behaviors.')
            sys.exit(1)
        logger.info("Freezing {} atoms {} Angstroms from '{}' on {}".format(len(freeze_idx), freeze_distance,
                                                                           freeze_center, system))
        cls.atomidx_to_atomlist(structure, freeze_idx)
        return utils.zero_masses(system, freeze_idx)


class SimulationFactory(object):
    """Builds the three Simulations (md / alch / ncmc) BLUES needs (``blues/simulation.py:483-600``)."""

    def __init__(self, systems, move_engine, config=None, md_reporters=None, ncmc_reporters=None):
        self._structure, self._system, self._alch_system = systems.structure, systems.md, systems.alch
        self._move_engine = move_engine
        self._atom_indices = move_engine.moves[0].atom_indices
        self.config = config
        if config:
            try:
                self.generateSimulationSet()
            except Exception as err:
                logger.exception(err)
                raise
        for which, reporters in (('md', md_reporters), ('ncmc', ncmc_reporters)):
            if reporters:
                setattr(self, '_%s_reporters' % which, reporters)
                setattr(self, which, self.attachReporters(getattr(self, which), reporters))

    @classmethod
    def addBarostat(cls, system, temperature=300 * unit.kelvin, pressure=1 * unit.atmospheres, frequency=25, **kwargs):
        """Attach a ``MonteCarloBarostat`` (``blues/simulation.py:602-626``).  A ``Context`` built from this system
        with a ``LangevinIntegrator`` (the MD leg) attempts a volume move every ``frequency`` steps
        (``blues_b200/barostat.py``); the NCMC context never does, as upstream."""
        logger.info('Adding MonteCarloBarostat with {}. MD simulation will be {} NPT.'.format(pressure, temperature))
        system.addForce(MonteCarloBarostat(pressure, temperature, frequency))
        return system

    @classmethod
    def generateIntegrator(cls, temperature=300 * unit.kelvin, dt=0.002 * unit.picoseconds, friction=1, seed=None,
                           **kwargs):
        """Langevin integrator of the MD / alch Simulations (``blues/simulation.py:628-648``).  ``seed`` (an addition:
        YAML ``simulation: seed:``) fixes the Philox key; without it every Context draws its own, like OpenMM."""
        integrator = openmm.LangevinIntegrator(temperature, friction, dt)
        if seed:
            integrator.setRandomNumberSeed(int(seed))
        return integrator

    @classmethod
    def generateNCMCIntegrator(cls, nstepsNC=None, alchemical_functions={
            'lambda_sterics': 'min(1, (1/0.3)*abs(lambda-0.5))',
            'lambda_electrostatics': 'step(0.2-lambda) - 1/0.2*lambda*step(0.2-lambda) + 1/0.2*(lambda-0.8)*step(lambda-0.8)'},
            splitting="H V R O R V H", temperature=300 * unit.kelvin, dt=0.002 * unit.picoseconds, nprop=1,
            propLambda=0.3, seed=None, **kwargs):
        """NCMC integrator with the reference's defaults (``blues/simulation.py:650-705``); note that, as in the
        reference, ``friction`` is not forwarded (collision rate stays 1/ps)."""
        integrator = AlchemicalExternalLangevinIntegrator(alchemical_functions=alchemical_functions, splitting=splitting,
                                                          temperature=temperature, nsteps_neq=nstepsNC, timestep=dt,
                                                          nprop=nprop, prop_lambda=propLambda)
        if seed:
            integrator.setRandomNumberSeed(int(seed) + 2)      # md, alch and ncmc contexts use distinct keys
        return integrator

    @classmethod
    def generateSimFromStruct(cls, structure, system, integrator, platform=None, properties={}, **kwargs):
        """Simulation with box, positions and Maxwell–Boltzmann velocities set (``blues/simulation.py:707-745``)."""
        # `nReplicas` (an addition: YAML `simulation: nReplicas:`) is the number of independent walkers of the JOB; under
        # torchrun each rank (one per GPU) holds its round-robin share (parallel.shard_walkers) on device LOCAL_RANK, or on
        # `devices[LOCAL_RANK]` when a device list is given.  Walkers never communicate.
        n_replicas = kwargs.get('nReplicas', None)
        if n_replicas:
            from . import parallel
            rank, world = parallel.rank_world()
            n_replicas = len(parallel.shard_walkers(int(n_replicas), rank, world))
            if n_replicas == 0:
                raise ValueError('nReplicas (%s) is smaller than the number of ranks (%d)' % (kwargs.get('nReplicas'), world))
            devices = kwargs.get('devices', None)
            local_rank = int(__import__('os').environ.get('LOCAL_RANK', '0'))
            properties = dict(properties)
            if devices:
                devices = devices if isinstance(devices, (list, tuple)) else [devices]
                properties.setdefault('DeviceIndex', int(devices[local_rank % len(devices)]))
            elif world > 1:
                properties.setdefault('DeviceIndex', local_rank)
            if platform is None:
                platform = 'CUDA'
        if platform is None:
            simulation = openmm.Simulation(structure.topology, system, integrator, n_replicas=n_replicas)
        else:
            plat = openmm.Platform.getPlatformByName(platform)
            props = {str(k): str(v) for k, v in properties.items()}
            simulation = openmm.Simulation(structure.topology, system, integrator, plat, props, n_replicas=n_replicas)
        if structure.box_vectors:
            simulation.context.setPeriodicBoxVectors(*structure.box_vectors)
        simulation.context.setPositions(structure.positions)
        simulation.context.setVelocitiesToTemperature(integrator.getTemperature())
        return simulation

    @staticmethod
    def attachReporters(simulation, reporter_list):
        """Append reporters to the Simulation (``blues/simulation.py:747-766``)."""
        for rep in reporter_list:
            simulation.reporters.append(rep)
        return simulation

    def generateSimulationSet(self, config=None):
        """md, alch (MD system, energy only) and ncmc Simulations (``blues/simulation.py:768-809``)."""
        cfg = config or self.config
        if cfg.get('seed'):
            # one Philox key per rank (walkers of a rank are subsequences of its key): ranks must not share noise
            from . import parallel
            rank, world = parallel.rank_world()
            if world > 1 and not cfg.get('_seed_is_per_rank'):
                cfg['seed'] = parallel.walker_seed(cfg['seed'], rank)
                cfg['_seed_is_per_rank'] = True
        if 'pressure' in cfg:
            self._system = self.addBarostat(self._system, **cfg)
            logger.warning('NCMC simulation will NOT have pressure control. NCMC will use pressure from last MD state.')
        else:
            logger.info('MD simulation will be {} NVT.'.format(cfg['temperature']))
        # the MD leg and the energy-only copy used by the alchemical correction share the MD system
        self.integrator = self.generateIntegrator(**cfg)
        alch_integrator = self.generateIntegrator(**cfg)
        if cfg.get('seed'):
            alch_integrator.setRandomNumberSeed(int(cfg['seed']) + 1)
        self.md, self.alch = (self.generateSimFromStruct(self._structure, self._system, integ, **cfg)
                              for integ in (self.integrator, alch_integrator))
        if 'moveStep' not in cfg:
            logger.warning('Did not find `moveStep` in configuration. Checking NCMC paramters')
            cfg.update(utils.calculateNCMCSteps(**cfg))
            self.config = cfg
        self.ncmc_integrator = self.generateNCMCIntegrator(**cfg)
        for move in self._move_engine.moves:          # a move may edit the alchemical system / integrator once
            self._alch_system, self.ncmc_integrator = move.initializeSystem(self._alch_system, self.ncmc_integrator)
        self.ncmc = self.generateSimFromStruct(self._structure, self._alch_system, self.ncmc_integrator, **cfg)
        utils.print_host_info(self.ncmc)


class BLUESSimulation(object):
    """NCMC + MD iteration driver (``blues/simulation.py:812-881``).

    >>> blues = BLUESSimulation(simulations)
    >>> blues.run()
    """

    def __init__(self, simulations, config=None):
        self._move_engine = simulations._move_engine
        self._md_sim, self._alch_sim, self._ncmc_sim = simulations.md, simulations.alch, simulations.ncmc
        self._config = config or getattr(simulations, 'config', None)
        if self._config:
            self._printSimulationTiming()
        self.accept = self.reject = 0
        self.acceptRatio = 0
        self.currentIter = 0
        self.stateTable = {sim: {'state0': {}, 'state1': {}} for sim in ('md', 'ncmc')}
        self._integrator_keys_ = ['lambda', 'shadow_work', 'protocol_work', 'Eold', 'Enew']
        self._state_keys = dict(getPositions=True, getVelocities=True, getForces=False, getEnergy=True,
                                getParameters=True, enforcePeriodicBox=True)

    # -- state plumbing -----------------------------------------------------------------------------------
    @classmethod
    def getStateFromContext(cls, context, state_keys):
        """positions / velocities / energies / box of a context as a dict (``blues/simulation.py:883-911``)."""
        state = context.getState(**state_keys)
        return {'positions': state.getPositions(asNumpy=True), 'velocities': state.getVelocities(asNumpy=True),
                'potential_energy': state.getPotentialEnergy(), 'kinetic_energy': state.getKineticEnergy(),
                'box_vectors': state.getPeriodicBoxVectors()}

    @classmethod
    def getIntegratorInfo(cls, ncmc_integrator, integrator_keys=['lambda', 'shadow_work', 'protocol_work', 'Eold', 'Enew']):
        """Work values and energies from the NCMC integrator (``blues/simulation.py:913-936``)."""
        return {key: ncmc_integrator.getGlobalVariableByName(key) for key in integrator_keys}

    @classmethod
    def setContextFromState(cls, context, state, box=True, positions=True, velocities=True):
        """Copy box / positions / velocities of a state dict into a context (``blues/simulation.py:938-963``)."""
        if box:
            context.setPeriodicBoxVectors(*state['box_vectors'])
        if positions:
            context.setPositions(state['positions'])
        if velocities:
            context.setVelocities(state['velocities'])
        return context

    def _printSimulationTiming(self):
        """Log the simulated time and force-evaluation budget (``blues/simulation.py:965-1011``)."""
        cfg = self._config
        dt = cfg['dt'].value_in_unit(unit.picoseconds)
        nIter, nprop, propLambda = cfg['nIter'], cfg['nprop'], cfg['propLambda']
        propSteps, nstepsNC, nstepsMD = cfg['propSteps'], cfg['nstepsNC'], cfg['nstepsMD']
        t_nc, t_md = propSteps * dt, nstepsMD * dt
        msg = 'Total BLUES Simulation Time = %s ps (%s ps/Iter)\n' % ((t_nc + t_md) * nIter, t_nc + t_md)
        msg += 'Total Force Evaluations = %s \n' % (nIter * (propSteps + nstepsMD))
        msg += 'Total NCMC time = %s ps (%s ps/iter)\n' % (t_nc * nIter, t_nc)
        if propSteps != nstepsNC:
            lo, hi = self._ncmc_sim.context._integrator._prop_lambda
            inside = int(nprop * (2 * math.floor(propLambda * nstepsNC)))
            outside = int(2 * math.ceil((0.5 - propLambda) * nstepsNC))
            msg += '\t%s lambda switching steps within %s total propagation steps.\n' % (nstepsNC, propSteps)
            msg += '\tExtra propgation steps between lambda [%s, %s]\n' % (lo, hi)
            msg += '\tLambda: 0.0 -> %s = %s propagation steps\n' % (lo, int(outside / 2))
            msg += '\tLambda: %s -> %s = %s propagation steps\n' % (lo, hi, inside)
            msg += '\tLambda: %s -> 1.0 = %s propagation steps\n' % (hi, int(outside / 2))
        msg += 'Total MD time = %s ps (%s ps/iter)\n' % (t_md * nIter, t_md)
        if 'md_trajectory_interval' in cfg.keys():
            frames = nstepsMD / cfg['md_trajectory_interval']
            msg += 'Trajectory Interval = %s ps/frame (%s frames/iter)' % ((t_nc + t_md) / frames, frames)
        logger.info(msg)

    def _setStateTable(self, simkey, stateidx, stateinfo):
        self.stateTable[simkey][stateidx] = stateinfo

    def _syncStatesMDtoNCMC(self):
        """MD state → NCMC context (``blues/simulation.py:1028-1037``)."""
        md_state0 = self.getStateFromContext(self._md_sim.context, self._state_keys)
        self._setStateTable('md', 'state0', md_state0)
        self._ncmc_sim.context = self.setContextFromState(self._ncmc_sim.context, md_state0)

    # -- NCMC leg ------------------------------------------------------------------------------------------
    def _stepNCMC(self, nstepsNC, moveStep, move_engine=None):
        """Advance the NCMC protocol with the move applied at ``moveStep`` (``blues/simulation.py:1039-1098``).

        Same observable sequence as the reference's per-step loop — ``beforeMove`` before step 0, ``move`` before
        step ``moveStep``, ``afterMove`` after the last step, any exception logged, ``_error`` called and the
        protocol abandoned — but executed as device-resident chunks: moves that provide ``device_move()`` run on
        the GPU inside ``bl_ncmc_run``; other moves cost one host round-trip at ``moveStep``.
        """
        logger.info('Advancing %i NCMC switching steps...' % (nstepsNC))
        ncmc_state0 = self.getStateFromContext(self._ncmc_sim.context, self._state_keys)
        self._setStateTable('ncmc', 'state0', ncmc_state0)
        if not move_engine:
            move_engine = self._move_engine
        self._ncmc_sim.currentIter = self.currentIter
        move_engine.selectMove()
        move = move_engine.selected_move
        nstepsNC, moveStep = int(nstepsNC), int(moveStep)
        integrator = self._ncmc_sim.integrator
        try:
            self._ncmc_sim.context = move.beforeMove(self._ncmc_sim.context)
            device_move = move.device_move() if hasattr(move, 'device_move') else None
            if hasattr(logger, 'report'):
                logger.info = logger.report
            if device_move is not None and 0 <= moveStep < nstepsNC:
                logger.info('Performing %s...' % move_engine.move_name)
                self._run_with_device_move(nstepsNC, moveStep, device_move)
                if hasattr(move, '_after_device_move'):
                    move._after_device_move(self._ncmc_sim.context)
            else:
                if moveStep > 0:
                    self._ncmc_sim.step(min(moveStep, nstepsNC))
                if moveStep < nstepsNC:
                    logger.info('Performing %s...' % move_engine.move_name)
                    self._ncmc_sim.context = move_engine.runEngine(self._ncmc_sim.context)
                    self._ncmc_sim.step(nstepsNC - max(moveStep, 0))
            self._ncmc_sim.context = move.afterMove(self._ncmc_sim.context)
        except Exception as e:
            import traceback
            traceback.print_tb(e.__traceback__)
            logger.error(e)
            move._error(self._ncmc_sim.context)
        finally:
            integrator._scheduled_move = None
        ncmc_state1 = self.getStateFromContext(self._ncmc_sim.context, self._state_keys)
        self._setStateTable('ncmc', 'state1', ncmc_state1)

    def _run_with_device_move(self, nstepsNC, moveStep, device_move):
        """Step through reporter boundaries; the chunk containing ``moveStep`` carries the on-device move."""
        sim = self._ncmc_sim
        integrator = sim.integrator
        done = 0
        start = sim.currentStep
        while done < nstepsNC:
            chunk = nstepsNC - done
            for rep in sim.reporters:
                r = rep.describeNextReport(sim)
                if 0 < r[0] < chunk:
                    chunk = r[0]
            limit = openmm.chunk_limit(sim.reporters, sim)
            if limit is not None and limit < chunk:
                chunk = limit
            if done <= moveStep < done + chunk:
                m = dict(device_move)
                m['step'] = moveStep - done
                integrator._scheduled_move = m
            else:
                integrator._scheduled_move = None
            sim.step(chunk)
            done = sim.currentStep - start
        integrator._scheduled_move = None

    def _computeAlchemicalCorrection(self):
        """−(E_ncmc0 − E_md0 + E_md(x1) − E_ncmc1)/kT (``blues/simulation.py:1100-1119``)."""
        md_state0_PE = self.stateTable['md']['state0']['potential_energy']
        ncmc_state0_PE = self.stateTable['ncmc']['state0']['potential_energy']
        ncmc_state1 = self.stateTable['ncmc']['state1']
        ncmc_state1_PE = ncmc_state1['potential_energy']
        self._alch_sim.context = self.setContextFromState(self._alch_sim.context, ncmc_state1, velocities=False)
        alch_PE = self._alch_sim.context.getState(getEnergy=True).getPotentialEnergy()
        return (ncmc_state0_PE - md_state0_PE + alch_PE - ncmc_state1_PE) * (-1.0 / self._ncmc_sim.context._integrator.kT)

    def _acceptRejectMove(self, write_move=False):
        """Metropolis test on the protocol work plus the alchemical correction (``blues/simulation.py:1121-1166``).
        NaN work skips the correction and can never pass the test; on acceptance the MD context takes the NCMC end
        positions (not the velocities); on rejection the MD potential energy must still equal its value before NCMC."""
        ncmc_context = self._ncmc_sim.context
        log_p = ncmc_context._integrator.getLogAcceptanceProbability(ncmc_context)
        log_u = math.log(np.random.random())
        if not np.isnan(log_p):
            correction = self._computeAlchemicalCorrection()
            logger.debug('NCMCLogAcceptanceProbability = %.6f + Alchemical Correction = %.6f' % (log_p, correction))
            log_p += correction
        if log_p > log_u:
            self.accept += 1
            logger.info('NCMC MOVE ACCEPTED: work_ncmc {} > randnum {}'.format(log_p, log_u))
            self._md_sim.context = self.setContextFromState(self._md_sim.context, self.stateTable['ncmc']['state1'],
                                                            velocities=False)
            if write_move:
                utils.saveSimulationFrame(self._md_sim, '{}acc-it{}.pdb'.format(self._config['outfname'], self.currentIter))
            return
        self.reject += 1
        logger.info('NCMC MOVE REJECTED: work_ncmc {} < {}'.format(log_p, log_u))
        before = self.stateTable['md']['state0']['potential_energy']
        now = self._md_sim.context.getState(getEnergy=True).getPotentialEnergy()
        if not math.isclose(before._value, now._value, rel_tol=10.0 ** -rtol):
            logger.error('Last MD potential energy %s != Current MD potential energy %s. Potential energy should '
                         'match the prior state.' % (before, now))
            sys.exit(1)

    def _resetSimulations(self, temperature=None):
        """Reset the NCMC integrator, redraw MD velocities (``blues/simulation.py:1168-1187``)."""
        if not temperature:
            temperature = self._md_sim.context._integrator.getTemperature()
        self._ncmc_sim.currentStep = 0
        self._ncmc_sim.context._integrator.reset()
        self._md_sim.context.setVelocitiesToTemperature(temperature)

    def _stepMD(self, nstepsMD):
        """Advance the MD simulation (``blues/simulation.py:1189-1213``); failure writes a PDB and exits."""
        logger.info('Advancing %i MD steps...' % (nstepsMD))
        sim = self._md_sim
        sim.currentIter = self.currentIter
        try:
            sim.step(int(nstepsMD))
        except Exception as err:
            start = self.stateTable['md']['state0']
            logger.error(err, exc_info=True)
            for key, label in (('potential_energy', 'potential'), ('kinetic_energy', 'kinetic')):
                logger.error('%s energy before NCMC: %s' % (label, start[key]))
            try:
                utils.saveSimulationFrame(sim, 'MD-fail-it%s-md%i.pdb' % (self.currentIter, sim.currentStep))
            except Exception:
                pass
            sys.exit(1)

    # -- many walkers ----------------------------------------------------------------------------------------
    def _iterateWalkers(self, nstepsNC, moveStep, nstepsMD, temperature):
        """One BLUES iteration of every walker held by this rank (``simulation: nReplicas``), the steps of
        ``blues/simulation.py:1028-1213`` per walker with the state resident on the device: MD -> NCMC copy
        (``bl_copy_state``), NCMC protocol with the move on the device, alchemical correction from per-walker energies of
        the three contexts, Metropolis test (``bl_accept_reject``), accepted walkers' positions NCMC -> MD
        (``bl_copy_state_masked``), reset + velocity redraw, MD.  Returns the per-walker record of the iteration."""
        md, alch, ncmc = (s.context._engine for s in (self._md_sim, self._alch_sim, self._ncmc_sim))
        integrator = self._ncmc_sim.integrator
        kT = integrator.kT.value_in_unit(unit.kilojoules_per_mole)
        self._move_engine.selectMove()
        move = self._move_engine.selected_move
        device_move = move.device_move() if hasattr(move, 'device_move') else None
        if device_move is None:
            raise NotImplementedError('many-walker runs need a move with an on-device descriptor '
                                      '(RandomLigandRotationMove, WaterTranslationMove)')
        # _syncStatesMDtoNCMC
        e_md0 = md.get_energy(True, False)[0]
        ncmc.copy_state_from(md)
        e_nc0 = ncmc.get_energy(True, False)[0]
        # _stepNCMC
        self._ncmc_sim.currentIter = self.currentIter
        failed = False
        try:
            self._ncmc_sim.context = move.beforeMove(self._ncmc_sim.context)
            if 0 <= moveStep < nstepsNC:
                self._run_with_device_move(int(nstepsNC), int(moveStep), device_move)
            else:
                self._ncmc_sim.step(int(nstepsNC))
            self._ncmc_sim.context = move.afterMove(self._ncmc_sim.context)
        except openmm.OpenMMException as err:      # a walker blew up: its work reads NaN and it is rejected below
            logger.error(err)
            failed = True
        finally:
            integrator._scheduled_move = None
        R = ncmc.n_replicas
        work = np.array([ncmc.get_global('protocol_work', r) for r in range(R)])
        e_nc1 = ncmc.get_energy(True, False)[0]
        # _computeAlchemicalCorrection: E_alch(x1) on the MD system
        alch.copy_state_from(ncmc, positions=True, velocities=False, box=True)
        e_alch1 = alch.get_energy(True, False)[0]
        correction = -(e_nc0 - e_md0 + e_alch1 - e_nc1) / kT
        correction = np.where(np.isfinite(work) & np.isfinite(correction), correction, 0.0)
        # _acceptRejectMove
        accepted, logp, logu = ncmc.accept_reject(correction)
        accepted = accepted.astype(bool) & np.isfinite(work)
        if accepted.any():
            md.copy_state_from(ncmc, positions=True, velocities=False, box=False, mask=accepted.astype(np.int32))
        self.accept += int(accepted.sum())
        self.reject += int(R - accepted.sum())
        # _resetSimulations, _stepMD
        self._ncmc_sim.currentStep = 0
        integrator.reset()
        self._md_sim.context.setVelocitiesToTemperature(temperature)
        self._md_sim.currentIter = self.currentIter
        self._md_sim.step(int(nstepsMD))
        return {'work_kT': work / kT, 'correction': correction, 'log_accept': logp, 'log_u': logu, 'accepted': accepted,
                'failed': failed, 'energies': {'md0': e_md0, 'ncmc0': e_nc0, 'ncmc1': e_nc1, 'alch1': e_alch1}}

    def _runWalkers(self, nIter, nstepsNC, moveStep, nstepsMD, temperature):
        """``run`` for contexts that hold several walkers; one all-gather of the per-walker statistics per iteration
        (NCCL when the job runs under torchrun with GPUs, none in a single process)."""
        from . import parallel
        import torch.distributed as dist
        rank, world = parallel.rank_world()
        R = self._ncmc_sim.context.getNumReplicas()
        distributed = world > 1 and dist.is_available() and dist.is_initialized()
        device = None
        if distributed and dist.get_backend() == 'nccl':
            device = 'cuda'
        local_ids = [rank + world * r for r in range(R)]
        temperature = temperature if unit.is_quantity(temperature) else temperature * unit.kelvin
        self.walker_history = []
        self.walker_records = []          # this rank's own per-walker records (work, correction, energies, flags)
        n_total = 0
        for it in range(int(nIter)):
            self.currentIter = it
            logger.info('BLUES Iteration: %s (%d walkers on this rank)' % (it, R))
            rec = self._iterateWalkers(nstepsNC, moveStep, nstepsMD, temperature)
            self.walker_records.append(rec)
            stats = parallel.gather_walker_stats(local_ids, rec['work_kT'], rec['log_accept'], rec['accepted'], device=device) \
                if distributed else {'walker': np.asarray(local_ids), 'work_kT': rec['work_kT'],
                                     'log_accept': rec['log_accept'], 'accepted': rec['accepted'].astype(int)}
            self.walker_history.append(stats)
            n_total += len(stats['walker'])
            logger.info('Iteration %d: %d / %d walkers accepted, mean work %.3f kT' % (
                it, int(np.sum(stats['accepted'])), len(stats['walker']), float(np.nanmean(stats['work_kT']))))
        total_acc = sum(int(np.sum(s['accepted'])) for s in self.walker_history)
        self.acceptRatio = total_acc / float(max(n_total, 1))
        logger.info('Acceptance Ratio: %s' % self.acceptRatio)
        logger.info('nIter: %s ' % nIter)

    def run(self, nIter=0, nstepsNC=0, moveStep=0, nstepsMD=0, temperature=300, write_move=False, **config):
        """NCMC → accept/reject → MD, ``nIter`` times (``blues/simulation.py:1215-1257``); arguments left at 0 come
        from the configuration."""
        cfg = self._config or {}
        nIter, nstepsNC, nstepsMD, moveStep = (given or cfg[key] for given, key in (
            (nIter, 'nIter'), (nstepsNC, 'nstepsNC'), (nstepsMD, 'nstepsMD'), (moveStep, 'moveStep')))
        logger.info('Running %i BLUES iterations...' % (nIter))
        if self._ncmc_sim.context.getNumReplicas() > 1:
            return self._runWalkers(nIter, nstepsNC, moveStep, nstepsMD, temperature)
        for it in range(int(nIter)):
            self.currentIter = it
            logger.info('BLUES Iteration: %s' % it)
            self._syncStatesMDtoNCMC()
            self._stepNCMC(nstepsNC, moveStep)
            self._acceptRejectMove(write_move)
            self._resetSimulations(temperature)
            self._stepMD(nstepsMD)
        self.acceptRatio = self.accept / float(nIter)
        logger.info('Acceptance Ratio: %s' % self.acceptRatio)
        logger.info('nIter: %s ' % nIter)


class MonteCarloSimulation(BLUESSimulation):
    """Plain Metropolis Monte Carlo with the same moves, no NCMC relaxation (``blues/simulation.py:1260-1335``)."""

    def __init__(self, simulations, config=None):
        super(MonteCarloSimulation, self).__init__(simulations, config)

    def _stepMC_(self):
        """Apply the selected move to the MD context and record the trial state."""
        self._move_engine.selectMove()
        trial = self._move_engine.runEngine(self._md_sim.context)
        self._setStateTable('md', 'state1', self.getStateFromContext(trial, self._state_keys))

    def _acceptRejectMove(self, temperature=None):
        """exp(−ΔE/kT) Metropolis test between the recorded MD states, then fresh velocities."""
        old, trial = self.stateTable['md']['state0'], self.stateTable['md']['state1']
        kT = self._ncmc_sim.context._integrator.kT
        log_p = -1.0 * (trial['potential_energy'] - old['potential_energy']) / kT
        log_u = math.log(np.random.random())
        accepted = log_p > log_u
        if accepted:
            self.accept += 1
            logger.info('MC MOVE ACCEPTED: work_mc {} > randnum {}'.format(log_p, log_u))
        else:
            self.reject += 1
            logger.info('MC MOVE REJECTED: work_mc {} < {}'.format(log_p, log_u))
        self._md_sim.context.setPositions((trial if accepted else old)['positions'])
        self._md_sim.context.setVelocitiesToTemperature(temperature)

    def run(self, nIter=0, mc_per_iter=0, nstepsMD=0, temperature=300, write_move=False):
        cfg = self._config or {}
        nIter = nIter or cfg['nIter']
        nstepsMD = nstepsMD or cfg['nstepsMD']
        mc_per_iter = mc_per_iter or cfg['mc_per_iter']
        self._syncStatesMDtoNCMC()
        for it in range(nIter):
            self.currentIter = it
            logger.info('MonteCarlo Iteration: %s' % it)
            for _ in range(mc_per_iter):
                self._syncStatesMDtoNCMC()
                self._stepMC_()
                self._acceptRejectMove(temperature)
            self._stepMD(nstepsMD)
