"""Reporters and logging with the surface of ``blues/reporters.py`` (+ the writers of ``blues/formats.py``).

The reporter *protocol* is OpenMM's: ``describeNextReport(simulation) -> (steps, pos, vel, frc, ene)`` and
``report(simulation, state)`` (``blues/reporters.py:534-561, 777-802``); ``steps == -1`` (frame-index mode
outside a frame) is tolerated by ``blues_b200.mm.Simulation``.  Provided: ``addLoggingLevel`` /
``init_logger`` / ``LoggerFormatter`` (``reporters.py:27-126``, ``formats.py:21-84``), ``ReporterConfig``
(``reporters.py:129-242``), ``BLUESStateDataReporter`` (``reporters.py:436-728``), ``NetCDF4Reporter``
(``reporters.py:731-865`` — AMBER NetCDF-3 trajectory with ``protocolWork`` / ``alchemicalLambda`` variables,
written with ``scipy.io.netcdf_file``), ``RestartReporter`` (ASCII ``.rst7``) and a plain ``StateDataReporter``.
"""
import logging
import sys
import time

import numpy as np

from . import unit

VELUNIT = unit.angstrom / unit.picosecond


def addLoggingLevel(levelName, levelNum, methodName=None):
    """Register a new logging level + convenience methods (``logging.REPORT`` / ``logger.report``)."""
    methodName = methodName or levelName.lower()
    for owner, attr in ((logging, levelName), (logging, methodName), (logging.getLoggerClass(), methodName)):
        if hasattr(owner, attr):
            logging.warning('{} already defined in {}'.format(attr, getattr(owner, '__name__', owner)))

    def logForLevel(self, message, *args, **kwargs):
        if self.isEnabledFor(levelNum):
            self._log(levelNum, message, args, **kwargs)

    def logToRoot(message, *args, **kwargs):
        logging.log(levelNum, message, *args, **kwargs)

    logging.addLevelName(levelNum, levelName)
    setattr(logging, levelName, levelNum)
    setattr(logging.getLoggerClass(), methodName, logForLevel)
    setattr(logging, methodName, logToRoot)


class LoggerFormatter(logging.Formatter):
    """Per-level formats: bare message for INFO/REPORT, level-tagged for the rest (``blues/formats.py:21-84``)."""
    FORMATS = {'DEBUG': '%(levelname)s: [%(module)s.%(funcName)s] %(message)s', 'INFO': '%(levelname)s: %(message)s',
               'REPORT': '%(message)s', 'WARNING': '%(levelname)s: %(message)s',
               'ERROR': '%(levelname)s: [%(module)s.%(funcName)s] %(message)s',
               'CRITICAL': '%(levelname)s: [%(module)s.%(funcName)s] %(message)s'}

    def __init__(self):
        if not hasattr(logging, 'REPORT'):
            addLoggingLevel('REPORT', logging.WARNING - 5)
        super(LoggerFormatter, self).__init__(fmt='%(levelname)s: %(msg)s', datefmt='%H:%M:%S')
        self._formatters = {k: logging.Formatter(v) for k, v in self.FORMATS.items()}

    def format(self, record):
        return self._formatters.get(record.levelname, self._formatters['INFO']).format(record)


def init_logger(logger, level=logging.INFO, stream=True, outfname=time.strftime("blues-%Y%m%d-%H%M%S")):
    """Attach stdout / file handlers with :class:`LoggerFormatter` (``blues/reporters.py:88-126``)."""
    fmt = LoggerFormatter()
    if stream:
        h = logging.StreamHandler(stream=sys.stdout)
        h.setFormatter(fmt)
        logger.addHandler(h)
    if outfname:
        fh = logging.FileHandler(str(outfname) + '.log')
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    logger.addHandler(logging.NullHandler())
    logger.setLevel(level)
    return logger


class ReporterConfig(object):
    """Reporter sections of the YAML → reporter objects (``blues/reporters.py:129-242``)."""

    def __init__(self, outfname, reporter_config, logger=None):
        self._outfname = str(outfname)
        self._cfg = reporter_config
        self._logger = logger
        self.trajectory_interval = 0

    def _name(self, section):
        return str(self._cfg[section].get('outfname', self._outfname))

    def _args(self, section):
        return {k: v for k, v in self._cfg[section].items() if k != 'outfname'}

    def makeReporters(self):
        out = []
        if 'state' in self._cfg:
            out.append(StateDataReporter(self._name('state') + '.ene', **self._args('state')))
        if 'traj_netcdf' in self._cfg:
            self.trajectory_interval = self._cfg['traj_netcdf'].get('reportInterval', 0)
            out.append(NetCDF4Reporter(self._name('traj_netcdf') + '.nc', **self._args('traj_netcdf')))
        if 'restart' in self._cfg:
            out.append(RestartReporter(self._name('restart') + '.rst7', **dict({'netcdf': True}, **self._args('restart'))))
        if 'progress' in self._cfg:
            args = self._args('progress')
            args.setdefault('progress', True)
            args.setdefault('step', True)
            out.append(StateDataReporter(self._name('progress') + '.prog', **args))
        if 'stream' in self._cfg:
            if not self._logger:
                self._logger = logging.getLogger(__name__)
            out.append(BLUESStateDataReporter(self._logger, **self._cfg['stream']))
        return out


class _Periodic(object):
    """describeNextReport logic shared by all reporters: fixed interval or explicit 1-based frame indices."""

    def _init_schedule(self, reportInterval, frame_indices):
        self._reportInterval = int(reportInterval) if reportInterval else 1
        self.frame_indices = [int(x) - 1 for x in frame_indices] if frame_indices else []

    def _next(self, simulation):
        if self.frame_indices:
            return 1 if simulation.currentStep in self.frame_indices else -1
        return self._reportInterval - simulation.currentStep % self._reportInterval

    def stepsToNextFrameIndex(self, simulation):
        """Distance from ``currentStep`` to the next listed frame index (0: reporting now; None: no index ahead or
        interval schedule).  ``describeNextReport`` keeps the reference's 1 / -1 answer; the chunked steppers
        (``mm.Simulation._simulate``, ``BLUESSimulation._run_with_device_move``) use this to stop at listed frames."""
        if not self.frame_indices:
            return None
        ahead = [i - simulation.currentStep for i in self.frame_indices if i >= simulation.currentStep]
        return min(ahead) if ahead else None


def _remaining(total, initial, elapsed_s, elapsed_steps):
    if elapsed_steps == 0:
        return '--'
    rem = int((total - initial) * elapsed_s / elapsed_steps - elapsed_s)
    d, rem = divmod(rem, 86400)
    h, rem = divmod(rem, 3600)
    m, s = divmod(rem, 60)
    if d > 0:
        return "%d:%d:%02d:%02d" % (d, h, m, s)
    if h > 0:
        return "%d:%02d:%02d" % (h, m, s)
    return "%d:%02d" % (m, s) if m > 0 else "0:%02d" % s


class BLUESStateDataReporter(_Periodic):
    """Streams ``title: v1<TAB>v2…`` rows to a logger (or file): step, speed in ns/day, lambda, protocol work…

    Column switches and their order follow ``blues/reporters.py:602-728``.
    """
    # (flag attribute, header, needs energy)
    _COLUMNS = [('currentIter', 'Iter', False), ('progress', 'Progress (%)', False), ('step', 'Step', False),
                ('time', 'Time (ps)', False), ('alchemicalLambda', 'alchemicalLambda', False),
                ('protocolWork', 'protocolWork', False), ('potentialEnergy', 'Potential Energy (kJ/mole)', True),
                ('kineticEnergy', 'Kinetic Energy (kJ/mole)', True), ('totalEnergy', 'Total Energy (kJ/mole)', True),
                ('temperature', 'Temperature (K)', True), ('volume', 'Box Volume (nm^3)', False),
                ('density', 'Density (g/mL)', False), ('speed', 'Speed (ns/day)', False),
                ('elapsedTime', 'Elapsed Time (s)', False), ('remainingTime', 'Time Remaining', False)]

    def __init__(self, file, reportInterval=1, frame_indices=[], title='', step=False, time=False,
                 potentialEnergy=False, kineticEnergy=False, totalEnergy=False, temperature=False, volume=False,
                 density=False, progress=False, remainingTime=False, speed=False, elapsedTime=False, separator='\t',
                 systemMass=None, totalSteps=None, protocolWork=False, alchemicalLambda=False, currentIter=False):
        self._init_schedule(reportInterval, frame_indices)
        self._flags = dict(currentIter=currentIter, progress=progress, step=step, time=time,
                           alchemicalLambda=alchemicalLambda, protocolWork=protocolWork,
                           potentialEnergy=potentialEnergy, kineticEnergy=kineticEnergy, totalEnergy=totalEnergy,
                           temperature=temperature, volume=volume, density=density, speed=speed,
                           elapsedTime=elapsedTime, remainingTime=remainingTime)
        if (progress or remainingTime) and totalSteps is None:
            raise ValueError('Reporting progress or remaining time requires total steps to be specified')
        self._openedFile = isinstance(file, str)
        self._out = open(file, 'w') if self._openedFile else file
        self.log = self._out
        self.title = title
        self._separator = separator
        self._totalSteps = totalSteps
        self._totalMass = systemMass
        self._needEnergy = any(self._flags[k] for k, _, e in self._COLUMNS if e)
        self._needsPositions = self._needsVelocities = self._needsForces = False
        self._hasInitialized = False

    def describeNextReport(self, simulation):
        return (self._next(simulation), self._needsPositions, self._needsVelocities, self._needsForces, self._needEnergy)

    def _emit(self, text):
        out = self.log
        if hasattr(out, 'report'):
            out.report(text)
        elif hasattr(out, 'info'):
            out.info(text)
        else:
            out.write(text + '\n')
            out.flush()

    def _initializeConstants(self, simulation):
        system = simulation.system
        masses = np.asarray(system.masses)
        dof = 3 * int(np.count_nonzero(masses > 0)) - len(simulation.context._topo['constraints'])
        if simulation.context._topo['remove_cm']:
            dof -= 3
        self._dof = max(dof, 1)
        if self._totalMass is None:
            self._totalMass = float(masses.sum())

    def _constructHeaders(self):
        return [h for k, h, _ in self._COLUMNS if self._flags[k]]

    def _checkForErrors(self, simulation, state):
        if self._needEnergy:
            e = state.getPotentialEnergy().value_in_unit(unit.kilojoules_per_mole)
            if np.isnan(e) or np.isinf(e):
                raise ValueError('Energy is %s' % e)

    def _constructReportValues(self, simulation, state):
        f = self._flags
        now = time.time()
        box = state.getPeriodicBoxVectors(asNumpy=True).value_in_unit(unit.nanometers)
        vol = float(box[0][0] * box[1][1] * box[2][2])
        v = []
        if f['currentIter']:
            v.append(getattr(simulation, 'currentIter', 0))
        if f['progress']:
            v.append('%.1f%%' % (100.0 * simulation.currentStep / self._totalSteps))
        if f['step']:
            v.append(simulation.currentStep)
        if f['time']:
            v.append(state.getTime().value_in_unit(unit.picosecond))
        if f['alchemicalLambda']:
            v.append(simulation.integrator.getGlobalVariableByName('lambda'))
        if f['protocolWork']:
            v.append(simulation.integrator.get_protocol_work(dimensionless=True))
        if self._needEnergy:
            pe = state.getPotentialEnergy().value_in_unit(unit.kilojoules_per_mole)
            ke = state.getKineticEnergy().value_in_unit(unit.kilojoules_per_mole)
            if f['potentialEnergy']:
                v.append(pe)
            if f['kineticEnergy']:
                v.append(ke)
            if f['totalEnergy']:
                v.append(pe + ke)
            if f['temperature']:
                v.append(2 * ke / (self._dof * unit.MOLAR_GAS_CONSTANT_R.value_in_unit(
                    unit.kilojoules_per_mole / unit.kelvin)))
        if f['volume']:
            v.append(vol)
        if f['density']:
            v.append(self._totalMass / vol * 1.66053886e-3)     # dalton/nm^3 → g/mL
        if f['speed']:
            days = (now - self._initialClockTime) / 86400.0
            ns = (state.getTime() - self._initialSimulationTime).value_in_unit(unit.nanosecond)
            v.append('%.3g' % (ns / days) if days > 0.0 else '--')
        if f['elapsedTime']:
            v.append(now - self._initialClockTime)
        if f['remainingTime']:
            v.append(_remaining(self._totalSteps, self._initialSteps, now - self._initialClockTime,
                                simulation.currentStep - self._initialSteps))
        return v

    def report(self, simulation, state):
        if not self._hasInitialized:
            self._initializeConstants(simulation)
            self._emit('#"%s"' % ('"' + self._separator + '"').join(self._constructHeaders()))
            self._initialClockTime = time.time()
            self._initialSimulationTime = state.getTime()
            self._initialSteps = simulation.currentStep
            self._hasInitialized = True
        self._checkForErrors(simulation, state)
        values = self._constructReportValues(simulation, state)
        line = self._separator.join(str(x) for x in values)
        self._emit('%s: %s' % (self.title, line) if self.title else line)

    def __del__(self):
        if getattr(self, '_openedFile', False):
            self._out.close()


class StateDataReporter(BLUESStateDataReporter):
    """File-backed state reporter (``.ene`` / ``.prog``); energies on by default like parmed's."""

    def __init__(self, file, reportInterval=1, **kwargs):
        if not any(k in kwargs for k in ('potentialEnergy', 'kineticEnergy', 'totalEnergy', 'temperature', 'progress')):
            kwargs.update(step=True, time=True, potentialEnergy=True, kineticEnergy=True, totalEnergy=True,
                          temperature=True, volume=True)
        kwargs.pop('title', None)
        super(StateDataReporter, self).__init__(file, reportInterval, **kwargs)


class RestartReporter(_Periodic):
    """AMBER restart (positions, velocities, box) every ``reportInterval`` steps — NetCDF (``AMBERRESTART``) when
    ``netcdf=True`` as the reference's ``ReporterConfig`` asks (``blues/reporters.py:224``), ASCII otherwise; resume
    with the YAML key ``structure: restart:`` (``blues/settings.py:76-85``), either format."""

    def __init__(self, file, reportInterval=1, write_multiple=False, netcdf=False, write_velocities=True, **kwargs):
        self._init_schedule(reportInterval, [])
        self.fname = file
        self.write_multiple = write_multiple
        self.netcdf = netcdf
        self.write_velocities = write_velocities

    def describeNextReport(self, simulation):
        return (self._next(simulation), True, self.write_velocities, False, False)

    def report(self, simulation, state):
        from .structure import Structure, write_inpcrd, write_netcdf_restart
        xyz = state.getPositions(asNumpy=True).value_in_unit(unit.angstroms)
        vel = state.getVelocities(asNumpy=True).value_in_unit(VELUNIT) if self.write_velocities else None
        bv = state.getPeriodicBoxVectors(asNumpy=True).value_in_unit(unit.angstroms)
        box = [bv[0][0], bv[1][1], bv[2][2], 90.0, 90.0, 90.0]
        fname = self.fname + ('.%d' % simulation.currentStep if self.write_multiple else '')
        if self.netcdf:
            write_netcdf_restart(fname, xyz, vel, box, time=state.getTime().value_in_unit(unit.picoseconds))
            return
        s = Structure()
        s.n_atoms = simulation.system.getNumParticles()
        s.coordinates = xyz
        if vel is not None:
            s._velocities = vel
        s.box = box
        with open(fname, 'w') as f:
            write_inpcrd(s, f)


class NetCDF4Reporter(_Periodic):
    """AMBER-convention NetCDF trajectory with optional ``protocolWork`` (kT) and ``alchemicalLambda`` variables
    (``blues/reporters.py:731-865``, ``blues/formats.py:476-690``).  AMBER trajectories are NetCDF-3 64-bit offset
    files, written here through ``scipy.io.netcdf_file``."""

    def __init__(self, file, reportInterval=1, frame_indices=[], crds=True, vels=False, frcs=False, protocolWork=False,
                 alchemicalLambda=False):
        self._init_schedule(reportInterval, frame_indices)
        self.fname = file
        self.crds, self.vels, self.frcs = crds, vels, frcs
        self.protocolWork, self.alchemicalLambda = protocolWork, alchemicalLambda
        self._nc = None
        self._frame = 0

    def describeNextReport(self, simulation):
        return (self._next(simulation), self.crds, self.vels, self.frcs, False)

    def _open(self, natom, has_box):
        from scipy.io import netcdf_file
        nc = netcdf_file(self.fname, 'w', version=2)
        nc.Conventions, nc.ConventionVersion = 'AMBER', '1.0'
        nc.application, nc.program, nc.programVersion = 'blues_b200', 'blues_b200', '0.1'
        nc.createDimension('frame', None)
        nc.createDimension('spatial', 3)
        nc.createDimension('atom', natom)
        v = nc.createVariable('spatial', 'c', ('spatial',))
        v[:] = np.asarray(list('xyz'), dtype='S1')
        t = nc.createVariable('time', 'f', ('frame',))
        t.units = 'picosecond'
        if has_box:
            nc.createDimension('cell_spatial', 3)
            nc.createDimension('cell_angular', 3)
            nc.createDimension('label', 5)
            nc.createVariable('cell_spatial', 'c', ('cell_spatial',))[:] = np.asarray(list('abc'), dtype='S1')
            ca = nc.createVariable('cell_angular', 'c', ('cell_angular', 'label'))
            ca[:] = np.asarray([list('alpha'), list('beta '), list('gamma')], dtype='S1')
            nc.createVariable('cell_lengths', 'd', ('frame', 'cell_spatial')).units = 'angstrom'
            nc.createVariable('cell_angles', 'd', ('frame', 'cell_angular')).units = 'degree'
        for flag, name, units in ((self.crds, 'coordinates', 'angstrom'), (self.vels, 'velocities', 'angstrom/picosecond'),
                                  (self.frcs, 'forces', 'kilocalorie/mole/angstrom')):
            if flag:
                nc.createVariable(name, 'f', ('frame', 'atom', 'spatial')).units = units
        if self.protocolWork:
            nc.createVariable('protocolWork', 'f', ('frame',)).units = 'kT'
        if self.alchemicalLambda:
            nc.createVariable('alchemicalLambda', 'f', ('frame',)).units = 'unitless'
        self._nc = nc

    def report(self, simulation, state):
        box = state.getPeriodicBoxVectors(asNumpy=True).value_in_unit(unit.angstroms)
        has_box = simulation.context._topo['nb_method'] != 0
        if self._nc is None:
            self._open(simulation.system.getNumParticles(), has_box)
        nc, k = self._nc, self._frame
        nc.variables['time'][k] = state.getTime().value_in_unit(unit.picosecond)
        if has_box:
            nc.variables['cell_lengths'][k] = [box[0][0], box[1][1], box[2][2]]
            nc.variables['cell_angles'][k] = [90.0, 90.0, 90.0]
        if self.crds:
            nc.variables['coordinates'][k] = state.getPositions(asNumpy=True).value_in_unit(unit.angstroms)
        if self.vels:
            nc.variables['velocities'][k] = state.getVelocities(asNumpy=True).value_in_unit(VELUNIT)
        if self.frcs:
            nc.variables['forces'][k] = state.getForces(asNumpy=True).value_in_unit(
                unit.kilocalories_per_mole / unit.angstroms)
        if self.protocolWork:
            nc.variables['protocolWork'][k] = simulation.integrator.get_protocol_work(dimensionless=True)
        if self.alchemicalLambda:
            nc.variables['alchemicalLambda'][k] = simulation.integrator.getGlobalVariableByName('lambda')
        self._frame += 1
        nc.flush()

    def close(self):
        if self._nc is not None:
            self._nc.close()
            self._nc = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
