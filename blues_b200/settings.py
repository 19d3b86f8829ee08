"""YAML / JSON / dict configuration → the nested dict BLUES' factories consume (``blues/settings.py:13-322``).

Sections and keys are the reference's (``examples/rotmove_cuda.yml:9-102``): ``output_dir``, ``outfname``,
``logger``, ``structure``, ``system`` (+ ``alchemical``), ``freeze``, ``restraints``, ``simulation``,
``md_reporters``, ``ncmc_reporters``.  Unit strings such as ``'10 * angstroms'`` become Quantities, enum names
(``PME``, ``HBonds``) become the ``blues_b200.system`` constants, NCMC step counts are completed by
``utils.calculateNCMCSteps`` and reporter sections are turned into reporter objects.
"""
import json
import logging
import os

import yaml

from . import reporters, system as app, unit, utils
from .structure import Rst7, load_file


class Settings(object):
    """``Settings(yaml_path_or_string_or_dict).asDict()``"""

    def __init__(self, config):
        config = Settings.load_yaml(config)
        if type(config) is dict:
            self.config = Settings.set_Parameters(config)

    @staticmethod
    def load_yaml(yaml_config):
        """A dict as is, the path of a YAML file, or YAML text (``blues/settings.py:33-58``)."""
        if isinstance(yaml_config, dict):
            return yaml_config
        is_file = os.path.isfile(str(yaml_config))
        try:
            if not is_file:
                return yaml.safe_load(yaml_config)
            with open(yaml_config, 'r') as stream:
                return yaml.safe_load(stream)
        except IOError:
            print("Unable to open file:", yaml_config)
            raise
        except yaml.YAMLError as err:
            mark = getattr(err, 'problem_mark', None)
            if mark is not None:
                print('YAML parsing error in file: {}\nError on Line:{} Column:{}'.format(yaml_config, mark.line + 1,
                                                                                        mark.column + 1))
            raise

    @staticmethod
    def set_Structure(config):
        """Load the structure; ``restart:`` overrides positions / velocities / box (``blues/settings.py:60-90``)."""
        sc = dict(config['structure'])
        restart = None
        if 'restart' in sc:
            rst7 = sc.pop('restart')
            config['Logger'].info('Restarting simulation from {}'.format(rst7))
            restart = Rst7(rst7)
        structure = load_file(**sc)
        if restart is not None:
            structure.positions = restart.positions
            structure.velocities = restart.velocities
            structure.box = restart.box
        config['structure'] = sc
        config['Structure'] = structure
        return config

    @staticmethod
    def set_Output(config):
        out_dir = config.setdefault('output_dir', '.')
        os.makedirs(str(out_dir), exist_ok=True)
        outfname = os.path.join(str(out_dir), config['outfname'])
        config['outfname'] = outfname
        config.setdefault('simulation', {})['outfname'] = outfname
        return config

    @staticmethod
    def set_Logger(config):
        lc = config.setdefault('logger', {'level': 'info', 'stream': True})
        level = str(lc.get('level', 'info')).upper()
        stream = lc.get('stream', True)
        outfname = lc.get('filename', config['outfname'])
        verbose = level == 'DEBUG'
        config['verbose'] = verbose
        config.setdefault('system', {})['verbose'] = verbose
        config['simulation']['verbose'] = verbose
        config['Logger'] = reporters.init_logger(logging.getLogger(), getattr(logging, level), stream, outfname)
        return config

    @staticmethod
    def set_Units(config):
        """Attach units to bare numbers / parse ``'value * unit'`` strings (``blues/settings.py:140-187``)."""
        default_units = {
            'nonbondedCutoff': unit.angstroms, 'switchDistance': unit.angstroms, 'implicitSolventKappa': unit.angstroms,
            'freeze_distance': unit.angstroms, 'temperature': unit.kelvins, 'hydrogenMass': unit.daltons,
            'dt': unit.picoseconds, 'friction': 1 / unit.picoseconds, 'pressure': unit.atmospheres,
            'weight': unit.kilocalories_per_mole / unit.angstroms ** 2}
        for param, unit_type in default_units.items():
            for section in ('system', 'simulation', 'freeze', 'restraints'):
                sec = config.get(section)
                if not isinstance(sec, dict) or param not in sec:
                    continue
                value = sec[param]
                if unit.is_quantity(value):
                    continue
                if '*' in str(value):
                    sec[param] = utils.parse_unit_quantity(value)
                else:
                    config['Logger'].warning("Units for '{} = {}' not specified. Setting units to '{}'".format(
                        param, value, unit_type))
                    sec[param] = value * unit_type
        return config

    @staticmethod
    def check_SystemModifications(config):
        if 'freeze' in config:
            for sel in ('freeze_center', 'freeze_solvent', 'freeze_selection'):
                if sel in config['freeze']:
                    utils.check_amber_selection(config['Structure'], config['freeze'][sel])
        if 'restraints' in config:
            utils.check_amber_selection(config['Structure'], config['restraints']['selection'])

    @staticmethod
    def set_Apps(config):
        valid = {'nonbondedMethod': ['NoCutoff', 'CutoffNonPeriodic', 'CutoffPeriodic', 'PME', 'Ewald'],
                 'constraints': [None, 'HBonds', 'HAngles', 'AllBonds']}
        for key, options in valid.items():
            if key in config.get('system', {}):
                value = config['system'][key]
                if value is None or str(value) == 'None':
                    config['system'][key] = None
                elif str(value) in options:
                    config['system'][key] = getattr(app, str(value))
                else:
                    config['Logger'].error("'{}' was not a valid option for '{}'. Valid options: {}".format(value, key, options))
        return config

    @staticmethod
    def set_ncmcSteps(config):
        for k, v in utils.calculateNCMCSteps(**config['simulation']).items():
            config['simulation'][k] = v
        return config

    @staticmethod
    def set_Reporters(config):
        logger = config['Logger']
        outfname = config['outfname']
        nstepsNC = config['simulation']['nstepsNC']
        moveStep = config['simulation']['moveStep']
        if 'md_reporters' in config:
            cfg = reporters.ReporterConfig(outfname, config['md_reporters'], logger)
            config['md_reporters'] = cfg.makeReporters()
            if cfg.trajectory_interval:
                config['simulation']['md_trajectory_interval'] = cfg.trajectory_interval
        else:
            logger.warning('Configuration for MD reporters were not set.')
        if 'ncmc_reporters' in config:
            for rep in config['ncmc_reporters'].values():
                if 'totalSteps' in rep:
                    rep['totalSteps'] = nstepsNC
                if 'frame_indices' in rep:
                    rep['frame_indices'] = [nstepsNC if x == -1 else (moveStep if x == 0.5 else x)
                                            for x in rep['frame_indices']]
            cfg = reporters.ReporterConfig(outfname + '-ncmc', config['ncmc_reporters'], logger)
            config['ncmc_reporters'] = cfg.makeReporters()
        else:
            logger.warning('Configuration for NCMC reporters were not set.')
        return config

    @staticmethod
    def set_Parameters(config):
        """Every section in the reference's order (``blues/settings.py:228-255``); errors are logged and re-raised."""
        steps = [Settings.set_Output, Settings.set_Logger]
        if 'structure' in config:
            steps += [Settings.set_Structure, lambda c: (Settings.check_SystemModifications(c), c)[1]]
        steps += [Settings.set_Units, Settings.set_Apps, Settings.set_ncmcSteps, Settings.set_Reporters]
        try:
            for step in steps:
                config = step(config)
        except Exception as err:
            if 'Logger' in config:
                config['Logger'].exception(err)
            raise
        return config

    def asDict(self):
        return self.config

    def asOrderedDict(self):
        from collections import OrderedDict
        return OrderedDict(sorted(self.config.items(), key=lambda t: t[0]))

    def asYAML(self):
        return yaml.dump(self.config)

    def asJSON(self, pprint=False):
        if pprint:
            return json.dumps(self.config, sort_keys=True, indent=2, skipkeys=True, default=str)
        return json.dumps(self.config, default=str)
