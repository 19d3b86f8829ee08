"""BLUES with water translation moves on the B200 engine — the reference's examples/example_water.py, same calls:
a water inside a sphere around two ligand atoms trades places with the alchemical water, is translated at the protocol
midpoint and must still be inside the sphere at the end (all three hooks run on the device).

    cd examples && python example_water.py [rotmove_b200.yml]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from blues_b200 import unit                                                # noqa: E402
from blues_b200.moves import WaterTranslationMove, MoveEngine              # noqa: E402
from blues_b200.simulation import SystemFactory, SimulationFactory, BLUESSimulation   # noqa: E402
from blues_b200.settings import Settings                                   # noqa: E402


def watermove(yaml_file, **simulation_overrides):
    opt = Settings(yaml_file).asDict()
    opt['simulation'].update(simulation_overrides)
    structure = opt['Structure']
    water = WaterTranslationMove(structure, water_name=['WAT', 'HOH'], protein_selection='(index 0) or (index 1)',
                                 radius=0.9 * unit.nanometers)
    water_mover = MoveEngine(water)
    systems = SystemFactory(structure, water.atom_indices, opt['system'])
    if 'restraints' in opt:
        systems.md = systems.restrain_positions(structure, systems.md, **opt['restraints'])
        systems.alch = systems.restrain_positions(structure, systems.alch, **opt['restraints'])
    simulations = SimulationFactory(systems, water_mover, opt['simulation'], opt['md_reporters'], opt['ncmc_reporters'])
    for sim in (simulations.md, simulations.alch, simulations.ncmc):
        sim.minimizeEnergy(maxIterations=200)
    blues = BLUESSimulation(simulations, opt['simulation'])
    blues.run()
    return blues


if __name__ == '__main__':
    watermove(sys.argv[1] if len(sys.argv) > 1 else 'rotmove_b200.yml')
