"""BLUES with random ligand rotations on the B200 engine — the reference's examples/example_rotmove.py, same calls.

    cd examples && python example_rotmove.py [rotmove_b200.yml]

Swap the imports for ``import blues_b200.compat; blues_b200.compat.install()`` followed by the reference's own
``from blues.moves import …`` lines and the script is the upstream one.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from blues_b200.moves import RandomLigandRotationMove, MoveEngine          # noqa: E402
from blues_b200.simulation import SystemFactory, SimulationFactory, BLUESSimulation   # noqa: E402
from blues_b200.settings import Settings                                   # noqa: E402


def rotmove(yaml_file, **simulation_overrides):
    cfg = Settings(yaml_file).asDict()
    cfg['simulation'].update(simulation_overrides)
    structure = cfg['Structure']
    ligand = RandomLigandRotationMove(structure, 'LIG')
    ligand_mover = MoveEngine(ligand)
    # the Systems are built outside SimulationFactory so that they can be modified first
    systems = SystemFactory(structure, ligand.atom_indices, cfg['system'])
    if 'freeze' in cfg:
        # freezing everything away from the ligand in the alchemical system speeds up the NCMC leg
        systems.alch = systems.freeze_radius(structure, systems.alch, **cfg['freeze'])
    simulations = SimulationFactory(systems, ligand_mover, cfg['simulation'], cfg['md_reporters'], cfg['ncmc_reporters'])
    for sim in (simulations.md, simulations.alch, simulations.ncmc):
        sim.minimizeEnergy(maxIterations=200)     # the TOL-parm start coordinates need relaxing
    blues = BLUESSimulation(simulations, cfg['simulation'])
    blues.run()
    return blues


if __name__ == '__main__':
    rotmove(sys.argv[1] if len(sys.argv) > 1 else 'rotmove_b200.yml')
