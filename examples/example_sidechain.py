"""BLUES with side-chain rotations on the B200 engine — the reference's examples/example_sidechain.py, same calls.

    cd examples && python example_sidechain.py [sidechain_b200.yml]

Valine dipeptide in vacuum: ``SideChainMove`` perceives the rotatable chi bond of residue 1 from the bond graph (the
reference needs OpenEye for that), NCMC relaxes every proposal; afterwards the chi1 dihedral of every stored MD frame is
read back from the NetCDF trajectory, as the reference's script does with mdtraj.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np                                                          # noqa: E402
from blues_b200 import trajectory as md                                     # noqa: E402
from blues_b200.moves import SideChainMove, MoveEngine                      # noqa: E402
from blues_b200.simulation import SystemFactory, SimulationFactory, BLUESSimulation   # noqa: E402
from blues_b200.settings import Settings                                    # noqa: E402


def sidechain(yaml_file, **simulation_overrides):
    cfg = Settings(yaml_file).asDict()
    cfg['simulation'].update(simulation_overrides)
    structure = cfg['Structure']
    move = SideChainMove(structure, [1])
    mover = MoveEngine(move)
    systems = SystemFactory(structure, move.atom_indices, cfg['system'])
    simulations = SimulationFactory(systems, mover, cfg['simulation'], cfg['md_reporters'], cfg['ncmc_reporters'])
    blues = BLUESSimulation(simulations, cfg['simulation'])
    blues.run()
    # analysis: N - CA - CB - CG1 of the valine (atoms 0, 4, 6, 8 of the reference's vacDivaline topology)
    out = os.path.join(cfg['output_dir'], cfg['outfname'])
    for rep in simulations.md.reporters:
        if hasattr(rep, 'close'):
            rep.close()
    traj = md.load_netcdf(out + '.nc')
    dihedrals = md.compute_dihedrals(traj, np.array([[0, 4, 6, 8]]))
    with open(out + '-dihedrals.txt', 'w') as fh:
        for value in dihedrals:
            fh.write('%s\n' % str(value)[1:-1])
    blues.dihedrals = dihedrals
    return blues


if __name__ == '__main__':
    sidechain(sys.argv[1] if len(sys.argv) > 1 else 'sidechain_b200.yml')
