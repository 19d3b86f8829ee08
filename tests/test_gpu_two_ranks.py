"""Two ranks on one GPU box: every rank holds its share of the walkers of `simulation: nReplicas` and the per-walker
statistics of an iteration are all-gathered (gloo here: NCCL refuses two ranks on one device; bench.py --gpus N uses
NCCL with one GPU per rank)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %r)
import torch.distributed as dist
dist.init_process_group('gloo')
from blues_b200 import unit, utils
from blues_b200.structure import Structure
from blues_b200.simulation import SystemFactory, SimulationFactory, BLUESSimulation
from blues_b200.moves import RandomLigandRotationMove, MoveEngine
import tests.test_gpu_api as api
structure = Structure.load_npz(os.path.join(api.GOLDEN, 'tol_parm.npz'))
idx = utils.atomIndexfromTop('LIG', structure.topology)
systems = SystemFactory(structure, idx, api.system_cfg())
cfg = api.sim_cfg()
cfg.update(nIter=2, nstepsNC=10, nstepsMD=4, nReplicas=5, seed=77, devices=[0])
sims = SimulationFactory(systems, MoveEngine(RandomLigandRotationMove(structure, 'LIG')), cfg)
for s in (sims.md, sims.alch, sims.ncmc):
    s.minimizeEnergy(maxIterations=100)
b = BLUESSimulation(sims)
b.run()
rank = dist.get_rank()
out = {'rank': rank, 'local': sims.ncmc.context.getNumReplicas(), 'seed': sims.ncmc.integrator.getRandomNumberSeed(),
       'walkers': [h['walker'].tolist() for h in b.walker_history], 'work': [h['work_kT'].tolist() for h in b.walker_history],
       'ratio': b.acceptRatio}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'result_%%d.json' %% rank), 'w') as fh:
    json.dump(out, fh)          # one file per rank: two ranks printing to one pipe interleave their lines
dist.destroy_process_group()
'''


def test_walkers_shard_over_two_ranks_and_statistics_are_gathered(tmp_path):
    import json
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    p = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr',
                        '127.0.0.1', '--master-port', '29731', str(script)], capture_output=True, text=True, timeout=600,
                       env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    res = sorted((json.load(open(str(f))) for f in tmp_path.glob('result_*.json')), key=lambda r: r['rank'])
    assert len(res) == 2
    assert [r['local'] for r in res] == [3, 2]                      # 5 walkers round-robin over 2 ranks
    assert res[0]['seed'] != res[1]['seed']
    for r in res:
        assert r['walkers'] == [[0, 1, 2, 3, 4], [0, 1, 2, 3, 4]]   # every rank sees the whole job after the gather
    assert res[0]['work'] == res[1]['work']
    assert len(set(res[0]['work'][0])) == 5                         # five different trajectories
    assert res[0]['ratio'] == res[1]['ratio']
