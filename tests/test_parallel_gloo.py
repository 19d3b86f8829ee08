"""CPU test of the N > 1 host path: world_size-2 gloo run of walker sharding + statistics gather + max-over-ranks."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_walkers, q):
    import torch.distributed as dist
    from blues_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ids = parallel.shard_walkers(n_walkers, rank, world)
    work = [10.0 * w + 0.5 for w in ids]                     # fake per-walker results, a function of the global id
    logp = [-x for x in work]
    acc = [w % 2 for w in ids]
    stats = parallel.gather_walker_stats(ids, work, logp, acc)
    tmax = parallel.max_over_ranks(1.0 + rank)
    q.put((rank, ids, {k: v.tolist() for k, v in stats.items()}, tmax, parallel.rank_world()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_walkers', [8, 5])
def test_sharding_and_gather_world2(n_walkers):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_walkers, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids0, ids1 = res[0][1], res[1][1]
    assert sorted(ids0 + ids1) == list(range(n_walkers)) and not set(ids0) & set(ids1)
    assert abs(len(ids0) - len(ids1)) <= 1
    for rank, ids, stats, tmax, rw in res:
        assert rw == (rank, 2)
        assert stats['walker'] == list(range(n_walkers))
        assert np.allclose(stats['work_kT'], [10.0 * w + 0.5 for w in range(n_walkers)])
        assert stats['accepted'] == [w % 2 for w in range(n_walkers)]
        assert tmax == 2.0


def test_single_process_passthrough():
    from blues_b200 import parallel
    stats = parallel.gather_walker_stats([1, 0], [2.0, 1.0], [-2.0, -1.0], [1, 0])
    assert stats['walker'].tolist() == [0, 1] and stats['work_kT'].tolist() == [1.0, 2.0]
    assert parallel.max_over_ranks(3.5) == 3.5
    assert parallel.walker_seed(7, 0) != parallel.walker_seed(7, 1)
