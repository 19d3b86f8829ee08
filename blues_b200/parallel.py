"""Walker sharding across ranks (one process per GPU) and the statistics gather.

Independent NCMC walkers are the only parallel axis: walker w lives on rank ``w % world``; there is
no per-step communication.  After an NCMC iteration each rank contributes ``{protocol_work, log_accept, accepted}``
per walker to one all-gather (NCCL on GPUs, gloo in the CPU tests) so that rank 0 can log acceptance statistics.
"""
import os

import numpy as np


def rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def shard_walkers(n_walkers, rank, world):
    """Global walker ids owned by ``rank`` (round-robin, so every rank gets ⌈n/world⌉ or ⌊n/world⌋ walkers)."""
    return list(range(rank, n_walkers, world))


def walker_seed(base_seed, walker_id):
    """One Philox key per job; the walker id selects the subsequence (engine: replica field of the counter)."""
    return int(base_seed) + 1000003 * int(walker_id)


def gather_walker_stats(local_ids, work_kT, log_accept, accepted, device=None):
    """All-gather per-walker statistics; returns dict of numpy arrays ordered by global walker id (on every rank)."""
    import torch
    import torch.distributed as dist
    local = np.stack([np.asarray(local_ids, float), np.asarray(work_kT, float), np.asarray(log_accept, float),
                      np.asarray(accepted, float)], axis=1).reshape(-1, 4)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        allrows = local
    else:
        world = dist.get_world_size()
        n = torch.tensor([len(local)], dtype=torch.int64, device=device)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        nmax = int(max(c.item() for c in counts))
        buf = torch.full((nmax, 4), float('nan'), dtype=torch.float64, device=device)
        if len(local):
            buf[:len(local)] = torch.from_numpy(local).to(buf.device)
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
        allrows = np.concatenate([o.cpu().numpy()[:int(c.item())] for o, c in zip(out, counts)], axis=0)
    order = np.argsort(allrows[:, 0], kind='stable')
    allrows = allrows[order]
    return {'walker': allrows[:, 0].astype(int), 'work_kT': allrows[:, 1], 'log_accept': allrows[:, 2],
            'accepted': allrows[:, 3].astype(int)}


def max_over_ranks(value, device=None):
    """Max of a Python float over ranks (timing rule: the slowest rank defines the step time)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
