"""Multi-GPU partitioning of a scoring batch (host logic, no compute).

The (candidate, config) index space is embarrassingly parallel once every simulation has
its own seed.  Each rank takes a strided shard of the CONFIG axis (so every
rank holds every candidate and a balanced mix of sender counts), scores it with no data-path
collective, and only the per-candidate partial scores (n_cands doubles) are all-reduced.
"""
import numpy as np


def config_shard(n_cfgs, rank, world):
    """Indices of the configs rank `rank` scores: rank, rank + world, ..."""
    return np.arange(rank, n_cfgs, world)


def partial_scores(sims, n_cands, n_local_cfgs):
    """Per-candidate sum of utilities over this rank's configs, in config order
    (the accumulation of src/evaluator.cc:96)."""
    u = np.asarray(sims["utility"], dtype=np.float64).reshape(n_cands, n_local_cfgs)
    out = np.zeros(n_cands, dtype=np.float64)
    for k in range(n_local_cfgs):
        out += u[:, k]
    return out


def all_reduce_scores(partial, dist=None, device=None):
    """Sum the per-candidate partial scores over ranks (NCCL on GPU, gloo on CPU)."""
    import torch
    t = torch.tensor(partial, dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
