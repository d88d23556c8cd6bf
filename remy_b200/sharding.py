"""Partitioning ONE scoring batch over the ranks of a torch.distributed job (host logic, no compute).

The (candidate, config) index space is embarrassingly parallel once every simulation has its own seed
(DESIGN.md §1).  Each rank takes a shard of the CONFIG axis — the cost-sorted configs dealt round-robin, so
every rank holds every candidate and the same mix of sender counts and run lengths — and scores it with no
data-path collective.  What crosses NVLink afterwards is 8 bytes per simulation: the utilities are
all-gathered and every rank adds them per candidate in CONFIG ORDER, the order of the reference's
`score += ...` loop (src/evaluator.cc:96), so the per-candidate scores are byte-identical for any number of
ranks.  (An all-reduce of partial sums would be cheaper still but ties the last bits to the rank count.)

Used by bench.py (--workload improve-step / policy-sweep, NCCL) and by tests/test_multigpu_gloo.py (gloo, CPU).
Inside ONE process the C ABI partitions a batch over several devices itself (remy_b200_init_devices).
"""
import numpy as np


def config_shard(n_cfgs, rank, world):
    """Indices of the configs rank `rank` scores: rank, rank + world, ..."""
    return np.arange(rank, n_cfgs, world)


def config_cost_order(cfgs):
    """Configs by descending cost — sender count first, then link rate x run length, the key the library
    sorts its work list by (batch_prep.h::plan_batch) — ties in index order."""
    cost = cfgs["link_ppt"] * cfgs["simulation_ticks"]
    return np.lexsort((np.arange(len(cfgs)), -cost, -cfgs["num_senders"].astype(np.int64)))


def config_shard_balanced(cfgs, rank, world):
    """Indices (ascending) of the configs rank `rank` scores when the cost-sorted list is dealt round-robin."""
    return np.sort(config_cost_order(cfgs)[rank::world])


def shard_layout(cfgs_all, world):
    """[indices of rank 0, indices of rank 1, ...] — what every rank needs to put gathered results back."""
    order = config_cost_order(cfgs_all)
    return [np.sort(order[r::world]) for r in range(world)]


def scores_in_config_order(util):
    """util[n_cands, n_cfgs] -> per-candidate score, added config by config like src/evaluator.cc:96."""
    out = np.zeros(util.shape[0], dtype=np.float64)
    for k in range(util.shape[1]):
        out += util[:, k]
    return out


def gather_utilities(util_local, n_cfgs_all, rank, world, dist, layout=None):
    """All-gather the [n_cands, n_local_cfgs] utilities of every rank (torch tensor on the rank's device; NCCL
    on GPUs, gloo on CPU) and return the full [n_cands, n_cfgs_all] matrix in the batch's config order.
    `layout` = shard_layout(...) (by default plain striding rank, rank + world, ...)."""
    import torch
    n_cands = util_local.shape[0]
    if layout is None:
        layout = [config_shard(n_cfgs_all, r, world) for r in range(world)]
    widest = max(len(ix) for ix in layout)
    pad = torch.zeros((n_cands, widest), dtype=torch.float64, device=util_local.device)
    pad[:, :util_local.shape[1]] = util_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    full = torch.empty((n_cands, n_cfgs_all), dtype=torch.float64, device=util_local.device)
    for r in range(world):
        ix = torch.as_tensor(layout[r], device=util_local.device)
        full[:, ix] = parts[r][:, :len(layout[r])]
    return full


def partial_scores(sims, n_cands, n_local_cfgs):
    """Per-candidate sum of utilities over this rank's configs, in config order."""
    return scores_in_config_order(np.asarray(sims["utility"], dtype=np.float64).reshape(n_cands, n_local_cfgs))


def all_reduce_scores(partial, dist=None, device=None):
    """Sum the per-candidate partial scores over ranks (NCCL on GPU, gloo on CPU): n_cands doubles on the wire,
    but the last bits depend on the number of ranks."""
    import torch
    t = torch.tensor(partial, dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
