"""N > 1 host logic on CPU: two gloo ranks shard the config axis of ONE batch, score their shards (with the
oracle standing in for the GPU), all-gather the per-simulation utilities and add them per candidate in config
order; the result must equal the single-process scores to the last bit.  Covers remy_b200/sharding.py, which
bench.py --workload improve-step / policy-sweep uses with NCCL (one process per GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from remy_b200 import abi, sharding, workloads
from oracle import bindings as checkers  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    tree = workloads.load_golden_tree("RemyCC-2013-delta1")
    cfgs = workloads.default_config_range(ticks=2e3)[::23]
    seeds = workloads.seeds_for(11, len(cfgs))
    cands = workloads.candidate_grid(4, base=(2, 0.9, 1.0), n=6)
    return tree, cfgs, seeds, cands


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree, cfgs, seeds, cands = _inputs()
    layout = sharding.shard_layout(cfgs, world)
    idx = layout[rank]
    sims, _, _ = checkers.Oracle().score(tree, cfgs[idx], seeds[idx], cands=cands, want_senders=False)
    part = sharding.partial_scores(sims, len(cands), len(idx))
    total = sharding.all_reduce_scores(part, dist)
    util = torch.from_numpy(np.ascontiguousarray(sims["utility"].reshape(len(cands), len(idx))))
    full = sharding.gather_utilities(util, len(cfgs), rank, world, dist, layout)
    exact = sharding.scores_in_config_order(full.numpy())
    ev = torch.tensor([float((sims["packets_sent"] + sims["packets_received"]).sum())], dtype=torch.float64)
    dist.all_reduce(ev, op=dist.ReduceOp.SUM)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)       # a per-rank "time": the bench takes the MAX
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([total, ev.numpy(), t.numpy()]))
    np.save(os.path.join(out_dir, f"exact{rank}.npy"), exact)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(built, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    tree, cfgs, seeds, cands = _inputs()
    sims, _, _ = checkers.Oracle().score(tree, cfgs, seeds, cands=cands, want_senders=False, threads=4)
    u = sims["utility"].reshape(len(cands), len(cfgs))
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert r0.tobytes() == r1.tobytes()                        # every rank holds the reduced result
    got, ev, t = r0[:len(cands)], r0[-2], r0[-1]
    # the reduction order differs from the serial config order, so compare to rounding error
    assert np.allclose(got, u.sum(axis=1), rtol=1e-12, atol=1e-12)
    assert ev == float((sims["packets_sent"] + sims["packets_received"]).sum())
    assert t == 2.0
    # the gathered utilities, added in config order, give the single-process scores to the last bit on every rank
    want = sharding.scores_in_config_order(u)
    assert np.load(tmp_path / "exact0.npy").tobytes() == want.tobytes()
    assert np.load(tmp_path / "exact1.npy").tobytes() == want.tobytes()
    # shards are disjoint, cover the config axis, and split the sender counts evenly
    a, b = sharding.shard_layout(cfgs, 2)
    assert sorted(np.concatenate([a, b]).tolist()) == list(range(len(cfgs)))
    assert abs(int(cfgs["num_senders"][a].sum()) - int(cfgs["num_senders"][b].sum())) <= 32
    assert sorted(np.concatenate([sharding.config_shard(7, 0, 2), sharding.config_shard(7, 1, 2)]).tolist()) == list(range(7))
