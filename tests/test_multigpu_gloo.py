"""N > 1 host logic on CPU: two gloo ranks shard the config axis, score their shards (with the
oracle standing in for the GPU) and all-reduce the per-candidate partial scores; the result
must equal the single-process scores.  Covers remy_b200/sharding.py, which bench.py and a
multi-GPU RatBreeder step use with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from remy_b200 import abi, sharding, workloads


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
    idx = sharding.config_shard(len(cfgs), rank, world)
    sims, _, _ = abi.Oracle().score(tree, cfgs[idx], seeds[idx], cands=cands, want_senders=False)
    part = sharding.partial_scores(sims, len(cands), len(idx))
    total = sharding.all_reduce_scores(part, dist)
    ev = torch.tensor([float((sims["packets_sent"] + sims["packets_received"]).sum())], dtype=torch.float64)
    dist.all_reduce(ev, op=dist.ReduceOp.SUM)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)       # a per-rank "time": the bench takes the MAX
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([total, ev.numpy(), t.numpy()]))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(built, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    tree, cfgs, seeds, cands = _inputs()
    sims, _, _ = abi.Oracle().score(tree, cfgs, seeds, cands=cands, want_senders=False, threads=4)
    u = sims["utility"].reshape(len(cands), len(cfgs))
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert r0.tobytes() == r1.tobytes()                        # every rank holds the reduced result
    got, ev, t = r0[:len(cands)], r0[-2], r0[-1]
    # the reduction order differs from the serial config order, so compare to rounding error
    assert np.allclose(got, u.sum(axis=1), rtol=1e-12, atol=1e-12)
    assert ev == float((sims["packets_sent"] + sims["packets_received"]).sum())
    assert t == 2.0
    # shards are disjoint and cover the config axis
    a, b = sharding.config_shard(len(cfgs), 0, 2), sharding.config_shard(len(cfgs), 1, 2)
    assert sorted(np.concatenate([a, b]).tolist()) == list(range(len(cfgs)))
