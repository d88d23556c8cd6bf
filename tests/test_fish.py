"""Second sender family: Evaluator<FinTree> / Fish (src/evaluator.cc:105-132,
src/fish.cc, src/fish-templates.cc): the oracle's restatement of the Fish pinned bit for bit against the
reference's own sources, the kernel's FISH variant (CPU build) against the oracle, and the GPU path
(remy_b200_score_batch_fish) against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import assert_same_results, random_configs
from remy_b200 import abi


@pytest.mark.parametrize("name,tree", [("default fin", abi.fin_tree([5.0])), ("three fins", abi.fin_tree([5.0, 1.0, 0.3], [2.0, 20.0]))])
def test_oracle_fish_equals_reference(oracle, reference, name, tree):
    rng = np.random.default_rng(1)
    cfgs = random_configs(rng, 32, max_ticks=5000)
    cfgs["simulation_ticks"] = np.floor(cfgs["simulation_ticks"])
    seeds = rng.integers(1, 2**31 - 2, len(cfgs)).astype(np.uint32)
    cands = np.zeros(3, dtype=abi.leaf_override_dtype)
    cands[0] = (0, 0, 0, abi.NO_OVERRIDE)
    cands[1] = (0, 0.25, 0, 0)                       # lambda alternatives (Fin::next_generation)
    cands[2] = (0, 12.5, 0, tree.n_leaves - 1)
    a = oracle.score_fish(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    b = reference.score_fish(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16, mode=0)
    assert (a[0]["error"] == 0).all() and a[0]["packets_sent"].sum() > 10 * a[0]["packets_received"].sum() > 1e6
    assert_same_results(a, b, what=name)
    pub = reference.score_fish(tree, cfgs, seeds, want_use_counts=False, threads=16, mode=1)   # the public static score()
    assert pub[0]["utility"].tobytes() == a[0]["utility"][:len(cfgs)].tobytes()


def fish_batch(seed, n):
    rng = np.random.default_rng(seed)
    cfgs = random_configs(rng, n, max_ticks=6000)
    seeds = rng.integers(1, 2**31 - 2, len(cfgs)).astype(np.uint32)
    trees = [abi.fin_tree([5.0]), abi.fin_tree([5.0, 1.0, 0.3], [2.0, 20.0]), abi.fin_tree([0.05, 30.0, 2.0, 0.7, 9.0], [0.5, 3.0, 10.0, 60.0])]
    out = []
    for tree in trees:
        cands = np.zeros(4, dtype=abi.leaf_override_dtype)
        cands[0] = (0, 0, 0, abi.NO_OVERRIDE)
        cands[1] = (0, 0.25, 0, 0)
        cands[2] = (0, 12.5, 0, tree.n_leaves - 1)
        cands[3] = (0, 0.01, 0, 0)
        out.append((tree, cfgs, seeds, cands))
    return out


def test_kernel_fish_equals_oracle(built, oracle):
    """The kernel source compiled for the host with FISH = true (tests/emu): batches of five packets into
    drop-tail queues, per-sender PRNG draws on every ack batch and send, lambda looked up lazily."""
    from helpers import ROOT
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "libremy_emu.so"))
    for tree, cfgs, seeds, cands in fish_batch(2, 48):
        a = abi._score_with(emu.remy_emu_score_batch_fish, "emu", tree, cfgs, seeds, cands, True, True, [(C.c_uint32, 0)])
        b = oracle.score_fish(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
        assert (a[0]["error"] == 0).all() and a[0]["packets_received"].sum() > 5e5
        assert_same_results(a, b, what="fish (kernel on CPU)")
    # the link ring is regrown when a flood outgrows it
    tree, cfgs, seeds, cands = fish_batch(3, 16)[0]
    cfgs["buffer_size"] = abi.BUFFER_INFINITE
    a = abi._score_with(emu.remy_emu_score_batch_fish, "emu", tree, cfgs, seeds, None, True, True, [(C.c_uint32, 64)])
    b = oracle.score_fish(tree, cfgs, seeds, want_use_counts=True, threads=16)
    assert_same_results(a, b, what="fish regrow")


@pytest.mark.gpu
def test_gpu_fish_equals_oracle(device, oracle):
    """remy_b200_score_batch_fish: bit-exact counters / PRNG end states / sender accumulators, utilities to 1e-9."""
    for tree, cfgs, seeds, cands in fish_batch(5, 96):
        a = device.score_fish(tree, cfgs, seeds, cands=cands, want_use_counts=True)
        b = oracle.score_fish(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
        assert (a[0]["error"] == 0).all()
        assert_same_results(a, b, exact_utility=False, what="fish (GPU)")
