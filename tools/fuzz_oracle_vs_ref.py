#!/usr/bin/env python3
"""Randomised check of the oracle against the reference build (oracle/_ref, only where /root/reference
was available to build it): both sender families, random trees / configs / candidates, and the running
medians of trace == true scoring.  usage: fuzz_oracle_vs_ref.py [seconds] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import helpers  # noqa: E402
from remy_b200 import abi, workloads  # noqa: E402
from oracle import bindings as checkers  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
orc, ref = checkers.Oracle(), checkers.Reference()
t_end = time.time() + budget
it = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed0 * 7919 + it)
    fish = bool(rng.integers(0, 3) == 0)
    flows = False
    n = int(rng.integers(3, 12))
    cfgs = helpers.random_configs(rng, n, max_ticks=float(rng.choice([2000, 6000])))
    cfgs["simulation_ticks"] = np.floor(cfgs["simulation_ticks"])
    seeds = rng.integers(0, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32)
    if fish:
        k = int(rng.integers(1, 5))
        cuts = sorted(float(x) for x in rng.choice([0.25, 1.0, 2.0, 5.0, 12.0, 40.0], k - 1, replace=False))
        tree = abi.fin_tree([float(rng.choice([0.01, 0.3, 1.0, 5.0, 12.5])) for _ in range(k)], cuts)
        a = orc.score_fish(tree, cfgs, seeds, want_use_counts=True, threads=8)
        b = ref.score_fish(tree, cfgs, seeds, want_use_counts=True, threads=8)
    else:
        kind = int(rng.integers(0, 3))
        tree = abi.default_tree() if kind == 0 else workloads.load_golden_tree(str(rng.choice(["RemyCC-2014-100x", "RemyCC-2013-delta1"]))) if kind == 1 \
            else helpers.random_tree(rng, n_children=int(rng.integers(2, 6)), depth=2, grid=True)
        cands = helpers.random_candidates(rng, tree, int(rng.integers(1, 4)))
        flows = bool(rng.integers(0, 4) == 0)   # ByteSwitchedSender<Rat>: the reference's own template instantiated for Rat
        if flows:
            cfgs["mean_on_duration"] = rng.choice([-2.0, 0.4, 1.0, 5.0, 80.0, 80.0, 700.0], n)
            cfgs["mean_off_duration"] = rng.choice([2.0, 20.0, 500.0, 500.0, 3000.0], n)
            a = orc.score_flows(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
            b = ref.score_flows(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
        else:
            a = orc.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
            b = ref.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
    try:
        helpers.assert_same_results(a, b, what=f"iteration {it}")
        if not fish and not flows and (a[0]["error"] == 0).all():
            _, counts, med = orc.trace_medians(tree, cfgs, seeds)
            _, rcounts, rmed, _ = ref.trace_split(tree, cfgs, seeds, split=False)
            mask = 0
            for nd in tree.nodes:
                mask |= int(nd["axis_mask"])
            axes = [x for x in range(6) if mask >> x & 1]
            assert (counts == rcounts).all(), "trace counts"
            leaf_axes = np.zeros((tree.n_leaves, 6), dtype=bool)
            for nd in tree.nodes:
                if nd["child_count"] == 0:
                    for x in range(6):
                        leaf_axes[int(nd["leaf_index"]), x] = bool(int(nd["axis_mask"]) >> x & 1)
            assert med[leaf_axes].tobytes() == rmed[leaf_axes].tobytes(), "trace medians"
    except AssertionError as e:
        print("MISMATCH at iteration", it, "seed0", seed0, "fish", fish, "flows", flows, e)
        sys.exit(1)
    it += 1
print(f"oracle == reference build: {it} iterations, seed0 {seed0}")
