#!/usr/bin/env python3
"""Parity spot-check of one or more builds of the CUDA library on the GPU (A/B work): golden vectors and a
random batch against the oracle.  usage: gpu_check.py lib1.so [lib2.so ...]   (no argument: the in-tree library)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import helpers  # noqa: E402
from remy_b200 import abi, workloads  # noqa: E402
from oracle import bindings as checkers  # noqa: E402

libs = sys.argv[1:] or [None]
orc = checkers.Oracle()
for lib in libs:
    dev = abi.Device(0, path=lib)
    bad = 0
    for case in helpers.load_vectors():
        tree, cfgs, seeds, cands = helpers.case_inputs(case)
        for mode in ("batch", "alone"):
            try:
                if mode == "batch":
                    sims, snd, cnt = dev.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
                    helpers.check_against_golden(case, sims, snd, cnt, exact_utility=False)
                else:
                    for k in range(len(cfgs)):
                        a = dev.score(tree, cfgs[k:k + 1], seeds[k:k + 1], cands=cands, want_use_counts=True)
                        b = orc.score(tree, cfgs[k:k + 1], seeds[k:k + 1], cands=cands, want_use_counts=True)
                        helpers.assert_same_results(a, b, exact_utility=False, what=f"{case['name']}[{k}] alone")
            except AssertionError as e:
                bad += 1
                print(f"  {lib}: {case['name']} ({mode}): {str(e)[:300]}")
    rng = np.random.default_rng(5)
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs = helpers.random_configs(rng, 200, max_ticks=2e4)
    seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
    cands = helpers.random_candidates(rng, tree, 4)
    a = dev.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    b = orc.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    try:
        helpers.assert_same_results(a, b, exact_utility=False, what="random")
    except AssertionError as e:
        bad += 1
        print(f"  {lib}: random batch: {str(e)[:300]}")
    print(f"{lib}: {'OK' if not bad else str(bad) + ' MISMATCHES'}", flush=True)
    dev.lib.remy_b200_shutdown()
