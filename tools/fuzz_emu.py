#!/usr/bin/env python3
"""Randomised hunt for divergences between the kernel source (CPU build, tests/emu) and the oracle:
random trees (regular, irregular, bisected, FinTrees), random NetConfigs incl. the reference's edge cases,
random candidates, both sender families and flow-switched Rat senders, with and without the lookup log.  Not part of the test suite
(minutes to hours of CPU); prints the first mismatch with everything needed to replay it.
usage: fuzz_emu.py [seconds] [seed]"""
import ctypes as C
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

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "libremy_emu.so"))
orc = checkers.Oracle()
for name in ("remy_emu_trace", "remy_emu_trace_fish"):
    getattr(emu, name).argtypes = [C.POINTER(abi.FlatTreeC), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint32)]
t_end = time.time() + budget
it, sims_done = 0, 0
while time.time() < t_end:
    rng = np.random.default_rng(seed0 * 1000003 + it)
    fish = bool(rng.integers(0, 3) == 0)
    if fish:
        k = int(rng.integers(1, 6))
        cuts = sorted(float(x) for x in rng.choice([0.25, 1.0, 2.0, 5.0, 12.0, 40.0, 120.0], k - 1, replace=False))
        tree = abi.fin_tree([float(rng.choice([0.01, 0.3, 1.0, 5.0, 12.5, 30.0])) for _ in range(k)], cuts)
    else:
        kind = int(rng.integers(0, 5))
        if kind == 0:
            tree = abi.default_tree()
        elif kind == 1:
            tree = workloads.load_golden_tree(str(rng.choice(["RemyCC-2014-100x", "RemyCC-2013-delta1", "RemyCC-2013-delta0.1", "RemyCC-2013-delta10"])))
        else:
            tree = helpers.random_tree(rng, n_children=int(rng.integers(2, 8)), depth=int(rng.integers(1, 4)), grid=(kind == 2),
                                       axes_mask=int(rng.choice([0b1111, 0b1111, 0b110111, 0b0101])))
    flows = False
    n = int(rng.integers(4, 24))
    cfgs = helpers.random_configs(rng, n, max_ticks=float(rng.choice([2000, 8000, 20000])))
    if rng.integers(0, 4) == 0:   # coincident events: round service times and pacing
        cfgs["link_ppt"] = rng.choice([0.5, 1.0, 2.0, 4.0], n)
        cfgs["delay"] = rng.choice([0.0, 1.0, 10.0, 50.0], n)
    seeds = rng.integers(0, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32)
    ncand = int(rng.integers(1, 5))
    if fish:
        cands = np.zeros(ncand, dtype=abi.leaf_override_dtype)
        cands[0] = (0, 0, 0, abi.NO_OVERRIDE)
        for c in range(1, ncand):
            cands[c] = (0, float(rng.choice([0.01, 0.25, 2.0, 29.99])), 0, int(rng.integers(0, tree.n_leaves)))
        a = abi._score_with(emu.remy_emu_score_batch_fish, "emu", tree, cfgs, seeds, cands, True, True, [(C.c_uint32, int(rng.choice([0, 0, 16])))])
        b = orc.score_fish(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
    elif rng.integers(0, 4) == 0:   # flow-switched Rat senders (ByteSwitchedSender, src/sendergang.cc:108-138,256-278)
        flows = True
        cfgs["mean_on_duration"] = rng.choice([-2.0, 0.4, 1.0, 5.0, 80.0, 80.0, 700.0], n)      # mean flow length in packets
        cfgs["mean_off_duration"] = rng.choice([2.0, 20.0, 500.0, 500.0, 3000.0], n)
        cands = helpers.random_candidates(rng, tree, ncand)
        a = abi._score_with(emu.remy_emu_score_batch_flows, "emu", tree, cfgs, seeds, cands, True, True, [(C.c_uint32, int(rng.choice([0, 0, 16])))])
        b = orc.score_flows(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
    else:
        cands = helpers.random_candidates(rng, tree, ncand)
        if rng.integers(0, 6) == 0:   # negative / overflowing window multiples (trees from files are not confined to [0, 1])
            for c in range(1, ncand):
                cands[c] = (float(rng.choice([-2.0, -1.0, -0.5, 2500.0])), float(rng.choice([0.0, 0.5, 1.0])), int(rng.choice([0, 1, 2, 5])), int(rng.integers(0, tree.n_leaves)))
        elif rng.integers(0, 3) == 0:   # pacing that is a multiple of the service time; zero windows
            for c in range(1, ncand):
                cands[c] = (float(rng.choice([0.0, 0.5, 1.0])), float(rng.choice([0.0, 0.5, 1.0, 2.0])), int(rng.choice([0, 1, 3])), int(rng.integers(0, tree.n_leaves)))
        a = abi._score_with(emu.remy_emu_score_batch, "emu", tree, cfgs, seeds, cands, True, True, [(C.c_uint32, int(rng.choice([0, 0, 16])))])
        b = orc.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
    # the library's rings stop growing at 2^24 entries (REMY_ERR_LINK_OVERFLOW, a documented limit); the oracle's deque does not
    big = (a[0]["error"] == 4) & (b[0]["link_queued"] > (1 << 23))
    if big.any():
        for part in (0, 1):
            if a[part] is not None:
                a[part][big] = b[part][big]
        a[2][:] = b[2]
    try:
        helpers.assert_same_results(a, b, what=f"iteration {it}")
        ok = (a[0]["error"][:n] == 0)
        if ok.all() and not flows and rng.integers(0, 2) == 0:   # the lookup log of the base tree
            sims = np.zeros(n, dtype=abi.sim_result_dtype)
            counts = np.zeros(tree.n_leaves, dtype=np.uint32)
            cap = int(a[2][0].sum()) + 16
            words = C.c_uint32()
            log = np.zeros(cap * 7, dtype=np.uint64)
            looks = np.zeros(n, dtype=np.uint64)
            fn = emu.remy_emu_trace_fish if fish else emu.remy_emu_trace
            rc = fn(tree.byref(), cfgs.ctypes.data, seeds.ctypes.data, n, sims.ctypes.data, counts.ctypes.data, log.ctypes.data, cap,
                    looks.ctypes.data, C.byref(words))
            assert rc == 0, rc
            w = words.value
            mask = 0
            for nd in tree.nodes:
                mask |= int(nd["axis_mask"])
            axes = [x for x in range(6) if mask >> x & 1]
            off = 0
            for k2 in range(n):
                _, leaf, mem = orc.trace_sim(tree, cfgs[k2], seeds[k2], fish=fish)
                assert len(leaf) == looks[k2], ("trace count", k2, len(leaf), looks[k2])
                rec = log[off * w:(off + len(leaf)) * w].reshape(-1, w)
                assert (rec[:, 0] == leaf).all(), ("trace leaf", k2)
                assert rec[:, 1:].tobytes() == np.ascontiguousarray(mem[:, axes]).tobytes(), ("trace values", k2)
                off += len(leaf)
    except AssertionError as e:
        print("MISMATCH at iteration", it, "seed0", seed0, "fish", fish, "flows", flows, e)
        sys.exit(1)
    it += 1
    sims_done += len(a[0])
print(f"fuzz ok: {it} iterations, {sims_done} simulations, seed0 {seed0}")
