"""Synthetic workloads of the shapes BASELINE.json names (host logic, numpy only)."""
import numpy as np

from . import abi


def config_range_points(low, high, incr):
    """The reference's `for (x = low; x <= high; x += incr)` loops with the isOne()
    early break (src/evaluator.cc:16-37, src/configrange.hh:30-33)."""
    out = []
    x = low
    while x <= high:
        out.append(x)
        if low == high or incr == 0:
            break
        x += incr
    return out


def default_config_range(ticks=1e6, link=(1.0, 2.0, 0.25), rtt=(100.0, 200.0, 25.0), nsrc=(1, 32, 1),
                         on=5000.0, off=5000.0, buffer_size=abi.BUFFER_INFINITE, loss=0.0):
    """BASELINE config 2: link 1..2 ppt step 0.25, rtt 100..200 step 25,
    nsrc 1..32, on = off = 5000 ms, infinite buffer, no loss -> 800 NetConfigs, in the
    nesting order of Evaluator::Evaluator (link outermost, then rtt, then senders)."""
    rows = []
    for l in config_range_points(*link):
        for r in config_range_points(*rtt):
            for n in config_range_points(*nsrc):
                rows.append(dict(link_ppt=l, delay=r, num_senders=int(n), mean_on_duration=on,
                                 mean_off_duration=off, buffer_size=buffer_size,
                                 stochastic_loss_rate=loss, simulation_ticks=ticks))
    return abi.make_configs(rows)


def seeds_for(evaluator_seed, n):
    """Per-config seeds of one Evaluator: config 0 runs on PRNG(evaluator_seed) itself, exactly
    like the reference (src/evaluator.cc:84), so a single-config score is bit-identical to the
    reference's; config k > 0 gets the k-th output of minstd_rand0(evaluator_seed) — the batched
    evaluator's replacement for the one PRNG the reference chains through all configs."""
    m = 2147483647
    x = evaluator_seed % m or 1
    out = np.empty(n, dtype=np.uint32)
    for i in range(n):
        if i == 0:
            out[0] = evaluator_seed & 0xFFFFFFFF
            continue
        x = (x * 16807) % m
        out[i] = x
    return out


def candidate_grid(leaf_index, base=(1, 1.0, 3.0), n=343):
    """<= 343 single-leaf replacements shaped like Whisker::next_generation
    (src/whisker.cc:46-81, src/action.hh:62-91) around `base` = (increment, multiple, intersend)."""
    def alts(v, lo, hi, mn, mx, mul, integer):
        out = [v]
        c = mn
        while c <= mx:
            for s in (1, -1):
                w = v + s * c
                if integer and w < 0:
                    continue
                if lo <= w <= hi:
                    out.append(w)
            c *= mul
        return out
    incs = alts(base[0], 0, 256, 1, 32, 4, True)
    muls = alts(base[1], 0.0, 1.0, 0.01, 0.5, 4, False)
    ints = alts(base[2], 0.25, 3.0, 0.05, 1.0, 4, False)
    cands = []
    for i in incs:
        for m in muls:
            for s in ints:
                cands.append((int(10000 * m) / 10000.0, int(10000 * s) / 10000.0, int(i), leaf_index))
    arr = np.zeros(min(n, len(cands)), dtype=abi.leaf_override_dtype)
    for k in range(len(arr)):
        arr[k] = cands[k]
    return arr


def load_golden_tree(name):
    """Flattened copies of the reference's tests/RemyCC-*.dna fixtures (tests/golden/trees,
    made by tools/make_golden.py); "default" is WhiskerTree()."""
    import os
    if name == "default":
        return abi.default_tree()
    z = np.load(os.path.join(abi.REPO_ROOT, "tests", "golden", "trees", name + ".npz"))
    return abi.FlatTree(z["nodes"], z["leaves"])


def random_sweep(n, seed=2026, ticks=1e5):
    """BASELINE config 5: randomized NetConfigs for a policy-robustness sweep —
    link ~ U[0.5, 10] ppt, delay ~ U[25, 200] ms, nsrc ~ U{1..32}, buffer in {10, 0.1 BDP, 1 BDP,
    infinite} equiprobable, loss in {0, 0.01, 0.05} w.p. {.6, .3, .1}, on = off = 5000 ms.
    (The survey names std::mt19937_64(2026) as the generator; any fixed generator will do for a
    synthetic sweep — numpy's PCG64 with the same seed is used here and recorded with the results.)"""
    rng = np.random.default_rng(seed)
    link = rng.uniform(0.5, 10.0, n)
    delay = rng.uniform(25.0, 200.0, n)
    nsrc = rng.integers(1, 33, n)
    bdp = link * delay
    kind = rng.integers(0, 4, n)
    buf = np.where(kind == 0, 10.0, np.where(kind == 1, np.maximum(1.0, np.floor(0.1 * bdp)),
                                             np.where(kind == 2, np.maximum(1.0, np.floor(bdp)), float(abi.BUFFER_INFINITE))))
    loss = rng.choice([0.0, 0.01, 0.05], size=n, p=[0.6, 0.3, 0.1])
    out = np.zeros(n, dtype=abi.net_config_dtype)
    out["mean_on_duration"] = 5000.0
    out["mean_off_duration"] = 5000.0
    out["link_ppt"] = link
    out["delay"] = delay
    out["stochastic_loss_rate"] = loss
    out["simulation_ticks"] = ticks
    out["num_senders"] = nsrc
    out["buffer_size"] = buf.astype(np.uint32)
    return out


def config4(ticks_cap=None):
    """BASELINE configs[3] (SURVEY.md §8d "config 4"): the 105 nsrc = 16 files of the reference's config/
    directory, decoded by tools/make_config4_fixture.py into tests/golden/config4.json (drop-tail buffers
    of 10..57 packets and 1e6, stochastic loss 0 / 0.01 / 0.05, on = 5000 ms or 80 ms, off = 500 ms,
    1e5..3.6e6 ticks).  ticks_cap bounds the run length for CPU checkers."""
    import json
    import os
    rows = json.load(open(os.path.join(abi.REPO_ROOT, "tests", "golden", "config4.json")))
    for r in rows:
        r.pop("file")
        if ticks_cap:
            r["simulation_ticks"] = min(r["simulation_ticks"], float(ticks_cap))
    return abi.make_configs(rows)
