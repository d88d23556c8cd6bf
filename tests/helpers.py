"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

from remy_b200 import abi, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INT_FIELDS = ["event_ticks", "packets_sent", "packets_received", "link_queued", "delay_inflight", "prng_state", "error"]
UTILITY_RTOL = 1e-9  # north_star: "utility scores must agree within 1e-9 relative"


def load_vectors():
    with open(os.path.join(ROOT, "tests", "golden", "vectors.json")) as f:
        return json.load(f)


def case_inputs(case):
    tree = workloads.load_golden_tree(case["tree"])
    cfgs = np.zeros(len(case["configs"]), dtype=abi.net_config_dtype)
    for i, c in enumerate(case["configs"]):
        for k, v in c.items():
            cfgs[i][k] = float.fromhex(v) if isinstance(v, str) else v
    seeds = np.array(case["seeds"], dtype=np.uint32)
    cands = None
    if case["cands"]:
        cands = np.zeros(len(case["cands"]), dtype=abi.leaf_override_dtype)
        for i, c in enumerate(case["cands"]):
            cands[i] = (float.fromhex(c[0]), float.fromhex(c[1]), c[2], c[3])
    return tree, cfgs, seeds, cands


def check_against_golden(case, sims, senders, counts, exact_utility):
    """exact_utility: True for the CPU oracle (same libm log2 as the reference build),
    False for the GPU (CUDA log2: utilities within UTILITY_RTOL, everything else bit-exact)."""
    n_cfg = len(case["configs"])
    for i, exp in enumerate(case["sims"]):
        got = sims[i]
        for f in INT_FIELDS[:-1]:
            assert int(got[f]) == exp[f], (case["name"], i, f, int(got[f]), exp[f])
        assert int(got["error"]) == 0
        assert float(got["last_tick"]).hex() == exp["last_tick"], (case["name"], i)
        eu = float.fromhex(exp["utility"])
        if exact_utility:
            assert float(got["utility"]).hex() == exp["utility"], (case["name"], i)
        else:
            assert abs(float(got["utility"]) - eu) <= UTILITY_RTOL * abs(eu), (case["name"], i, got["utility"], eu)
        for s, es in enumerate(case["senders"][i]):
            g = senders[i][s]
            assert float(g["tick_share_sending"]).hex() == es[0], (case["name"], i, s)
            assert float(g["total_delay"]).hex() == es[1], (case["name"], i, s)
            assert int(g["packets_received"]) == es[2] and int(g["packets_sent"]) == es[3], (case["name"], i, s)
    assert [[int(v) for v in row] for row in counts] == case["use_counts"], case["name"]
    assert len(sims) == n_cfg * (len(case["cands"]) if case["cands"] else 1)


def assert_same_results(a, b, exact_utility=True, what=""):
    """a, b: (sims, senders, counts) triples."""
    sa, na, ca = a
    sb, nb, cb = b
    for f in INT_FIELDS:
        assert (sa[f] == sb[f]).all(), (what, f, np.nonzero(sa[f] != sb[f])[0][:5], sa[f][sa[f] != sb[f]][:5], sb[f][sa[f] != sb[f]][:5])
    assert (sa["last_tick"] == sb["last_tick"]).all(), what
    if exact_utility:
        assert sa["utility"].tobytes() == sb["utility"].tobytes(), what
    else:
        ok = np.abs(sa["utility"] - sb["utility"]) <= UTILITY_RTOL * np.abs(sb["utility"])
        assert ok.all(), (what, sa["utility"][~ok][:5], sb["utility"][~ok][:5])
    if na is not None and nb is not None:
        assert na.tobytes() == nb.tobytes(), (what, "per-sender results differ")
    if ca is not None and cb is not None:
        assert (ca == cb).all(), (what, "use counts differ")


def random_configs(rng, n, max_ticks=2e4):
    """Random NetConfigs covering the reference's edge cases (no buffer, off = 0, delay 0,
    heavy loss, 1..32 senders) with run lengths bounded so a CPU checker finishes quickly:
    ticks x link rate stays below ~1e5 packets per simulation."""
    rows = []
    for _ in range(n):
        link = float(rng.choice([0.1, 0.5, 1, 1.5, 2, 4, 7.5, 10, rng.uniform(0.2, 10)]))
        ticks = float(rng.choice([1000, 5000, max_ticks, 12345.678]))
        ticks = min(ticks, 4e4 / link)
        rows.append(dict(
            num_senders=int(rng.integers(1, 33)), link_ppt=link,
            delay=float(rng.choice([0, 1, 25, 30, 50, 100, 150, rng.uniform(1, 200)])),
            buffer_size=int(rng.choice([0, 1, 2, 10, 23, 57, 1000, abi.BUFFER_INFINITE])),
            stochastic_loss_rate=float(rng.choice([0, 0, 0.01, 0.05, 0.5])),
            mean_on_duration=float(rng.choice([5000, 1000, 80, 10, 0.5])),
            mean_off_duration=float(rng.choice([5000, 500, 1000, 0, 3])),
            simulation_ticks=ticks))
    return abi.make_configs(rows)


def random_candidates(rng, tree, n):
    """Single-leaf overrides.  Window growth is kept bounded (multiple < 1, or multiple 1
    with increment <= 1 and a non-zero pacing interval): an unbounded window with zero
    pacing makes one simulated burst cost millions of event ticks on a CPU checker."""
    cands = np.zeros(n, dtype=abi.leaf_override_dtype)
    cands[0] = (0, 0, 0, abi.NO_OVERRIDE)
    for c in range(1, n):
        mult = float(rng.choice([0, 0.5, 0.9, 1.0]))
        if mult == 1.0:
            incr, inter = int(rng.choice([-5, 0, 1])), float(rng.choice([0.25, 1, 3]))
        else:
            incr, inter = int(rng.choice([-5, 0, 1, 2, 32])), float(rng.choice([0, 0.001, 0.25, 1, 3]))
        cands[c] = (mult, inter, incr, int(rng.integers(0, tree.n_leaves)))
    return cands
