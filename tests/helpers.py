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
    """a, b: (sims, senders, counts) triples.  Error words must be equal.  A simulation that
    ended with a tree-lookup error (where the reference calls exit(1) / fails an assert) has
    no defined results beyond its error word, so the remaining fields are compared on the
    error-free simulations only."""
    sa, na, ca = a
    sb, nb, cb = b
    assert (sa["error"] == sb["error"]).all(), (what, "error", sa["error"], sb["error"])
    ok = sa["error"] == 0
    for f in INT_FIELDS:
        x, y = sa[f][ok], sb[f][ok]
        assert (x == y).all(), (what, f, np.nonzero(x != y)[0][:5], x[x != y][:5], y[x != y][:5])
    assert (sa["last_tick"][ok] == sb["last_tick"][ok]).all(), what
    if exact_utility:
        assert sa["utility"][ok].tobytes() == sb["utility"][ok].tobytes(), what
    else:
        close = np.abs(sa["utility"][ok] - sb["utility"][ok]) <= UTILITY_RTOL * np.abs(sb["utility"][ok])
        assert close.all(), (what, sa["utility"][ok][~close][:5], sb["utility"][ok][~close][:5])
    if na is not None and nb is not None:
        assert na[ok].tobytes() == nb[ok].tobytes(), (what, "per-sender results differ")
    if ca is not None and cb is not None and ok.all():
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


def random_tree(rng, n_children=6, depth=2, grid=False, axes_mask=0b1111):
    """An irregular WhiskerTree: children are random (overlapping, not necessarily covering)
    boxes inside their parent, so the first-match order matters and some lookups fall in
    holes.  Cut points are drawn where the Rat's signals actually live."""
    pools = [np.array([0.0, 0.3, 0.7, 1.0, 1.5, 2.5, 5.0, 20.0, 163840.0]),    # rec_send_ewma
             np.array([0.0, 0.4, 0.6, 1.0, 1.4, 2.0, 4.0, 30.0, 163840.0]),    # rec_rec_ewma
             np.array([0.0, 1.0, 1.05, 1.2, 1.5, 2.0, 3.0, 163840.0]),         # rtt_ratio
             np.array([0.0, 0.2, 0.5, 0.9, 1.3, 3.0, 163840.0]),               # slow_rec_rec_ewma
             np.array([0.0, 5.0, 50.0, 163840.0]), np.array([0.0, 10.0, 100.0, 163840.0])]
    nodes, leaves = [], []

    def new_node(lo, hi):
        nd = np.zeros((), dtype=abi.tree_node_dtype)
        nd["lower"][:] = lo
        nd["upper"][:] = hi
        nd["axis_mask"] = axes_mask
        nodes.append(nd)
        return len(nodes) - 1

    def fill(idx, lo, hi, d):
        if d == 0:
            nodes[idx]["leaf_index"] = len(leaves)
            leaves.append((float(rng.choice([0.5, 0.8, 0.95, 1.0])), float(rng.choice([0.1, 0.5, 1.0, 3.0])),
                           int(rng.choice([1, 1, 2, 5])), 0))
            return
        boxes = []
        if grid:  # bisect every active axis at a pool value strictly inside the box
            mids = []
            for a in range(abi.NUM_AXES):
                inside = [v for v in pools[a] if lo[a] < v < hi[a]]
                mids.append(float(rng.choice(inside)) if inside and (axes_mask >> a) & 1 else None)
            combos = [[]]
            for a in range(abi.NUM_AXES):
                if mids[a] is None:
                    combos = [c + [(lo[a], hi[a])] for c in combos]
                else:
                    combos = [c + [iv] for c in combos for iv in ((lo[a], mids[a]), (mids[a], hi[a]))]
            for c in combos:
                boxes.append(([iv[0] for iv in c], [iv[1] for iv in c]))
        else:
            for _ in range(n_children):
                blo, bhi = list(lo), list(hi)
                for a in range(abi.NUM_AXES):
                    if not (axes_mask >> a) & 1 or rng.random() < 0.4:
                        continue
                    inside = [v for v in pools[a] if lo[a] <= v <= hi[a]]
                    x, y = sorted(rng.choice(inside, 2, replace=len(inside) < 2))
                    if x < y:
                        blo[a], bhi[a] = float(x), float(y)
                boxes.append((blo, bhi))
            if rng.random() < 0.7:
                boxes.append((list(lo), list(hi)))  # catch-all last child: fewer holes
        first = len(nodes)
        for blo, bhi in boxes:
            new_node(blo, bhi)
        nodes[idx]["child_begin"], nodes[idx]["child_count"] = first, len(boxes)
        for k, (blo, bhi) in enumerate(boxes):
            fill(first + k, blo, bhi, d - 1 if rng.random() < 0.6 else 0)

    lo0, hi0 = [0.0] * abi.NUM_AXES, [163840.0] * abi.NUM_AXES
    fill(new_node(lo0, hi0), lo0, hi0, depth)
    lv = np.zeros(len(leaves), dtype=abi.leaf_action_dtype)
    for i, l in enumerate(leaves):
        lv[i] = l
    # leaves must be numbered depth-first; renumber by DFS over the BFS-ish node array
    arr = np.array(nodes, dtype=abi.tree_node_dtype)
    order = []

    def dfs(i):
        if arr[i]["child_count"] == 0:
            order.append(i)
        for c in range(int(arr[i]["child_count"])):
            dfs(int(arr[i]["child_begin"]) + c)

    dfs(0)
    new_leaves = np.zeros_like(lv)
    for new_idx, node_i in enumerate(order):
        new_leaves[new_idx] = lv[int(arr[node_i]["leaf_index"])]
        arr[node_i]["leaf_index"] = new_idx
    return abi.FlatTree(arr, new_leaves)
