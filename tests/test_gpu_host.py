"""GPU tests of the C++ host mirror: Evaluator<WhiskerTree>::score and score_candidates (the
replacement for ActionImprover::evaluate_replacements) against the oracle."""
import numpy as np
import pytest

from remy_b200 import abi, hostlib, workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host(built, device):
    built.build_host()
    return hostlib.HostLib()


def _range(ticks):
    inf = float(abi.BUFFER_INFINITE)
    return hostlib.encode_config_range(link=(1, 2, 0.5), rtt=(100, 200, 50), nsrc=(1, 32, 6), buf=(inf, inf, 0),
                                       off=(5000, 5000, 0), on=(5000, 5000, 0), ticks=ticks)


@pytest.mark.parametrize("tree_name", ["default", "RemyCC-2014-100x"])
def test_evaluator_score_matches_oracle(host, oracle, tree_name):
    flat = workloads.load_golden_tree(tree_name)
    tree = host.tree_from_flat(flat)
    raw = _range(20000)
    cfgs, _ = host.expand(raw)
    assert len(cfgs) == 3 * 3 * 6
    seed = 77
    score, counts, td = tree.score(raw, seed)
    seeds = workloads.seeds_for(seed, len(cfgs))
    sims, snd, want_counts = oracle.score(flat, cfgs, seeds, want_use_counts=True, threads=16)
    want = 0.0
    for u in sims["utility"]:
        want += float(u)                                       # src/evaluator.cc:96
    assert abs(score - want) <= 1e-9 * abs(want)
    assert (counts == want_counts[0]).all()
    k = 0
    for i, c in enumerate(cfgs):
        for s in range(int(c["num_senders"])):
            r = snd[i][s]
            tput = 0.0 if r["tick_share_sending"] == 0 else float(r["packets_received"]) / r["tick_share_sending"]
            dly = 0.0 if r["packets_received"] == 0 else r["total_delay"] / float(r["packets_received"])
            assert td[k][0] == tput and td[k][1] == dly
            k += 1
    assert k == len(td)
    # carefulness scales the run length (src/evaluator.cc:196-201)
    score2, _, _ = tree.score(raw, seed, carefulness=0.5)
    cfgs2, _ = host.expand(raw, carefulness=0.5)
    sims2, _, _ = oracle.score(flat, cfgs2, seeds, want_senders=False, threads=16)
    assert abs(score2 - float(np.sum(sims2["utility"]))) <= 1e-9 * abs(score2)


def test_score_candidates_matches_oracle(host, oracle):
    """One batch call scores every Whisker::next_generation candidate of a leaf (80 for the
    default whisker) over all configs, with common random numbers across candidates."""
    flat = abi.default_tree()
    tree = host.tree_default()
    raw = _range(5000)
    cfgs, _ = host.expand(raw)
    seed = 5
    cands = tree.next_generation(0)
    got = tree.score_next_generation(raw, seed, 0)
    assert len(got) == len(cands) == 80
    sims, _, _ = oracle.score(flat, cfgs, workloads.seeds_for(seed, len(cfgs)), cands=cands, want_senders=False, threads=16)
    want = sims["utility"].reshape(len(cands), len(cfgs))
    for c in range(len(cands)):
        w = 0.0
        for u in want[c]:
            w += float(u)
        assert abs(got[c] - w) <= 1e-9 * abs(w), c
    assert got[0] == tree.score(raw, seed)[0]                  # candidate 0 is the unchanged whisker


def test_single_config_equals_reference(host, reference):
    """sender-runner's shape: a ConfigRange with one point.  The reference's Evaluator runs its
    only Network on PRNG(seed); so does ours (seed_0 = seed), so score and per-sender
    (throughput, delay) must equal the reference's PUBLIC Evaluator<WhiskerTree>::score — called
    here through oracle/_ref, i.e. the reference's own unmodified evaluator.cc."""
    inf = float(abi.BUFFER_INFINITE)
    for tree_name, link, rtt, n, on, off, buf in (("RemyCC-2014-100x", 0.993, 150, 2, 1000, 1000, inf),
                                                  ("RemyCC-2013-delta1", 1.5, 150, 8, 10000, 10000, inf),
                                                  ("default", 2.0, 50, 2, 5000, 500, 10.0)):
        flat = workloads.load_golden_tree(tree_name)
        tree = host.tree_from_flat(flat)
        raw = hostlib.encode_config_range(link=(link, link, 0), rtt=(rtt, rtt, 0), nsrc=(n, n, 0), buf=(buf, buf, 0),
                                          off=(off, off, 0), on=(on, on, 0), ticks=200000)
        cfgs, _ = host.expand(raw)
        assert len(cfgs) == 1
        for seed in (1, 4242, 2147483648 + 17):
            score, counts, td = tree.score(raw, seed)
            sims, snd, ref_counts = reference.score(flat, cfgs, np.array([seed], dtype=np.uint32), want_use_counts=True, mode=1)
            assert abs(score - sims[0]["utility"]) <= 1e-9 * abs(score)
            for s in range(n):   # mode 1 returns (normalised throughput, average delay) per sender
                assert td[s][0] == snd[0][s]["tick_share_sending"] and td[s][1] == snd[0][s]["total_delay"]
            assert (counts == ref_counts[0]).all()


def test_chained_prng_mode_equals_reference_multi_config_score(host, reference):
    """The reference chains ONE PRNG through all configs of a score() call (src/evaluator.cc:84,92-94), so its number for
    a multi-config range depends on the order of the configs.  The batched evaluator gives every config its own seed; the
    compatibility mode (set_chained_prng) runs the configs one after the other, each seeded with the state its predecessor
    left behind, and must reproduce the reference's PUBLIC multi-config score: same per-sender throughputs and delays, same
    use counts, score within 1e-9."""
    flat = workloads.load_golden_tree("RemyCC-2014-100x")
    tree = host.tree_from_flat(flat)
    raw = hostlib.encode_config_range(link=(1.0, 2.0, 0.5), rtt=(100, 150, 50), nsrc=(1, 7, 3), buf=(float(abi.BUFFER_INFINITE),) * 2 + (0,),
                                      off=(2000, 2000, 0), on=(3000, 3000, 0), ticks=30000)
    cfgs, _ = host.expand(raw)
    assert len(cfgs) == 18
    host.set_chained_prng(True)
    try:
        for seed in (7, 31337):
            score, counts, td = tree.score(raw, seed)
            want_score, want_td, want_counts = reference.score_chained(flat, cfgs, seed)
            assert abs(score - want_score) <= 1e-9 * abs(want_score)
            k = 0
            for c, cfg in enumerate(cfgs):
                for s in range(int(cfg["num_senders"])):
                    assert td[k][0] == want_td[c, s, 0] and td[k][1] == want_td[c, s, 1], (c, s)
                    k += 1
            assert (counts == want_counts).all()
    finally:
        host.set_chained_prng(False)
    # and the default (per-config seeds) is a different, order-independent number
    assert tree.score(raw, 7)[0] != score


def _alternatives(value, lo, hi, min_change, max_change, mult, unsigned):
    """OptimizationSetting::alternatives (src/action.hh:62-91)."""
    out = [value]
    c = min_change
    while c <= max_change:
        up, down = value + c, value - c
        if unsigned and down < 0:
            down = down + 2 ** 32                                  # unsigned wrap: never eligible
        if lo <= up <= hi:
            out.append(up)
        if lo <= down <= hi:
            out.append(down)
        c *= mult
    return out


def _next_generation(w):
    """Whisker::next_generation + round (src/whisker.cc:46-95) for action w = (incr, mult, intersend)."""
    rnd = lambda v: (1.0 / 10000.0) * int(10000 * v)
    return [(i, rnd(m), rnd(s)) for i in _alternatives(w[0], 0, 256, 1, 32, 4, True)
            for m in _alternatives(w[1], 0.0, 1.0, 0.01, 0.5, 4, False)
            for s in _alternatives(w[2], 0.25, 3.0, 0.05, 1.0, 4, False)]


def test_whisker_improver_matches_oracle_driven_restatement(host, oracle):
    """ActionImprover::improve / early_bail_out / evaluate_replacements (src/breeder.cc:53-150) on the GPU
    path against a line-by-line Python restatement of the same control flow that scores with the oracle:
    same number of improve() rounds, same chosen action, same final score."""
    import math
    flat = abi.default_tree()
    tree = host.tree_default()
    inf = float(abi.BUFFER_INFINITE)
    raw = hostlib.encode_config_range(link=(1, 2, 1), rtt=(100, 200, 100), nsrc=(1, 9, 4), buf=(inf, inf, 0),
                                      off=(2000, 2000, 0), on=(2000, 2000, 0), ticks=20000)
    seed = 99
    calls, score, action, n_scored, batches = tree.improve_whisker(raw, seed, 0)

    cfg_full, _ = host.expand(raw)
    cfg_fast, _ = host.expand(raw, carefulness=0.1)
    seeds = workloads.seeds_for(seed, len(cfg_full))

    def score_many(cands, cfgs):
        arr = np.zeros(len(cands), dtype=abi.leaf_override_dtype)
        for i, (inc, mul, isend) in enumerate(cands):
            arr[i] = (mul, isend, inc, 0)
        sims, _, _ = oracle.score(flat, cfgs, seeds, cands=arr, want_senders=False, threads=16)
        u = sims["utility"].reshape(len(cands), len(cfgs))
        out = []
        for c in range(len(cands)):
            t = 0.0
            for v in u[c]:
                t += float(v)
            out.append(t)
        return out

    cache = {}

    def evaluate(cands, cfgs):
        fresh = [c for c in cands if c not in cache]
        got = dict(zip(fresh, score_many(fresh, cfgs))) if fresh else {}
        return [(c in got, got[c] if c in got else cache[c]) for c in cands]

    w = (1, 1.0, 3.0)
    to_beat = score_many([w], cfg_full)[0]
    py_calls = 0
    while True:
        reps = _next_generation(w)
        raw_scores = [s for _, s in evaluate(reps, cfg_fast)]
        lower = min(to_beat * 1.05, to_beat * 0.95)
        ranked = sorted(raw_scores, reverse=True)
        cutoff = max(lower, ranked[math.ceil(len(ranked) * 0.5) - 1])
        top = [r for r, s in zip(reps, raw_scores) if s >= cutoff]
        before = to_beat
        for r, (new, s) in zip(top, evaluate(top, cfg_full)):
            if new:
                cache[r] = s
            if s > to_beat:
                to_beat, w = s, r
        py_calls += 1
        if to_beat == before:
            break
    assert calls == py_calls and action == w, (calls, py_calls, action, w)
    assert abs(score - to_beat) <= 1e-9 * abs(to_beat)
    assert batches <= 2 * calls and n_scored >= 80          # two batch calls per improve(), not one thread per candidate


def test_parse_problem_and_evaluate(host, oracle):
    """The Problem -> Answer wire path (src/evaluator.cc:134-146, src/scoring-example.cc): a Problem
    written by Evaluator::DNA is solved on the GPU; the Answer's score and per-sender numbers equal
    what the oracle gives for the Problem's configs, seed and tick count."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import pb_schema
    cls = pb_schema.classes()
    flat = workloads.load_golden_tree("RemyCC-2013-delta1")
    tree = host.tree_from_flat(flat)
    inf = float(abi.BUFFER_INFINITE)
    raw_range = hostlib.encode_config_range(link=(1, 2, 1), rtt=(100, 200, 100), nsrc=(2, 10, 8), buf=(inf, inf, 0),
                                            off=(3000, 3000, 0), on=(3000, 3000, 0), ticks=15000)
    problem = tree.problem_dna(raw_range, 31337)
    answer = host.solve_problem(problem)
    a = cls["Outcome"]()
    a.ParseFromString(answer)
    assert a.SerializeToString() == answer
    cfgs, _ = host.expand(raw_range)
    sims, snd, _ = oracle.score(flat, cfgs, workloads.seeds_for(31337, len(cfgs)), threads=8)
    want = 0.0
    for u in sims["utility"]:
        want += float(u)
    assert abs(a.score - want) <= 1e-9 * abs(want)
    assert len(a.throughputs_delays) == len(cfgs)
    for k, td in enumerate(a.throughputs_delays):
        assert td.config.link_ppt == cfgs[k]["link_ppt"] and td.config.num_senders == cfgs[k]["num_senders"]
        for s, r in enumerate(td.results):
            x = snd[k][s]
            assert r.throughput == (0.0 if x["tick_share_sending"] == 0 else float(x["packets_received"]) / x["tick_share_sending"])
            assert r.delay == (0.0 if x["packets_received"] == 0 else x["total_delay"] / float(x["packets_received"]))
