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


@pytest.fixture(scope="module")
def host(built):
    from helpers import ROOT
    from remy_b200 import hostlib
    if not os.path.exists(os.path.join(ROOT, "remy_b200", "csrc", "libremy_b200.so")):
        built.build_cuda()
    built.build_host()
    return hostlib.HostLib()


def test_host_fintree_wire_format_and_generation(host):
    """FinTree / Fin (src/fintree.cc, src/fin.cc) on the host: default tree, RemyBuffers.FinTree bytes
    (fields 90-92, 37-38) cross-checked against python-protobuf, flatten() for the C ABI, Fin::next_generation."""
    import pb_schema
    t = host.fintree_default()
    raw = t.serialize()
    FinTreeMsg = pb_schema.classes()["FinTree"]
    msg = FinTreeMsg()
    msg.ParseFromString(raw)
    assert msg.SerializeToString() == raw
    assert list(msg.domain.active_axis) == [4] and msg.domain.upper.rtt_diff == 163840.0
    assert getattr(msg.leaf, "lambda") == 5.0
    flat = t.flatten()
    want = abi.fin_tree([5.0])
    assert flat.nodes.tobytes() == want.nodes.tobytes() and flat.leaves["intersend"][0] == 5.0
    # a three-fin tree built with python-protobuf parses, flattens and re-serialises byte for byte
    m = FinTreeMsg()
    m.domain.CopyFrom(msg.domain)
    for lo, hi, lam in ((0.0, 2.0, 5.0), (2.0, 20.0, 1.0), (20.0, 163840.0, 0.3)):
        c = m.children.add()
        c.domain.CopyFrom(msg.domain)
        c.domain.lower.rtt_diff, c.domain.upper.rtt_diff = lo, hi
        setattr(c.leaf, "lambda", lam)
        c.leaf.domain.CopyFrom(c.domain)
    raw3 = m.SerializeToString()
    t3 = host.fintree_parse(raw3)
    assert t3.serialize() == raw3
    f3 = t3.flatten()
    w3 = abi.fin_tree([5.0, 1.0, 0.3], [2.0, 20.0])
    assert f3.nodes.tobytes() == w3.nodes.tobytes() and f3.leaves["intersend"].tolist() == [5.0, 1.0, 0.3]
    assert "lambda=1.000000" in t3.str()
    # Fin::next_generation: value, then +-0.01 * 4^k up to 3 inside [0.01, 30], rounded down to 1e-4 (src/fin.cc:33-54,64-67)
    got = t.next_generation(0).tolist()
    want = [5.0]
    c = 0.01
    while c <= 3:
        for v in (5.0 + c, 5.0 - c):
            if 0.01 <= v <= 30:
                want.append(v)
        c *= 4
    assert got == [(1.0 / 10000.0) * int(10000 * v) for v in want] and len(got) == 11


@pytest.mark.gpu
def test_gpu_host_fish_evaluator_and_improver(host, device, oracle):
    """Evaluator<FinTree>::score and FinImprover::improve on the GPU path against the oracle and an
    oracle-driven restatement of the improver's control flow (src/breeder.cc:79-150)."""
    from remy_b200 import hostlib, workloads
    inf = float(abi.BUFFER_INFINITE)
    raw = hostlib.encode_config_range(link=(1, 2, 0.5), rtt=(100, 150, 50), nsrc=(1, 9, 4), buf=(50, 50, 0),
                                      off=(0, 0, 0), on=(1e9, 1e9, 0), ticks=100000)
    cfgs, _ = host.expand(raw)
    assert len(cfgs) == 3 * 2 * 3
    flat = abi.fin_tree([5.0, 1.0, 0.3], [2.0, 20.0])
    tree = host.fintree_from_flat(flat)
    seed = 31
    seeds = workloads.seeds_for(seed, len(cfgs))
    score, counts = tree.score(raw, seed)
    sims, _, want_counts = oracle.score_fish(flat, cfgs, seeds, want_use_counts=True, threads=16)
    want = 0.0
    for u in sims["utility"]:
        want += float(u)
    assert abs(score - want) <= 1e-9 * abs(want) and (counts == want_counts[0]).all()

    def score_many(lams, c):
        cands = np.zeros(len(lams), dtype=abi.leaf_override_dtype)
        for i, lam in enumerate(lams):
            cands[i] = (0, lam, 0, 0)
        s, _, _ = oracle.score_fish(flat, c, seeds, cands=cands, want_senders=False, threads=16)
        return [sum(float(u) for u in s["utility"][i * len(c):(i + 1) * len(c)]) for i in range(len(lams))]

    cfgs_fast, _ = host.expand(raw, carefulness=0.1)
    calls, got_score, got_lambda, n_scored, batches = tree.improve_fin(raw, seed, 0)
    lam, best, cache, want_calls = 5.0, want, {}, 0
    while True:
        want_calls += 1
        alts = [lam]
        c = 0.01
        while c <= 3:
            for v in (lam + c, lam - c):
                if 0.01 <= v <= 30:
                    alts.append(v)
            c *= 4
        alts = [(1.0 / 10000.0) * int(10000 * v) for v in alts]
        fast = [cache[a] if a in cache else None for a in alts]
        fresh = [a for a, f in zip(alts, fast) if f is None]
        for a, s in zip(fresh, score_many(fresh, cfgs_fast) if fresh else []):
            fast[alts.index(a)] = s
        fast = [f if f is not None else cache[a] for a, f in zip(alts, fast)]
        lower = min(best * 1.05, best * 0.95)
        srt = sorted(fast, reverse=True)
        cutoff = max(lower, srt[int(np.ceil(len(srt) * (1.0 - (1 - 0.5)))) - 1])
        top = [a for a, f in zip(alts, fast) if f >= cutoff]
        fresh = [a for a in top if a not in cache]
        full = dict(zip(fresh, score_many(fresh, cfgs) if fresh else []))
        before = best
        for a in top:
            s = full[a] if a in full else cache[a]
            if a in full:
                cache[a] = s
            if s > best:
                best, lam = s, a
        if best == before:
            break
    assert calls == want_calls and got_lambda == lam and abs(got_score - best) <= 1e-9 * abs(best)
    assert batches <= 2 * calls


def test_kernel_fish_trace_equals_oracle(built, oracle):
    """trace == true for Fish on the CPU build of the kernel: per sender, in the shuffled order, the ack's
    lookup or the first send's lookup on the all-zero Memory of the reset; one rtt_diff value per record."""
    from helpers import ROOT
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "libremy_emu.so"))
    emu.remy_emu_trace_fish.argtypes = [C.POINTER(abi.FlatTreeC), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint32)]
    rng = np.random.default_rng(9)
    cfgs = random_configs(rng, 24, max_ticks=5000)
    seeds = rng.integers(1, 2**31 - 2, len(cfgs)).astype(np.uint32)
    for tree in (abi.fin_tree([5.0]), abi.fin_tree([5.0, 1.0, 0.3], [2.0, 20.0])):
        n = len(cfgs)
        sims = np.zeros(n, dtype=abi.sim_result_dtype)
        counts = np.zeros(tree.n_leaves, dtype=np.uint32)
        cap = 2_000_000
        log = np.zeros(cap * 2, dtype=np.uint64)
        looks = np.zeros(n, dtype=np.uint64)
        words = C.c_uint32()
        rc = emu.remy_emu_trace_fish(tree.byref(), cfgs.ctypes.data, seeds.ctypes.data, n, sims.ctypes.data, counts.ctypes.data,
                                     log.ctypes.data, cap, looks.ctypes.data, C.byref(words))
        assert rc == 0 and words.value == 2
        off = 0
        for k in range(n):
            _, leaf, mem = oracle.trace_sim(tree, cfgs[k], seeds[k], fish=True)
            assert len(leaf) == looks[k], k
            rec = log[off * 2:(off + len(leaf)) * 2].reshape(-1, 2)
            assert (rec[:, 0] == leaf).all() and rec[:, 1].tobytes() == np.ascontiguousarray(mem[:, 4]).tobytes(), k
            off += len(leaf)
        assert off == counts.sum() > 100000


@pytest.mark.gpu
def test_gpu_fish_trace_and_split(host, device, oracle):
    """remy_b200_score_trace_fish against the oracle's lookup log; Evaluator<FinTree>::score(trace = true) leaves the
    oracle's running medians on every fin; Breeder<FinTree>::apply_best_split cuts the most used fin at its median."""
    from remy_b200 import hostlib, workloads
    rng = np.random.default_rng(19)
    cfgs = random_configs(rng, 32, max_ticks=5000)
    seeds = rng.integers(1, 2**31 - 2, len(cfgs)).astype(np.uint32)
    flat = abi.fin_tree([5.0, 1.0, 0.3], [2.0, 20.0])
    sims, snd, counts, recs = device.score_trace(flat, cfgs, seeds, fish=True)
    plain = device.score_fish(flat, cfgs, seeds, want_use_counts=True)
    assert sims.tobytes() == plain[0].tobytes() and (counts == plain[2]).all()
    for k in range(len(cfgs)):
        _, leaf, mem = oracle.trace_sim(flat, cfgs[k], seeds[k], fish=True)
        gl, gv = recs[k]
        assert len(gl) == len(leaf) and (gl == leaf).all() and gv.tobytes() == np.ascontiguousarray(mem[:, 4:5]).tobytes(), k
    # host: medians after a traced score, then the split
    raw = hostlib.encode_config_range(link=(1, 2, 0.5), rtt=(100, 150, 50), nsrc=(1, 9, 4), buf=(50, 50, 0),
                                      off=(0, 0, 0), on=(1e9, 1e9, 0), ticks=20000)
    hc, _ = host.expand(raw)
    breeder_seed = 77
    eval_seed = (breeder_seed * 16807) % 2147483647
    hs = workloads.seeds_for(eval_seed, len(hc))
    _, ocounts, omed = oracle.trace_medians(flat, hc, hs, fish=True)
    tree = host.fintree_from_flat(flat)
    tree.score_trace(hc, eval_seed)
    assert tree.medians(flat.n_leaves)[:, 4].tobytes() == omed[:, 4].tobytes()
    tree2 = host.fintree_from_flat(flat)
    tree2.apply_best_split(raw, breeder_seed, 0)
    got = tree2.flatten()
    top = int(np.argmax(np.where(ocounts >= ocounts.max(), np.arange(len(ocounts)), -1)))   # most_used: the last of the maxima
    assert got.n_leaves == flat.n_leaves + 1
    lo, hi = [0.0, 2.0, 20.0][top], [2.0, 20.0, 163840.0][top]
    cuts = sorted(set(float(n["lower"][4]) for n in got.nodes) | set(float(n["upper"][4]) for n in got.nodes))
    assert float(omed[top, 4]) in cuts and lo < float(omed[top, 4]) < hi
