"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the
C ABI, against the committed golden vectors (made from the reference's own sources), against
the oracle on seeded random inputs, and through size-independent properties at full size.

Bar: every integer observable, the PRNG end state, every event time and the per-sender
accumulators (tick_share_sending, total_delay) are BIT-EXACT; utilities (the only value
that goes through CUDA's log2) agree within 1e-9 relative (helpers.UTILITY_RTOL)."""
import numpy as np
import pytest

from helpers import (INT_FIELDS, UTILITY_RTOL, assert_same_results, case_inputs, check_against_golden, load_vectors,
                     random_candidates, random_configs)
from remy_b200 import abi, workloads

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", load_vectors(), ids=lambda c: c["name"])
def test_gpu_reproduces_golden(device, case):
    tree, cfgs, seeds, cands = case_inputs(case)
    sims, snd, cnt = device.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    check_against_golden(case, sims, snd, cnt, exact_utility=False)


@pytest.mark.parametrize("tree_name", ["default", "RemyCC-2014-100x", "RemyCC-2013-delta1", "RemyCC-2013-delta0.1",
                                       "RemyCC-2013-delta10"])
def test_gpu_equals_oracle_random(device, oracle, tree_name):
    rng = np.random.default_rng(1000 + len(tree_name))
    tree = workloads.load_golden_tree(tree_name)
    cfgs = random_configs(rng, 96, max_ticks=2e4)
    seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
    cands = random_candidates(rng, tree, 4)
    got = device.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    want = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    assert_same_results(got, want, exact_utility=False, what=tree_name)


def test_gpu_simultaneous_sends_and_ring_regrowth(device, oracle):
    tree = abi.default_tree()
    rows = [dict(link_ppt=1, delay=d, num_senders=n, buffer_size=b, mean_on_duration=50, mean_off_duration=o,
                 simulation_ticks=1500)
            for d in (0, 1, 50) for n in (2, 3, 8, 31, 32) for b in (0, 5, abi.BUFFER_INFINITE) for o in (0, 20)]
    cfgs = abi.make_configs(rows)
    seeds = np.arange(len(rows), dtype=np.uint32) + 7
    cands = np.zeros(3, dtype=abi.leaf_override_dtype)
    cands[0] = (0.9, 0.0, 5, 0)
    cands[1] = (0.5, 0.0, 0, 0)
    cands[2] = (1.0, 1.0, 1, 0)
    want = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    got = device.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    assert_same_results(got, want, exact_utility=False, what="burst")


def test_gpu_spread_and_packed_work_lists_agree(device, oracle, monkeypatch):
    """A batch with fewer simulations than resident threads is spread over all warps (capi.cu: only some lanes
    of a warp take work, cost strata interleaved); REMY_B200_NO_SPREAD=1 packs it into full warps like a batch
    that fills the GPU.  Which simulations share a warp must not show in any result."""
    rng = np.random.default_rng(77)
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs = random_configs(rng, 150, max_ticks=1e4)
    seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
    cands = random_candidates(rng, tree, 3)
    want = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    spread = device.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    monkeypatch.setenv("REMY_B200_NO_SPREAD", "1")
    packed = device.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    monkeypatch.delenv("REMY_B200_NO_SPREAD")
    assert_same_results(spread, want, exact_utility=False, what="spread")
    assert_same_results(packed, want, exact_utility=False, what="packed")
    assert_same_results(spread, packed, exact_utility=True, what="spread vs packed")


def test_gpu_coincident_events(device, oracle):
    """Timer sends, departures and deliveries on the same instants (see tests/test_kernel_emu.py)."""
    from test_kernel_emu import coincident_event_batch
    tree, cfgs, seeds, cands = coincident_event_batch()
    a = device.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    b = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    assert_same_results(a, b, exact_utility=False, what="coincident events")


def test_gpu_config4_sweep(device, oracle):
    """BASELINE configs[3]: the 105 16-sender config files (drop-tail buffers of 10..57 packets and 1e6,
    loss 0 / 0.01 / 0.05, 80 ms bursts), three candidate actions each, 1e5 ticks (the files' own run
    lengths go up to 3.6e6 ticks, minutes per simulation for the CPU checker): bit-exact counters, drops
    included (sent - received - in flight)."""
    cfgs = workloads.config4(ticks_cap=1e5)
    seeds = workloads.seeds_for(4, len(cfgs))
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    rng = np.random.default_rng(44)
    cands = random_candidates(rng, tree, 3)
    a = device.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    b = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    assert (a[0]["error"] == 0).all()
    dropped = a[0]["packets_sent"] - a[0]["packets_received"] - a[0]["link_queued"] - a[0]["delay_inflight"]
    assert (dropped >= 0).all() and dropped.sum() > 1e5
    assert_same_results(a, b, exact_utility=False, what="config4")


def test_gpu_link_ring_overflow_is_rerun(device, oracle):
    """An unpaced, ever-growing window on an infinite buffer queues far more packets than
    the first-try ring holds; the library re-runs such simulations with larger rings."""
    tree = abi.default_tree()
    cfgs = abi.make_configs([dict(link_ppt=0.5, delay=20, num_senders=n, simulation_ticks=3000, mean_off_duration=10)
                             for n in (2, 8)])
    seeds = np.array([5, 6], dtype=np.uint32)
    cands = np.zeros(1, dtype=abi.leaf_override_dtype)
    cands[0] = (1.0, 0.0, 40, 0)
    want = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    assert int(want[0]["link_queued"].max()) > 16384
    got = device.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    assert_same_results(got, want, exact_utility=False, what="overflow")


@pytest.mark.parametrize("kind", ["irregular", "grid", "wide", "six-axes"])
def test_gpu_irregular_trees(device, oracle, kind):
    """Cell-table lookups (batch_prep.h::build_cell_table) and the linear fallback against the
    oracle's first-match scan on trees with overlapping children, holes and six axes."""
    from helpers import random_tree
    rng = np.random.default_rng({"irregular": 11, "grid": 12, "wide": 13, "six-axes": 14}[kind])
    for rep in range(4):
        if kind == "grid":
            tree = random_tree(rng, depth=2, grid=True)
        elif kind == "wide":
            tree = random_tree(rng, n_children=40, depth=1)
        elif kind == "six-axes":
            tree = random_tree(rng, n_children=5, depth=2, axes_mask=0b111111)
        else:
            tree = random_tree(rng, n_children=6, depth=2)
        cfgs = abi.make_configs([dict(link_ppt=float(l), delay=float(d), num_senders=int(n), buffer_size=int(b),
                                      mean_on_duration=500.0, mean_off_duration=200.0, simulation_ticks=4000.0)
                                 for l, d, n, b in zip(rng.choice([0.5, 1, 2, 5], 32), rng.choice([10, 50, 150], 32),
                                                       rng.integers(1, 33, 32), rng.choice([5, 50, abi.BUFFER_INFINITE], 32))])
        seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
        got = device.score(tree, cfgs, seeds, want_use_counts=True)
        want = oracle.score(tree, cfgs, seeds, want_use_counts=True, threads=16)
        assert_same_results(got, want, exact_utility=False, what=f"{kind}/{rep}")


def test_gpu_error_words(device):
    nodes = np.zeros(3, dtype=abi.tree_node_dtype)
    leaves = np.zeros(2, dtype=abi.leaf_action_dtype)
    nodes["upper"][:] = 163840.0
    nodes["axis_mask"] = 15
    nodes[0]["child_begin"], nodes[0]["child_count"] = 1, 2
    nodes[1]["upper"][2] = 1.5
    nodes[2]["lower"][2] = 2.0
    nodes[2]["leaf_index"] = 1
    leaves[0] = (1.0, 0.5, 2, 0)
    leaves[1] = (0.8, 1.0, 0, 0)
    cfgs = abi.make_configs([dict(link_ppt=1, delay=50, num_senders=n, simulation_ticks=20000) for n in (1, 4, 16)])
    seeds = np.array([1, 2, 3], dtype=np.uint32)
    sims, _, _ = device.score(abi.FlatTree(nodes, leaves), cfgs, seeds)
    assert (sims["error"] == abi.ERR_NO_CHILD).all()
    assert sims["event_ticks"].tolist() == [301, 301, 331]
    root = nodes[:1].copy()
    root[0]["child_count"] = 0
    root[0]["upper"][2] = 1.2
    sims, _, _ = device.score(abi.FlatTree(root, leaves[:1]), cfgs, seeds)
    assert (sims["error"] == abi.ERR_NO_WHISKER).all()
    bad = abi.make_configs([dict(link_ppt=1, num_senders=0, simulation_ticks=100),
                            dict(link_ppt=1, num_senders=2, simulation_ticks=0),
                            dict(link_ppt=0, num_senders=2, simulation_ticks=10),
                            dict(link_ppt=1, num_senders=33, simulation_ticks=10),
                            dict(link_ppt=1, num_senders=2, simulation_ticks=10, mean_on_duration=0, mean_off_duration=0),
                            dict(link_ppt=1, num_senders=2, simulation_ticks=10, mean_on_duration=-2, mean_off_duration=500),
                            dict(link_ppt=1, num_senders=2, simulation_ticks=100)])
    sims, _, _ = device.score(abi.default_tree(), bad, np.arange(len(bad), dtype=np.uint32))
    assert sims["error"].tolist() == [abi.ERR_BAD_CONFIG] * 6 + [0]


def test_gpu_argument_errors(device):
    import ctypes as C
    tree = abi.default_tree()
    cfgs = abi.make_configs([dict(link_ppt=1, num_senders=2, simulation_ticks=100)])
    seeds = np.array([1], dtype=np.uint32)
    cands = np.zeros(1, dtype=abi.leaf_override_dtype)
    cands[0] = (1.0, 1.0, 1, 7)  # leaf 7 does not exist
    with pytest.raises(RuntimeError, match="does not exist"):
        device.score(tree, cfgs, seeds, cands=cands)
    nodes = tree.nodes.copy()
    nodes[0]["child_begin"], nodes[0]["child_count"] = 0, 1  # a node that is its own child
    with pytest.raises(RuntimeError, match="invalid child range"):
        device.score(abi.FlatTree(nodes, tree.leaves), cfgs, seeds)
    # empty batch is legal and returns nothing
    sims, _, _ = device.score(tree, cfgs[:0], seeds[:0])
    assert len(sims) == 0


def test_gpu_full_size_properties(device, oracle):
    """BASELINE config 2 at full width (800 NetConfigs x 8 evaluator seeds, RemyCC-2014-100x,
    1e5 ticks) — too big for the CPU checker, so checked through properties:
      * determinism: two runs give identical bytes;
      * batch independence: a simulation's result does not depend on what else is in the
        batch, nor on its position (a strided sub-batch reproduces its rows);
      * conservation: sent == received + queued + in flight + dropped, with no drops at all
        on an infinite, loss-free link;
      * a 1 % sample is bit-exact against the oracle."""
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs = np.tile(workloads.default_config_range(ticks=1e5), 8)
    seeds = workloads.seeds_for(2026, len(cfgs))
    a = device.score(tree, cfgs, seeds, want_use_counts=True)
    b = device.score(tree, cfgs, seeds, want_use_counts=True)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes() and (a[2] == b[2]).all()
    sims = a[0]
    assert (sims["error"] == 0).all()
    assert (sims["packets_sent"] == sims["packets_received"] + sims["link_queued"] + sims["delay_inflight"]).all()
    assert (sims["last_tick"] > 1e5).all() and (sims["event_ticks"] > 0).all()
    idx = np.arange(3, len(cfgs), 101)
    sub = device.score(tree, cfgs[idx], seeds[idx])
    assert sub[0].tobytes() == sims[idx].tobytes()
    assert sub[1][:, :].tobytes() == a[1][idx].tobytes()
    pick = np.arange(7, len(cfgs), 97)
    want = oracle.score(tree, cfgs[pick], seeds[pick], threads=16)
    for f in INT_FIELDS:
        assert (sims[pick][f] == want[0][f]).all(), f
    assert a[1][pick].tobytes() == want[1].tobytes()
    assert (np.abs(sims[pick]["utility"] - want[0]["utility"]) <= UTILITY_RTOL * np.abs(want[0]["utility"])).all()


def test_gpu_random_sweep_sample(device, oracle):
    """BASELINE config 5 shape at reduced size (16 384 randomized NetConfigs x 2 candidates instead
    of 1 M): the whole batch runs on the GPU, a 1 % sample is checked bit-exactly against the oracle,
    and the whole batch through conservation (no packet is created or lost unaccounted)."""
    tree = workloads.load_golden_tree("RemyCC-2013-delta0.1")
    cfgs = workloads.random_sweep(16384, seed=2026, ticks=1e4)
    seeds = workloads.seeds_for(5, len(cfgs))
    cands = np.zeros(2, dtype=abi.leaf_override_dtype)
    cands[0] = (0, 0, 0, abi.NO_OVERRIDE)
    cands[1] = (0.8, 0.5, 2, 3)
    sims, snd, _ = device.score(tree, cfgs, seeds, cands=cands)
    assert (sims["error"] == 0).all()
    # sent = received + still queued + in flight + dropped (buffer drops and stochastic losses)
    dropped = sims["packets_sent"].astype(np.int64) - sims["packets_received"] - sims["link_queued"] - sims["delay_inflight"]
    assert (dropped >= 0).all()
    k = np.arange(len(sims)) % len(cfgs)
    lossless = (cfgs["stochastic_loss_rate"][k] == 0) & (cfgs["buffer_size"][k] == abi.BUFFER_INFINITE)
    assert (dropped[lossless] == 0).all() and (dropped[~lossless] > 0).any()
    pick = np.arange(11, len(cfgs), 100)
    want = oracle.score(tree, cfgs[pick], seeds[pick], cands=cands, threads=16)
    rows = np.concatenate([pick, pick + len(cfgs)])
    assert_same_results((sims[rows], snd[rows], None), want, exact_utility=False, what="sweep sample")


def test_gpu_staged_api_matches_one_shot(device):
    tree = workloads.load_golden_tree("RemyCC-2013-delta1")
    cfgs = workloads.default_config_range(ticks=5e3)[::7]
    seeds = workloads.seeds_for(3, len(cfgs))
    cands = workloads.candidate_grid(3, n=5)
    one = device.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    batch = device.create_batch(tree, cfgs, seeds, cands=cands, want_senders=True, want_use_counts=True)
    for _ in range(2):  # a batch can be re-run; outputs are reset
        ms = batch.run(timed=True)
        assert ms > 0 and batch.launches() >= 1
        got = batch.fetch()
        assert got[0].tobytes() == one[0].tobytes() and got[1].tobytes() == one[1].tobytes() and (got[2] == one[2]).all()
    batch.destroy()
