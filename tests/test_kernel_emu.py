"""The CUDA kernel's event loop is a restructuring, not a translation, of the reference's
loop.  These CPU tests compile the SAME kernel source (remy_b200/csrc/sim_kernel.cuh) with
g++ through tests/emu/host_emu.h and check it against the oracle, so that logic errors are
caught where there is no GPU.  (The GPU parity tests proper are in test_gpu_parity.py.)"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import ROOT, assert_same_results, case_inputs, check_against_golden, load_vectors, random_candidates, random_configs
from remy_b200 import abi, workloads


class Emu(abi._ScoreLib):
    def __init__(self):
        super().__init__(os.path.join(ROOT, "tests", "emu", "libremy_emu.so"), "remy_emu_score_batch", (C.c_uint32,))

    def score(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False, link_cap=0):
        return super().score(tree, cfgs, seeds, cands, want_senders, want_use_counts, extra=(link_cap,))


@pytest.fixture(scope="module")
def emu(built):
    return Emu()


@pytest.mark.parametrize("case", load_vectors(), ids=lambda c: c["name"])
def test_emu_reproduces_golden(emu, case):
    tree, cfgs, seeds, cands = case_inputs(case)
    sims, snd, cnt = emu.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    check_against_golden(case, sims, snd, cnt, exact_utility=True)


@pytest.mark.parametrize("tree_name", ["default", "RemyCC-2014-100x", "RemyCC-2013-delta0.1"])
def test_emu_equals_oracle_random(emu, oracle, tree_name):
    rng = np.random.default_rng(len(tree_name))
    tree = workloads.load_golden_tree(tree_name)
    cfgs = random_configs(rng, 30, max_ticks=1e4)
    seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
    cands = random_candidates(rng, tree, 3)
    a = emu.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    b = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=4)
    assert_same_results(a, b, exact_utility=True, what=tree_name)


def test_emu_simultaneous_sends_and_ring_regrowth(emu, oracle):
    """intersend 0 makes several senders transmit in the same tick (the only case in which
    the shuffled order is observable); a 2-entry ring forces the overflow re-run path."""
    tree = abi.default_tree()
    rows = [dict(link_ppt=1, delay=d, num_senders=n, buffer_size=b, mean_on_duration=50, mean_off_duration=o,
                 simulation_ticks=1500)
            for d in (0, 1, 50) for n in (2, 3, 8, 31, 32) for b in (0, 5, abi.BUFFER_INFINITE) for o in (0, 20)]
    cfgs = abi.make_configs(rows)
    seeds = np.arange(len(rows), dtype=np.uint32) + 7
    cands = np.zeros(3, dtype=abi.leaf_override_dtype)
    cands[0] = (0.9, 0.0, 5, 0)     # bursts: windows up to 50 packets, no pacing
    cands[1] = (0.5, 0.0, 0, 0)     # windows collapse to 0: Rat::send's repeated lookups
    cands[2] = (1.0, 1.0, 1, 0)
    want = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=4)
    assert_same_results(emu.score(tree, cfgs, seeds, cands=cands, want_use_counts=True), want, what="burst")
    assert int(want[0]["link_queued"].max()) > 64
    assert_same_results(emu.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, link_cap=2), want, what="regrow")


def test_emu_error_words(emu, oracle):
    """exit(1)/assert paths of the reference become per-simulation error words."""
    nodes = np.zeros(3, dtype=abi.tree_node_dtype)
    leaves = np.zeros(2, dtype=abi.leaf_action_dtype)
    nodes["upper"][:] = 163840.0
    nodes["axis_mask"] = 15
    nodes[0]["child_begin"], nodes[0]["child_count"] = 1, 2
    nodes[1]["upper"][2] = 1.5
    nodes[2]["lower"][2] = 2.0      # hole: rtt_ratio in [1.5, 2) matches no child
    nodes[2]["leaf_index"] = 1
    leaves[0] = (1.0, 0.5, 2, 0)
    leaves[1] = (0.8, 1.0, 0, 0)
    hole = abi.FlatTree(nodes, leaves)
    cfgs = abi.make_configs([dict(link_ppt=1, delay=50, num_senders=n, simulation_ticks=20000) for n in (1, 4, 16)])
    seeds = np.array([1, 2, 3], dtype=np.uint32)
    a, b = emu.score(hole, cfgs, seeds), oracle.score(hole, cfgs, seeds)
    assert (a[0]["error"] == abi.ERR_NO_CHILD).all() and (b[0]["error"] == abi.ERR_NO_CHILD).all()
    assert (a[0]["event_ticks"] == b[0]["event_ticks"]).all()
    root = nodes[:1].copy()
    root[0]["child_count"] = 0
    root[0]["upper"][2] = 1.2       # root does not cover rtt_ratio >= 1.2: "No whisker found"
    small = abi.FlatTree(root, leaves[:1])
    a, b = emu.score(small, cfgs, seeds), oracle.score(small, cfgs, seeds)
    assert (a[0]["error"] == abi.ERR_NO_WHISKER).all() and (b[0]["error"] == abi.ERR_NO_WHISKER).all()
    assert (a[0]["event_ticks"] == b[0]["event_ticks"]).all()
    bad = abi.make_configs([dict(link_ppt=1, num_senders=0, simulation_ticks=100),
                            dict(link_ppt=1, num_senders=2, simulation_ticks=0),
                            dict(link_ppt=0, num_senders=2, simulation_ticks=10),
                            dict(link_ppt=1, num_senders=33, simulation_ticks=10),
                            # durations with which the on/off switch loop (src/sendergang.cc:95-105) never ends
                            dict(link_ppt=1, num_senders=2, simulation_ticks=10, mean_on_duration=0, mean_off_duration=0),
                            dict(link_ppt=1, num_senders=2, simulation_ticks=10, mean_on_duration=-2, mean_off_duration=500),
                            dict(link_ppt=1, num_senders=2, simulation_ticks=10, mean_on_duration=50, mean_off_duration=float("nan"))])
    a = emu.score(abi.default_tree(), bad, np.arange(len(bad), dtype=np.uint32))
    assert (a[0]["error"] == abi.ERR_BAD_CONFIG).all()


@pytest.mark.parametrize("kind", ["irregular", "grid", "wide", "six-axes"])
def test_emu_irregular_trees(emu, oracle, kind):
    """The kernel answers lookups from per-node cell tables built on the host
    (batch_prep.h::build_cell_table) and falls back to the linear first-match scan for nodes
    whose table would be too large.  Overlapping children, holes and all six axes must give
    the oracle's leaf (or the oracle's error word) every time."""
    from helpers import random_tree
    rng = np.random.default_rng({"irregular": 1, "grid": 2, "wide": 3, "six-axes": 4}[kind])
    for rep in range(4):
        if kind == "grid":
            tree = random_tree(rng, depth=2, grid=True)
        elif kind == "wide":
            tree = random_tree(rng, n_children=40, depth=1)   # > 4096 cells: linear fallback
        elif kind == "six-axes":
            tree = random_tree(rng, n_children=5, depth=2, axes_mask=0b111111)
        else:
            tree = random_tree(rng, n_children=6, depth=2)
        cfgs = abi.make_configs([dict(link_ppt=float(l), delay=float(d), num_senders=int(n), buffer_size=int(b),
                                      mean_on_duration=500.0, mean_off_duration=200.0, simulation_ticks=4000.0)
                                 for l, d, n, b in zip(rng.choice([0.5, 1, 2, 5], 12), rng.choice([10, 50, 150], 12),
                                                       rng.integers(1, 17, 12), rng.choice([5, 50, abi.BUFFER_INFINITE], 12))])
        seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
        a = emu.score(tree, cfgs, seeds, want_use_counts=True)
        b = oracle.score(tree, cfgs, seeds, want_use_counts=True, threads=4)
        assert_same_results(a, b, what=f"{kind}/{rep}")


def coincident_event_batch():
    """Pacing intervals that are multiples of the link's service time, zero delay, zero pacing: timer
    sends, departures and deliveries fall on the same instants, which is where the kernel's fused /
    split passes must still do the reference's ticks one by one."""
    rows = []
    for n in (1, 2, 5, 16, 32):
        for link, delay in ((1.0, 50.0), (2.0, 100.0), (0.5, 0.0), (1.0, 1.0), (4.0, 25.0)):
            for buf in (0, 1, 3, 10, abi.BUFFER_INFINITE):
                rows.append(dict(num_senders=n, link_ppt=link, delay=delay, buffer_size=buf, stochastic_loss_rate=0.0,
                                 mean_on_duration=1000.0, mean_off_duration=200.0, simulation_ticks=1500.0))
    cfgs = abi.make_configs(rows)
    seeds = np.arange(1, len(cfgs) + 1, dtype=np.uint32)
    acts = ((1, 1.0, 1), (1, 0.5, 1), (1, 2.0, 1), (0.9, 1.0, 2), (1, 0.25, 1), (0.5, 0.0, 4), (1, 3.0, 1), (0.8, 1.0, 32), (1, 0.0, 1))
    cands = np.zeros(len(acts), dtype=abi.leaf_override_dtype)
    for k, (mult, inter, incr) in enumerate(acts):
        cands[k] = (mult, inter, incr, 0)
    return abi.default_tree(), cfgs, seeds, cands


def test_emu_coincident_events(emu, oracle):
    tree, cfgs, seeds, cands = coincident_event_batch()
    a = emu.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    b = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    assert_same_results(a, b, what="coincident events")


def test_emu_config4_sweep(emu, oracle):
    """BASELINE configs[3]: all 16-sender config files (finite buffers, stochastic loss, bursty on-times)."""
    cfgs = workloads.config4(ticks_cap=2500)
    assert len(cfgs) == 105 and set(cfgs["stochastic_loss_rate"].tolist()) == {0.0, 0.01, 0.05}
    seeds = workloads.seeds_for(4, len(cfgs))
    for name in ("default", "RemyCC-2014-100x"):
        tree = workloads.load_golden_tree(name)
        a = emu.score(tree, cfgs, seeds, want_use_counts=True)
        b = oracle.score(tree, cfgs, seeds, want_use_counts=True, threads=16)
        assert (a[0]["error"] == 0).all() and a[0]["packets_sent"].sum() > a[0]["packets_received"].sum() > 100000
        assert_same_results(a, b, what="config4 " + name)


def test_emu_negative_window_multiple(emu, oracle):
    """Trees loaded from files are not confined to the optimiser's 0 <= window_multiple <= 1.  With a
    negative multiple an ack can close the window (int(w m + b) <= 0) that Rat::send's own lookup at
    window 0 reopens to b in the same tick (reference src/rat-templates.cc:13-18)."""
    tree = abi.default_tree()
    rows = [dict(link_ppt=p, delay=d, num_senders=n, buffer_size=b, mean_on_duration=300, mean_off_duration=100,
                 simulation_ticks=3000)
            for p in (0.5, 2) for d in (5, 40) for n in (1, 3, 8) for b in (4, abi.BUFFER_INFINITE)]
    cfgs = abi.make_configs(rows)
    seeds = np.arange(len(rows), dtype=np.uint32) + 99
    cands = np.zeros(5, dtype=abi.leaf_override_dtype)
    cands[0] = (-1.0, 0.5, 2, 0)      # window 2 -> 0 -> (send) 2 -> ...
    cands[1] = (-0.5, 0.0, 3, 0)
    cands[2] = (-2.0, 1.0, 1, 0)
    cands[3] = (-1.0, 0.25, 0, 0)     # stays closed: b = 0
    cands[4] = (3000.0, 0.5, 1, 0)    # w m overflows the int conversion after a few acks
    want = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=4)
    assert_same_results(emu.score(tree, cfgs, seeds, cands=cands, want_use_counts=True), want, what="negative multiple")
    assert int(want[0]["packets_received"][:len(rows)].min()) > 10


def test_small_divisor_division(built):
    """The kernel divides sending-time increments by the number of active senders with a multiply and
    two FMAs (sim_kernel.cuh::div_small_int); the result must be the correctly rounded quotient the
    reference's `/` gives (src/sendergang.cc:169-171)."""
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "libremy_emu.so"))
    lib.remy_emu_check_small_div.restype = C.c_uint64
    lib.remy_emu_check_small_div.argtypes = [C.c_uint64, C.c_uint64]
    assert lib.remy_emu_check_small_div(3_200_000, 1) == 0


def test_shuffle_skip_is_the_draw_by_draw_loop(built):
    """sim_kernel.cuh::shuffle_skip advances the run's PRNG by std::shuffle's draws (bits/stl_algo.h:3742-3805,
    src/sendergang.cc:75-82) on a lazily reduced residue; it must leave exactly the state the uniform_int loop
    leaves, including for the states where a draw is rejected or the lazy reduction wraps."""
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "libremy_emu.so"))
    lib.remy_emu_check_shuffle_skip.restype = C.c_uint64
    lib.remy_emu_check_shuffle_skip.argtypes = [C.c_uint64, C.c_uint64]
    assert lib.remy_emu_check_shuffle_skip(200_000, 3) == 0


def test_emu_full_queues_at_ring_boundaries(emu, oracle):
    """The packet that enters service is read from its ring slot only at the top of the next pass of the event
    loop (sim_kernel.cuh, REMY_TOP_LOADS), so the ring keeps one entry in reserve.  Finite buffers around the
    powers of two, kept FULL by bursting senders, and forced ring sizes from 2 entries up must all give the
    reference's result: no slot is overwritten before it is read."""
    tree = abi.default_tree()
    rows = [dict(link_ppt=p, delay=d, num_senders=n, buffer_size=b, mean_on_duration=200, mean_off_duration=10,
                 simulation_ticks=1200)
            for p in (0.5, 3.0) for d in (0, 20) for n in (4, 16) for b in (1, 2, 3, 4, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64)]
    cfgs = abi.make_configs(rows)
    seeds = np.arange(len(rows), dtype=np.uint32) * 3 + 11
    cands = np.zeros(2, dtype=abi.leaf_override_dtype)
    cands[0] = (1.0, 0.0, 40, 0)    # the window grows by 40 per ack, no pacing: the queue is at its limit all the time
    cands[1] = (0.9, 0.05, 9, 0)
    want = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=4)
    full = want[0]["link_queued"].reshape(2, -1)[0] >= np.minimum(cfgs["buffer_size"], 10 ** 9)
    assert full.mean() > 0.5, "the queues were meant to be full at the end of most runs"
    for cap in (0, 2, 16, 32):
        assert_same_results(emu.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, link_cap=cap), want, what=f"link_cap {cap}")
