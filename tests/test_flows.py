"""Flow-length workloads: Rat senders behind the reference's ByteSwitchedSender (reference src/sendergang.cc:108-138,
256-278; Rat::send's packets_sent_cap, src/rat-templates.cc:22-26) - what `on = -2` / `on = 80` mean in the shipped
config/*.cfg files.  The reference instantiates the template for its RL sender only; the checker build
(oracle/ref_driver.cc) instantiates the reference's own template for Rat, the oracle restates it, the kernel runs it
as a template variant.  All three must agree bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import ROOT, assert_same_results, random_candidates, random_configs
from remy_b200 import abi, workloads


def flow_configs(rng, n, max_ticks):
    """random_configs with mean_on_duration reinterpreted as a mean flow length in packets: tiny, typical (80, the
    figure4 files), large, and unbounded (-2, most shipped files)."""
    cfgs = random_configs(rng, n, max_ticks=max_ticks)
    cfgs["mean_on_duration"] = rng.choice([-2.0, 0.4, 1.0, 5.0, 80.0, 80.0, 700.0], n)
    cfgs["mean_off_duration"] = rng.choice([2.0, 20.0, 500.0, 500.0, 3000.0], n)   # (0 is refused: see test_..._simultaneous_ends)
    return cfgs


@pytest.mark.parametrize("tree_name", ["default", "RemyCC-2014-100x"])
def test_oracle_flows_equals_reference_build(oracle, reference, tree_name):
    rng = np.random.default_rng(31 + len(tree_name))
    tree = workloads.load_golden_tree(tree_name)
    cfgs = flow_configs(rng, 60, 8000)
    seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
    cands = random_candidates(rng, tree, 3)
    a = oracle.score_flows(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
    b = reference.score_flows(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
    assert_same_results(a, b, exact_utility=True, what="flows: oracle vs reference template")
    assert a[0].tobytes() == b[0].tobytes()
    # the workload is not degenerate: flows end and restart, some senders never get to send a whole flow
    assert int(a[0]["packets_sent"].max()) > 1000 and int((a[0]["packets_sent"] == 0).sum()) < len(a[0]) // 2


class EmuFlows:
    def __init__(self):
        self.lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "libremy_emu.so"))

    def score(self, tree, cfgs, seeds, cands=None, want_use_counts=True, link_cap=0):
        return abi._score_with(self.lib.remy_emu_score_batch_flows, "emu", tree, cfgs, seeds, cands, True, want_use_counts,
                               [(C.c_uint32, link_cap)])


@pytest.mark.parametrize("tree_name", ["default", "RemyCC-2014-100x", "RemyCC-2013-delta1"])
def test_kernel_source_flows_equals_oracle(built, oracle, tree_name):
    emu = EmuFlows()
    rng = np.random.default_rng(77 + len(tree_name))
    tree = workloads.load_golden_tree(tree_name)
    cfgs = flow_configs(rng, 80, 8000)
    seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
    cands = random_candidates(rng, tree, 3)
    want = oracle.score_flows(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
    assert_same_results(emu.score(tree, cfgs, seeds, cands=cands), want, exact_utility=True, what="flows: kernel source vs oracle")
    assert_same_results(emu.score(tree, cfgs, seeds, cands=cands, link_cap=16), want, exact_utility=True, what="flows, regrown rings")


def test_kernel_source_flows_simultaneous_ends(built, oracle):
    """No pacing and one-packet flows: several senders finish flows in the same tick, and the idle times to their next
    flows are drawn from the run's PRNG in the shuffled order of that tick."""
    emu = EmuFlows()
    tree = abi.default_tree()
    rows = [dict(link_ppt=1, delay=d, num_senders=n, buffer_size=b, mean_on_duration=on, mean_off_duration=off, simulation_ticks=2000)
            for d in (0, 20) for n in (2, 5, 32) for b in (3, abi.BUFFER_INFINITE) for on in (0.3, 2.0) for off in (0.5, 15.0)]
    cfgs = abi.make_configs(rows)
    seeds = np.arange(len(rows), dtype=np.uint32) + 5
    cands = np.zeros(2, dtype=abi.leaf_override_dtype)
    cands[0] = (0.9, 0.0, 5, 0)
    cands[1] = (1.0, 0.0, 1, 0)
    want = oracle.score_flows(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
    assert_same_results(emu.score(tree, cfgs, seeds, cands=cands), want, exact_utility=True, what="flows: simultaneous ends")
    # mean flow length 0: the drawn length is 0 packets and the reference's assert( new_flow_length > 0 ) fires
    zero = abi.make_configs([dict(link_ppt=1, delay=10, num_senders=2, mean_on_duration=0.0, mean_off_duration=10.0, simulation_ticks=500)])
    got = emu.score(tree, zero, np.array([1], dtype=np.uint32))
    assert int(got[0]["error"][0]) & abi.ERR_ASSERT
    # a mean idle time of 0: a sender whose flow fits its window would start flow after flow in the same instant (the
    # reference never leaves that tick); refused like the other configs whose switch loop cannot end
    stuck = abi.make_configs([dict(link_ppt=1, delay=10, num_senders=2, mean_on_duration=1.0, mean_off_duration=0.0, simulation_ticks=500)])
    got = emu.score(tree, stuck, np.array([1], dtype=np.uint32))
    assert int(got[0]["error"][0]) == abi.ERR_BAD_CONFIG


@pytest.mark.gpu
def test_gpu_flows_equals_oracle(device, oracle):
    rng = np.random.default_rng(9)
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs = flow_configs(rng, 400, 2e4)
    seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
    cands = random_candidates(rng, tree, 4)
    got = device.score_flows(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    want = oracle.score_flows(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    assert_same_results(got, want, exact_utility=False, what="flows on the GPU")


@pytest.mark.gpu
def test_gpu_flows_on_the_shipped_16_sender_configs(device, oracle):
    """BASELINE configs[3] as the files mean it: on = -2 (unbounded flows) / on = 80 packets, off = 500 ms."""
    import json
    rows = json.load(open(os.path.join(ROOT, "tests", "golden", "config4.json")))
    for r in rows:
        r.pop("file")
        r["simulation_ticks"] = min(r["simulation_ticks"], 5e4)
        if r["mean_on_duration"] == 5000.0:
            r["mean_on_duration"] = -2.0   # (the fixture holds the time-switched mapping of on <= 0)
    cfgs = abi.make_configs(rows)
    seeds = workloads.seeds_for(3, len(cfgs))
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    got = device.score_flows(tree, cfgs, seeds, want_use_counts=True)
    want = oracle.score_flows(tree, cfgs, seeds, want_use_counts=True, threads=16)
    assert_same_results(got, want, exact_utility=False, what="config 4 with flow-switched senders")
