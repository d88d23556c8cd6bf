"""trace == true scoring on the GPU (remy_b200_score_trace) against the oracle, and
Breeder::apply_best_split on the host mirror against the reference's own classes."""
import os

import numpy as np
import pytest

from remy_b200 import abi, hostlib, workloads
from test_trace import canon, small_configs, trees, zero_window_tree

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host(built, device):
    built.build_host()
    return hostlib.HostLib()


@pytest.mark.parametrize("name,tree", trees())
@pytest.mark.parametrize("log_bytes", [0, 200_000])
def test_device_trace_equals_oracle(device, oracle, name, tree, log_bytes):
    """Every use_whisker call of every simulation, in the reference's order, bit for bit; with a
    tiny log budget the configs are traced in several groups and must still arrive in order."""
    cfgs, seeds = small_configs(21, 40)
    if log_bytes:
        os.environ["REMY_B200_TRACE_LOG_BYTES"] = str(log_bytes)
    try:
        sims, snd, counts, recs = device.score_trace(tree, cfgs, seeds)
    finally:
        os.environ.pop("REMY_B200_TRACE_LOG_BYTES", None)
    plain, psnd, pcounts = device.score(tree, cfgs, seeds, want_use_counts=True)
    assert sims.tobytes() == plain.tobytes() and snd.tobytes() == psnd.tobytes() and (counts == pcounts).all()
    assert (sims["error"] == 0).all()
    total = 0
    for k in range(len(cfgs)):
        osim, leaf, mem = oracle.trace_sim(tree, cfgs[k], seeds[k])
        gl, gv = recs[k]
        assert len(gl) == len(leaf), (k, len(gl), len(leaf))
        assert (gl == leaf).all(), k
        assert gv.tobytes() == np.ascontiguousarray(mem[:, :4]).tobytes(), k
        total += len(leaf)
    assert total == counts.sum() > 10000


def test_host_trace_medians_equal_oracle(host, oracle):
    """Evaluator::score(tree, seed, configs, trace = true, ticks) on the host mirror leaves the same
    running medians on every whisker as the oracle's serial restatement."""
    flat = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs, _ = small_configs(22, 24)
    cfgs["simulation_ticks"] = 6000
    seed = 4242
    tree = host.tree_from_flat(flat)
    score = tree.score_trace(cfgs, seed)
    seeds = workloads.seeds_for(seed, len(cfgs))
    sims, counts, med = oracle.trace_medians(flat, cfgs, seeds)
    want = 0.0
    for u in sims["utility"]:
        want += float(u)
    assert abs(score - want) <= 1e-9 * abs(want)
    assert tree.medians(flat.n_leaves)[:, :4].tobytes() == med[:, :4].tobytes()


@pytest.mark.parametrize("tree_name", ["default", "RemyCC-2014-100x"])
def test_apply_best_split_equals_reference(host, reference, tree_name):
    """Breeder::apply_best_split (src/breeder.cc:15-41) through RatBreeder on the GPU == the same
    steps done by the reference's own Evaluator / MemoryRange / WhiskerTree code."""
    flat = workloads.load_golden_tree(tree_name)
    inf = float(abi.BUFFER_INFINITE)
    raw = hostlib.encode_config_range(link=(1, 2, 0.5), rtt=(100, 200, 50), nsrc=(1, 17, 8), buf=(inf, inf, 0),
                                      off=(2000, 2000, 0), on=(2000, 2000, 0), ticks=10000)
    cfgs, _ = host.expand(raw)
    assert len(cfgs) == 27
    breeder_seed = 99
    eval_seed = (breeder_seed * 16807) % 2147483647   # RatBreeder::next_seed(): one draw per Evaluator
    seeds = workloads.seeds_for(eval_seed, len(cfgs))
    _, rcounts, rmed, rtree = reference.trace_split(flat, cfgs, seeds, generation=0, split=True)
    tree = host.tree_from_flat(flat)
    tree.apply_best_split(raw, breeder_seed, 0)
    got = tree.flatten()
    assert got.n_leaves == rtree.n_leaves == flat.n_leaves + 15
    assert canon(got) == canon(rtree)
