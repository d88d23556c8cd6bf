"""trace == true scoring on the CPU side: the oracle's restatement of the lookup order and of the
running medians against the reference build, the kernel's restructured loop (CPU build) against
the oracle record for record, and the host mirror of MemoryRange::bisect / WhiskerTree(whisker,
bisect) / replace(src, dst) against the reference's own code (src/memoryrange.cc:8-41,
src/whiskertree.cc:17-30,159-180, src/breeder.cc:15-41)."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import ROOT, random_configs
from remy_b200 import abi, hostlib, workloads


def zero_window_tree():
    """Two whiskers split on rtt_ratio: the upper one halves the window with increment 0, so
    windows decay to 0 and Rat::send's window-0 lookup (src/rat-templates.cc:13-18) fires every tick."""
    nodes = np.zeros(3, dtype=abi.tree_node_dtype)
    hi = [163840.0] * abi.NUM_AXES
    for i, (lo2, hi2) in enumerate([(0.0, 163840.0), (0.0, 1.3), (1.3, 163840.0)]):
        nodes[i]["lower"][:] = 0.0
        nodes[i]["upper"][:] = hi
        nodes[i]["lower"][2], nodes[i]["upper"][2] = lo2, hi2
        nodes[i]["axis_mask"] = 0b1111
    nodes[0]["child_begin"], nodes[0]["child_count"] = 1, 2
    nodes[1]["leaf_index"], nodes[2]["leaf_index"] = 0, 1
    leaves = np.zeros(2, dtype=abi.leaf_action_dtype)
    leaves[0] = (1.0, 0.5, 2, 0)
    leaves[1] = (0.5, 1.0, 0, 0)
    return abi.FlatTree(nodes, leaves)


def trees():
    return [("default", abi.default_tree()), ("RemyCC-2014-100x", workloads.load_golden_tree("RemyCC-2014-100x")),
            ("RemyCC-2013-delta1", workloads.load_golden_tree("RemyCC-2013-delta1")), ("zero-window", zero_window_tree())]


def small_configs(seed, n=10):
    rng = np.random.default_rng(seed)
    cfgs = random_configs(rng, n, max_ticks=8000)
    cfgs["simulation_ticks"] = np.floor(cfgs["simulation_ticks"])  # the reference's score() takes `unsigned int ticks_to_run`
    return cfgs, rng.integers(1, 2**31 - 2, len(cfgs)).astype(np.uint32)


def canon(flat, i=0):
    nd = flat.nodes[i]
    box = (tuple(float(v) for v in nd["lower"]), tuple(float(v) for v in nd["upper"]), int(nd["axis_mask"]))
    if nd["child_count"] == 0:
        l = flat.leaves[int(nd["leaf_index"])]
        return box + ((float(l["window_multiple"]), float(l["intersend"]), int(l["window_increment"])),)
    return box + tuple(canon(flat, int(nd["child_begin"]) + c) for c in range(int(nd["child_count"])))


@pytest.fixture(scope="module")
def host(built):
    if not os.path.exists(os.path.join(ROOT, "remy_b200", "host", "libremy_host.so")):
        built.build_cuda()
        built.build_host()
    return hostlib.HostLib()


@pytest.fixture(scope="module")
def emu(built):
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "libremy_emu.so"))
    lib.remy_emu_trace.argtypes = [C.POINTER(abi.FlatTreeC), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint32)]
    return lib


@pytest.mark.parametrize("name,tree", trees())
def test_oracle_trace_equals_reference(oracle, reference, name, tree):
    """Same lookup order and same estimator => bit-identical medians for every leaf and axis
    (the reference's own MemoryRange::track runs on top of the P^2 stand-in of oracle/gen_pb_stubs.py)."""
    cfgs, seeds = small_configs(11)
    sims, counts, med = oracle.trace_medians(tree, cfgs, seeds)
    assert (sims["error"] == 0).all()
    score, rcounts, rmed, _ = reference.trace_split(tree, cfgs, seeds, split=False)
    assert (counts == rcounts).all() and counts.sum() > 1000
    assert med[:, :4].tobytes() == rmed[:, :4].tobytes()
    want = 0.0
    for u in sims["utility"]:
        want += float(u)
    assert want == score


@pytest.mark.parametrize("name,tree", trees())
def test_kernel_trace_equals_oracle(emu, oracle, name, tree):
    """The kernel source (CPU build) reports every use_whisker call in the reference's order:
    switch-on resets in sender order, then per sender in shuffled order the ack's lookup and the
    window-0 lookup of Rat::send."""
    cfgs, seeds = small_configs(12, 16)
    n = len(cfgs)
    sims = np.zeros(n, dtype=abi.sim_result_dtype)
    counts = np.zeros(tree.n_leaves, dtype=np.uint32)
    cap = 3_000_000
    log = np.zeros(cap * 7, dtype=np.uint64)
    looks = np.zeros(n, dtype=np.uint64)
    words = C.c_uint32()
    rc = emu.remy_emu_trace(tree.byref(), cfgs.ctypes.data, seeds.ctypes.data, n, sims.ctypes.data, counts.ctypes.data,
                            log.ctypes.data, cap, looks.ctypes.data, C.byref(words))
    assert rc == 0 and words.value == 5
    off = 0
    multi = 0
    for k in range(n):
        osim, leaf, mem = oracle.trace_sim(tree, cfgs[k], seeds[k])
        assert osim["error"] == 0 and sims["error"][k] == 0
        assert len(leaf) == looks[k], (k, len(leaf), looks[k])
        rec = log[off * 5:(off + len(leaf)) * 5].reshape(-1, 5)
        assert (rec[:, 0] == leaf).all(), k
        assert rec[:, 1:].tobytes() == np.ascontiguousarray(mem[:, :4]).tobytes(), k
        off += len(leaf)
        multi += int(cfgs["num_senders"][k] > 1)
    assert off == counts.sum() and multi >= 4


def test_p2_median_restatements_agree(host, oracle):
    """Host (product) and oracle P^2 agree bit for bit, also below five samples and on ties, and
    track the true median of a long stream."""
    L = oracle.lib
    L.remy_oracle_p2_sizeof.restype = C.c_size_t
    L.remy_oracle_p2_median.restype = C.c_double
    L.remy_oracle_p2_median.argtypes = [C.c_void_p]
    L.remy_oracle_p2_init.argtypes = [C.c_void_p]
    L.remy_oracle_p2_add.argtypes = [C.c_void_p, C.c_double]
    rng = np.random.default_rng(3)
    streams = [rng.exponential(1.0, 20000), rng.normal(5, 2, 5000), np.round(rng.uniform(0, 3, 3000)), np.zeros(100),
               np.array([3.0, 1.0, 2.0]), np.array([]), np.array([5.0, 4.0, 3.0, 2.0, 1.0]), rng.uniform(0, 1, 6)]
    for x in streams:
        acc = np.zeros(L.remy_oracle_p2_sizeof(), dtype=np.uint8)
        L.remy_oracle_p2_init(acc.ctypes.data)
        for v in x:
            L.remy_oracle_p2_add(acc.ctypes.data, float(v))
        want = L.remy_oracle_p2_median(acc.ctypes.data)
        got = host.p2_median(x)
        assert np.float64(got).tobytes() == np.float64(want).tobytes()
        if len(x) >= 5000:
            assert abs(got - np.median(x)) < 0.05 * (np.std(x) + 1e-9)


def test_p2_median_reproduces_the_published_worked_example(host, oracle):
    """Known-answer test from the paper that defines the estimator (Jain & Chlamtac, "The P^2 algorithm for
    dynamic calculation of quantiles and histograms without storing observations", CACM 28(10), 1985, Table I):
    twenty observations, p = 0.5, final marker heights 0.02, 0.49, 4.44, 17.20, 38.62 - the paper's estimate
    of the median is 4.44.  (The per-observation values of the median marker below are this restatement's,
    rounded like the table's column; they are kept as a regression record.)  Boost's p_square_quantile_impl
    (what src/memoryrange.hh:24-26 instantiates) is that algorithm; Boost itself is not in the image, so this
    pins the restatements to the published vector, not to Boost's code."""
    obs = [0.02, 0.15, 0.74, 3.39, 0.83, 22.37, 10.15, 15.43, 38.62, 15.92,
           34.60, 10.28, 1.47, 0.40, 0.05, 11.39, 0.27, 0.42, 0.09, 11.37]
    table = {5: 0.74, 6: 0.74, 7: 0.74, 8: 2.18, 9: 4.75, 10: 4.75, 11: 9.27, 12: 9.27, 13: 9.27, 14: 9.27,
             15: 6.30, 16: 6.30, 17: 6.30, 18: 6.30, 19: 4.44, 20: 4.44}
    L = oracle.lib
    L.remy_oracle_p2_sizeof.restype = C.c_size_t
    L.remy_oracle_p2_median.restype = C.c_double
    L.remy_oracle_p2_median.argtypes = [C.c_void_p]
    L.remy_oracle_p2_init.argtypes = [C.c_void_p]
    L.remy_oracle_p2_add.argtypes = [C.c_void_p, C.c_double]
    acc = np.zeros(L.remy_oracle_p2_sizeof(), dtype=np.uint8)
    L.remy_oracle_p2_init(acc.ctypes.data)
    for i, v in enumerate(obs, 1):
        L.remy_oracle_p2_add(acc.ctypes.data, v)
        if i in table:
            assert abs(L.remy_oracle_p2_median(acc.ctypes.data) - table[i]) < 0.006, (i, L.remy_oracle_p2_median(acc.ctypes.data))
            assert abs(host.p2_median(np.array(obs[:i])) - table[i]) < 0.006, i
    heights = acc.view(np.float64)[:5]
    assert np.allclose(heights, [0.02, 0.49, 4.44, 17.20, 38.62], atol=0.006), heights


@pytest.mark.parametrize("name,tree", trees()[:3])
def test_host_split_equals_reference(host, oracle, reference, name, tree):
    """Breeder::apply_best_split's body on the host mirror == the reference's own classes: the tree
    is fed the same lookups (oracle trace, in order), then the most used whisker is bisected."""
    cfgs, seeds = small_configs(13, 6)
    ht = host.tree_from_flat(tree)
    counts = np.zeros(tree.n_leaves, dtype=np.uint32)
    for k in range(len(cfgs)):
        sim, leaf, mem = oracle.trace_sim(tree, cfgs[k], seeds[k])
        assert sim["error"] == 0
        for l in np.unique(leaf):
            ht.track(int(l), mem[leaf == l])
        counts += np.bincount(leaf, minlength=tree.n_leaves).astype(np.uint32)
    score, rcounts, rmed, rtree = reference.trace_split(tree, cfgs, seeds, split=True)
    assert (counts == rcounts).all()
    assert ht.medians(tree.n_leaves)[:, :4].tobytes() == rmed[:, :4].tobytes()
    ht.set_counts(counts)
    ht.split_most_used(0)
    got = ht.flatten()
    assert got.n_leaves == rtree.n_leaves > tree.n_leaves
    assert canon(got) == canon(rtree)


def test_host_sink_feeds_in_parallel_like_serial(host, oracle):
    """The host sink of trace scoring splits the (leaf, axis) estimators over several threads for big
    chunks; the medians must equal the oracle's strictly serial feed bit for bit."""
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs, seeds = small_configs(14, 12)
    cfgs["simulation_ticks"] = 20000
    _, counts, med = oracle.trace_medians(tree, cfgs, seeds)
    ht = host.tree_from_flat(tree)
    recs = []
    for k in range(len(cfgs)):
        _, leaf, mem = oracle.trace_sim(tree, cfgs[k], seeds[k])
        r = np.zeros((len(leaf), 5), dtype=np.uint64)
        r[:, 0] = leaf
        r[:, 1:] = np.ascontiguousarray(mem[:, :4]).view(np.uint64)
        recs.append(r)
    allrec = np.concatenate(recs)
    assert len(allrec) > 200000                     # well above the serial threshold
    half = len(allrec) // 2
    ht.feed_records(allrec[:half])                  # two chunks, like a log that spans staging buffers
    ht.feed_records(allrec[half:])
    assert ht.medians(tree.n_leaves)[:, :4].tobytes() == med[:, :4].tobytes()
