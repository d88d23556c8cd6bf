"""CPU tests of the oracle: its arithmetic restatements against the real libstdc++/glibc
code, the whole simulator against the reference's own sources (oracle/_ref, where built)
and against the committed golden vectors (always)."""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import assert_same_results, case_inputs, check_against_golden, load_vectors, random_candidates, random_configs
from remy_b200 import abi, workloads


def _olib(oracle):
    lib = oracle.lib
    lib.remy_oracle_exact_log.restype = C.c_double
    lib.remy_oracle_exact_log.argtypes = [C.c_double]
    lib.remy_oracle_exponential.restype = C.c_double
    lib.remy_oracle_exponential.argtypes = [C.POINTER(C.c_uint32), C.c_double]
    lib.remy_oracle_canonical.restype = C.c_double
    lib.remy_oracle_canonical.argtypes = [C.POINTER(C.c_uint32)]
    lib.remy_oracle_shuffle.argtypes = [C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32]
    return lib


def test_exact_log_equals_libm(oracle):
    """remy_b200/csrc/exact_log.h must return glibc's bits (the host's libm picks the FMA
    build on any x86-64 with FMA; the GPU box's CPU has it too)."""
    lib = _olib(oracle)
    rng = np.random.default_rng(7)
    xs = np.concatenate([1.0 - rng.random(200000), rng.random(50000) * 2.0 + 1e-300,
                         np.exp(rng.uniform(-700, 700, 50000)), [1.0, 0.5, 2.0 ** -53, 1 - 2.0 ** -53, 0.9375, 1.0646972656]])
    bad = [x for x in xs if lib.remy_oracle_exact_log(float(x)).hex() != math.log(float(x)).hex()]
    assert not bad, bad[:5]


def test_distributions_equal_libstdcxx(oracle, reference):
    """exponential / bernoulli / shuffle restatements vs the real std:: code (compiled into
    oracle/_ref with the reference's PRNG typedef)."""
    lib = _olib(oracle)
    ref = reference.lib
    ref.remy_ref_exponential.restype = C.c_double
    ref.remy_ref_exponential.argtypes = [C.POINTER(C.c_uint32), C.c_double]
    ref.remy_ref_bernoulli.argtypes = [C.POINTER(C.c_uint32), C.c_double]
    ref.remy_ref_shuffle.argtypes = [C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32]
    rng = np.random.default_rng(3)
    for seed in rng.integers(1, 2147483646, 3000):
        a, b = C.c_uint32(int(seed)), C.c_uint32(int(seed))
        for rate in (1.0 / 5000.0, 1.0 / 3.0, 2.0):
            assert lib.remy_oracle_exponential(C.byref(a), rate).hex() == ref.remy_ref_exponential(C.byref(b), rate).hex()
            assert a.value == b.value
        for p in (0.0, 0.01, 0.5):
            c = C.c_uint32(a.value)
            assert int(lib.remy_oracle_canonical(C.byref(c)) < p) == ref.remy_ref_bernoulli(C.byref(b), p)
            a = C.c_uint32(c.value)
            assert a.value == b.value
    for n in (1, 2, 3, 4, 5, 8, 15, 16, 31, 32, 33, 100):
        for seed in rng.integers(1, 2147483646, 200):
            a, b = C.c_uint32(int(seed)), C.c_uint32(int(seed))
            ia = np.arange(n, dtype=np.uint32)
            ib = np.arange(n, dtype=np.uint32)
            lib.remy_oracle_shuffle(C.byref(a), ia.ctypes.data, n)
            ref.remy_ref_shuffle(C.byref(b), ib.ctypes.data, n)
            assert (ia == ib).all() and a.value == b.value, (n, seed)
    # states whose next draw is rejected by uniform_int (x - 1 >= past): the two largest
    # engine outputs 2147483645 and 2147483646 are rejected for every range used here
    m = 2147483647
    inv = pow(16807, -1, m)
    for top in (2147483646, 2147483645):
        start = (top * inv) % m
        assert (start * 16807) % m == top
        for n in (2, 3, 16, 32):
            a, b = C.c_uint32(start), C.c_uint32(start)
            ia = np.arange(n, dtype=np.uint32)
            ib = np.arange(n, dtype=np.uint32)
            lib.remy_oracle_shuffle(C.byref(a), ia.ctypes.data, n)
            ref.remy_ref_shuffle(C.byref(b), ib.ctypes.data, n)
            assert (ia == ib).all() and a.value == b.value


@pytest.mark.parametrize("case", load_vectors(), ids=lambda c: c["name"])
def test_oracle_reproduces_golden(oracle, case):
    """Golden vectors were produced by the reference's own sources (tools/make_golden.py)."""
    tree, cfgs, seeds, cands = case_inputs(case)
    sims, snd, cnt = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=4)
    check_against_golden(case, sims, snd, cnt, exact_utility=True)


@pytest.mark.parametrize("tree_name", ["default", "RemyCC-2014-100x", "RemyCC-2013-delta1"])
def test_oracle_equals_reference_random(oracle, reference, tree_name):
    rng = np.random.default_rng(hash(tree_name) % 1000)
    tree = workloads.load_golden_tree(tree_name)
    cfgs = random_configs(rng, 24, max_ticks=1e4)
    cfgs = cfgs[cfgs["num_senders"] <= 32]
    seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
    cands = random_candidates(rng, tree, 3)
    a = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=4)
    b = reference.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=4)
    assert_same_results(a, b, exact_utility=True, what=tree_name)


def test_reference_whitebox_loop_equals_public_score(reference):
    """oracle/ref_driver.cc steps Network::run_simulation's loop itself to read counters; its
    utilities must equal the reference's public static Evaluator<WhiskerTree>::score."""
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs = workloads.default_config_range(ticks=2e4)[::97]
    seeds = workloads.seeds_for(5, len(cfgs))
    w, wn, _ = reference.score(tree, cfgs, seeds, threads=4, mode=0)
    p, pn, _ = reference.score(tree, cfgs, seeds, threads=4, mode=1)
    assert w["utility"].tobytes() == p["utility"].tobytes()
    for i, c in enumerate(cfgs):
        for s in range(int(c["num_senders"])):
            share, delay, recv = wn[i][s]["tick_share_sending"], wn[i][s]["total_delay"], int(wn[i][s]["packets_received"])
            tput = 0.0 if share == 0 else recv / share       # utility.hh:30-36
            avg = 0.0 if recv == 0 else delay / recv         # utility.hh:38-44
            assert tput == pn[i][s]["tick_share_sending"] and avg == pn[i][s]["total_delay"]


def test_oracle_threads_do_not_change_results(oracle):
    tree = workloads.load_golden_tree("RemyCC-2013-delta10")
    cfgs = workloads.default_config_range(ticks=3e3)[::41]
    seeds = workloads.seeds_for(9, len(cfgs))
    a = oracle.score(tree, cfgs, seeds, want_use_counts=True, threads=1)
    b = oracle.score(tree, cfgs, seeds, want_use_counts=True, threads=5)
    assert_same_results(a, b)
