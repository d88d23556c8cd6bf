"""Second sender family: Evaluator<FinTree> / Fish (src/evaluator.cc:105-132,
src/fish.cc, src/fish-templates.cc).  Round 1 holds the checker only: the oracle's restatement of the
Fish pinned bit for bit against the reference's own sources.  The CUDA path for it is next round's."""
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
