"""ONE batch partitioned over several GPUs inside one process (remy_b200_init_devices): what RatBreeder /
the `remy` CLI use on a multi-GPU box (dev=all), mirroring the reference's fan-out of one improve step over
the cores of a machine (reference src/breeder.cc:53-77).  The results must be byte-identical to the
single-device results, and bit-exact against the oracle.  Needs >= 2 visible devices (gpurun --gpus 2)."""
import os

import numpy as np
import pytest

from helpers import assert_same_results, random_candidates, random_configs
from remy_b200 import abi, workloads

pytestmark = pytest.mark.gpu


def _need_two():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def multi(built):
    n = _need_two()
    built.build_cuda()
    dev = abi.Device(list(range(n)))
    assert dev.n_devices == n
    yield dev
    dev.lib.remy_b200_init(0)   # leave the library bound to one device for the other test modules


def test_partitioned_batch_equals_oracle_and_single_device(multi, oracle):
    rng = np.random.default_rng(2)
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs = random_configs(rng, 150, max_ticks=1.5e4)
    seeds = rng.integers(0, 2 ** 32, size=len(cfgs), dtype=np.uint64).astype(np.uint32)
    cands = random_candidates(rng, tree, 7)          # 1050 simulations: spread over every device
    b = multi.create_batch(tree, cfgs, seeds, cands=cands, want_senders=True, want_use_counts=True)
    b.run()
    assert b.devices() == multi.n_devices
    got = b.fetch()
    b.destroy()
    want = oracle.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    assert_same_results(got, want, exact_utility=False, what="partitioned batch")
    # the same through the one-shot call, and on ONE device: byte-identical per-simulation results
    one_shot = multi.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    single = abi.Device(0)
    ref = single.score(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    abi.Device(list(range(multi.n_devices)))         # bind all devices again for the following tests
    for x, y in ((got, ref), (one_shot, ref)):
        assert x[0].tobytes() == y[0].tobytes()
        assert x[1].tobytes() == y[1].tobytes()
        assert (x[2] == y[2]).all()


def test_partitioned_improve_step_scores_do_not_depend_on_device_count(multi):
    """Per-candidate scores, added in config order like reference src/evaluator.cc:96, from a batch spread
    over all devices and from the same batch on one device."""
    from remy_b200 import sharding
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs = workloads.default_config_range(ticks=1e4)
    seeds = workloads.seeds_for(5, len(cfgs))
    cands = workloads.candidate_grid(3, base=(2, 0.9, 1.0), n=24)
    sims, _, _ = multi.score(tree, cfgs, seeds, cands=cands, want_senders=False)
    many = sharding.scores_in_config_order(sims["utility"].reshape(len(cands), len(cfgs)))
    sims1, _, _ = abi.Device(0).score(tree, cfgs, seeds, cands=cands, want_senders=False)
    one = sharding.scores_in_config_order(sims1["utility"].reshape(len(cands), len(cfgs)))
    abi.Device(list(range(multi.n_devices)))
    assert many.tobytes() == one.tobytes()
    assert sims.tobytes() == sims1.tobytes()


def test_partitioned_fish_batch_and_ring_regrowth(multi, oracle):
    tree = abi.fin_tree([5.0, 1.0, 0.3], [2.0, 20.0])
    rng = np.random.default_rng(4)
    cfgs = random_configs(rng, 300, max_ticks=4e3)
    seeds = rng.integers(1, 2 ** 31 - 2, len(cfgs)).astype(np.uint32)
    cands = np.zeros(3, dtype=abi.leaf_override_dtype)
    cands[0] = (0, 0, 0, abi.NO_OVERRIDE)
    cands[1] = (0, 2.0, 0, 1)
    cands[2] = (0, 0.25, 0, 0)
    got = multi.score_fish(tree, cfgs, seeds, cands=cands, want_use_counts=True)
    want = oracle.score_fish(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=16)
    assert_same_results(got, want, exact_utility=False, what="fish, partitioned")
    # tiny first-try rings: every device re-runs its own overflowed simulations
    os.environ["REMY_B200_LINK_CAP"] = "16"
    try:
        t2 = abi.default_tree()
        c2 = abi.make_configs([dict(link_ppt=1, delay=d, num_senders=n, mean_on_duration=200, mean_off_duration=50, simulation_ticks=3000)
                               for d in (1, 50) for n in range(1, 33)] * 5)
        s2 = np.arange(len(c2), dtype=np.uint32) + 3
        k2 = np.zeros(2, dtype=abi.leaf_override_dtype)
        k2[0] = (0.9, 0.0, 5, 0)
        k2[1] = (1.0, 1.0, 1, 0)
        b = multi.create_batch(t2, c2, s2, cands=k2, want_senders=True, want_use_counts=True)
        b.run()
        assert b.rerun() > 0 and b.devices() == multi.n_devices
        got2 = b.fetch()
        b.destroy()
    finally:
        del os.environ["REMY_B200_LINK_CAP"]
    assert_same_results(got2, oracle.score(t2, c2, s2, cands=k2, want_use_counts=True, threads=16), exact_utility=False, what="regrowth")
