"""The configs at the reference's own run lengths, on the GPU against the oracle (through the C ABI):
BASELINE config 2 at its default 1e6 ticks (reference src/configrange.cc:21) with a real tree, where the
infinite buffer holds thousands of packets and the link ring has to be regrown; config 4's files at the run
lengths they carry (6e5 .. 3.6e6 ticks); sender-runner's native 10 x 1e6 ticks
(reference src/sender-runner.cc:57,143-145).  Slow tests: a single simulation is one GPU thread."""
import os

import numpy as np
import pytest

from helpers import assert_same_results
from remy_b200 import abi, workloads

pytestmark = pytest.mark.gpu


def test_config2_at_1e6_ticks_with_ring_regrowth(device, oracle):
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    full = workloads.default_config_range(ticks=1e6)
    pick = [i for i, c in enumerate(full) if c["num_senders"] in (32, 16, 3) and c["link_ppt"] in (1.0, 2.0) and c["delay"] in (100.0, 200.0)]
    cfgs = full[pick]
    assert len(cfgs) == 12 and cfgs["simulation_ticks"].min() == 1e6
    seeds = workloads.seeds_for(77, len(full))[pick]
    want = oracle.score(tree, cfgs, seeds, want_use_counts=True, threads=16)
    # first with rings of 1024 entries: every simulation whose queue grows beyond that is run again with 16x larger rings
    os.environ["REMY_B200_LINK_CAP"] = "1024"
    try:
        b = device.create_batch(tree, cfgs, seeds, want_senders=True, want_use_counts=True)
        b.run()
        rerun = b.rerun()
        got = b.fetch()
        b.destroy()
    finally:
        del os.environ["REMY_B200_LINK_CAP"]
    assert_same_results(got, want, exact_utility=False, what="config 2, 1e6 ticks, regrown rings")
    assert rerun > 0, "no queue outgrew 1024 packets: the case does not exercise the re-run path"
    assert int(want[0]["event_ticks"].max()) > 4_000_000 and int(want[0]["packets_received"].max()) > 1_000_000
    # and with the default rings
    assert_same_results(device.score(tree, cfgs, seeds, want_use_counts=True), want, exact_utility=False, what="config 2, 1e6 ticks")


def test_config4_files_at_their_own_run_lengths(device, oracle):
    """Every distinct (ticks, link rate, lossy?) of the 105 sixteen-sender config files, at the file's own ticks: 6e5 ticks at
    up to 7.5 packets/ms, 1.2e6 ticks, 3.6e6 ticks at 1.5 packets/ms.  Left out: the ten files with 3.6e6 ticks at 7.5
    packets/ms and the six with 6e5 ticks at 10 (27 M / 6 M packet slots: 130 M / 43 M event ticks, minutes for the ONE
    GPU thread that runs a simulation; they run in the same code with the same ring sizes as the cases kept)."""
    tree = workloads.load_golden_tree("RemyCC-2014-100x")
    cfgs = workloads.config4()
    keys, pick = set(), []
    for i, c in enumerate(cfgs):
        if float(c["simulation_ticks"]) * float(c["link_ppt"]) > 5.4e6:
            continue
        k = (float(c["simulation_ticks"]), float(c["link_ppt"]), float(c["stochastic_loss_rate"]) > 0)
        if k not in keys:
            keys.add(k)
            pick.append(i)
    cfgs = cfgs[pick]
    assert cfgs["simulation_ticks"].max() == 3.6e6 and len(cfgs) >= 10
    seeds = workloads.seeds_for(4, len(cfgs))
    got = device.score(tree, cfgs, seeds, want_use_counts=True)
    want = oracle.score(tree, cfgs, seeds, want_use_counts=True, threads=16)
    assert_same_results(got, want, exact_utility=False, what="config 4 at the files' ticks")
    assert int(want[0]["event_ticks"].max()) > 20_000_000


def test_sender_runner_native_run_length(device, oracle):
    """sender-runner scores 10 x 1 000 000 ticks (carefulness 10) on one config: BASELINE config 1 on the default tree."""
    tree = abi.default_tree()
    cfgs = abi.make_configs([dict(link_ppt=2.0, delay=50.0, num_senders=2, buffer_size=10, mean_on_duration=5000.0,
                                  mean_off_duration=500.0, simulation_ticks=1e7)])
    seeds = np.array([2026], dtype=np.uint32)
    got = device.score(tree, cfgs, seeds, want_use_counts=True)
    want = oracle.score(tree, cfgs, seeds, want_use_counts=True)
    assert_same_results(got, want, exact_utility=False, what="sender-runner, 1e7 ticks")
    assert int(want[0]["event_ticks"][0]) > 10_000_000
