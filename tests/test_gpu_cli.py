"""sender-runner CLI (reference src/sender-runner.cc) on the GPU: stdout format and numbers.

Mirrors the reference's own CLI regression test tests/verify-2014-10.test (via
tests/verify-2014-100x-results.pm): RemyCC-2014-100x, 2 senders, link 9.93 Mbit/s, RTT 150 ms,
on = off = 1000 ms -> every sender's throughput within 5 % of 6.0 Mbit/s and delay within 5 % of
165 ms — here with a fixed seed (an extension; the reference seeds from time ^ pid), and
additionally pinned digit-for-digit against the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import ROOT
from remy_b200 import abi, hostlib, workloads

pytestmark = pytest.mark.gpu


def test_sender_runner_verify_2014_10(built, device, oracle, tmp_path):
    built.build_host()
    host = hostlib.HostLib()
    flat = workloads.load_golden_tree("RemyCC-2014-100x")
    dna_path = tmp_path / "RemyCC-2014-100x.dna"
    dna_path.write_bytes(host.tree_from_flat(flat).serialize())
    link_mbps = 9.93
    exe = os.path.join(ROOT, "remy_b200", "host", "sender-runner")
    out = subprocess.run([exe, f"if={dna_path}", "nsrc=2", f"link={link_mbps / 10}", "rtt=150", "on=1000", "off=1000",
                          "seed=7", "ticks=100000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    senders = re.findall(r"sender: \[tp=([0-9.eE+-]+), del=([0-9.eE+-]+)\]", out.stdout)
    assert len(senders) == 2
    for tp, dl in senders:   # tests/verify-2014-100x-results.pm:20-37
        raw_tput = float(tp) * 0.75 * link_mbps
        raw_delay = float(dl) * 150
        assert abs(raw_tput - 6.0) <= 0.05 * 6.0 and abs(raw_delay - 165) <= 0.05 * 165, (raw_tput, raw_delay)
    # digit-for-digit against the oracle: carefulness 10 x 100000 ticks, seed_0 = seed
    cfg = abi.make_configs([dict(link_ppt=link_mbps / 10, delay=150, num_senders=2, mean_on_duration=1000,
                                 mean_off_duration=1000, simulation_ticks=1e6)])
    sims, snd, _ = oracle.score(flat, cfg, np.array([7], dtype=np.uint32))
    want = []
    for s in range(2):
        r = snd[0][s]
        want.append(("%f" % (r["packets_received"] / r["tick_share_sending"] / cfg[0]["link_ppt"]),
                     "%f" % (r["total_delay"] / r["packets_received"] / 150.0)))
    assert senders == want
    assert ("score = %f" % sims[0]["utility"]) in out.stdout
    assert "config: mean_on=1000.000000, mean_off=1000.000000, nsrc=2.000000, link_ppt=0.993000, delay=150.000000" in out.stdout
    assert re.search(r"^normalized_score = -?[0-9.]+$", out.stdout, re.M) and "Rules: [{(lo=< sewma=" in out.stdout
    # (the tree was re-serialised from the golden POD arrays, so it carries no config/optimizer
    # sub-messages and print_tree prints nothing; tests/test_host.py covers the real .dna files)
    assert out.stdout.startswith("score = ")


def test_remy_cli_writes_checkpoints(built, device, tmp_path):
    """`remy cf=... of=... opt=bmr runs=1` (src/remy.cc): banner, per-config report, and a checkpoint
    `<of>.0` that is a RemyBuffers.WhiskerTree carrying config, optimizer and configvector, readable by
    a real protobuf implementation and by sender-runner."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import pb_schema
    built.build_host()
    inf = float(abi.BUFFER_INFINITE)
    cfg_path = tmp_path / "range.cfg"
    cfg_path.write_bytes(hostlib.encode_config_range(link=(1, 2, 1), rtt=(100, 100, 0), nsrc=(2, 6, 4), buf=(inf, inf, 0),
                                                     off=(1000, 1000, 0), on=(1000, 1000, 0), ticks=4000))
    exe = os.path.join(ROOT, "remy_b200", "host", "remy")
    of = tmp_path / "cc"
    out = subprocess.run([exe, f"cf={cfg_path}", f"of={of}", "opt=bmr", "runs=1", "seed=3"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Evaluator simulations will run for 4000 ticks" in out.stdout
    assert "Optimizing window increment: 1, window multiple: 1, intersend: 1" in out.stdout
    assert "Optimizing for link packets_per_ms over [1.000000 : 1.000000 : 2.000000]" in out.stdout
    assert "Optimizing for infinitely sized buffers." in out.stdout
    assert re.search(r"^run = 0, score = -?[0-9.]+$", out.stdout, re.M)
    assert out.stdout.count("config: mean_on=1000.000000") == 4 and out.stdout.count("sender: [tp=") == 2 * (2 + 6)
    raw = (tmp_path / "cc.0").read_bytes()
    msg = pb_schema.classes()["WhiskerTree"]()
    msg.ParseFromString(raw)
    assert msg.SerializeToString() == raw
    assert msg.config.simulation_ticks == 4000 and msg.config.num_senders.high == 6
    assert msg.optimizer.intersend.min_value == 0.25 and msg.optimizer.window_increment.max_value == 256
    assert len(msg.configvector.config) == 4
    # one run = optimise the single whisker, then bisect it at the running medians of its four signals
    # (Breeder::apply_best_split) -- unless the careful re-evaluation found a regression and kept the input
    leaves = [msg.leaf] if msg.HasField("leaf") else [c.leaf for c in msg.children]
    assert len(leaves) in (1, 16) and ("Bisecting ===" in out.stdout)
    if len(leaves) == 16:
        assert all(c.HasField("leaf") and c.domain.upper.rtt_ratio > c.domain.lower.rtt_ratio for c in msg.children)
    for leaf in leaves:
        assert 0 <= leaf.window_increment <= 256 and 0 <= leaf.window_multiple <= 1 and 0.25 <= leaf.intersend <= 3
    # the optimiser only ever accepts strict improvements, and the checkpoint feeds straight back into sender-runner
    assert "Regression" not in out.stderr or "Ending search." in out.stderr
    sr = subprocess.run([os.path.join(ROOT, "remy_b200", "host", "sender-runner"), f"if={of}.0", "nsrc=2", "link=1", "rtt=100",
                         "seed=1", "ticks=2000"], capture_output=True, text=True, timeout=300)
    assert sr.returncode == 0 and sr.stdout.startswith("Prior assumptions:\nlink_packets_per_ms {\n  low: 1\n  high: 2\n  incr: 1\n}")


def test_sender_runner_poisson(built, device, oracle, tmp_path):
    """`sender-runner sender=poisson if=<FinTree>` (src/sender-runner.cc:59-68,80-87,138-141): Fish senders over a
    FinTree file, digits pinned against the oracle's Fish."""
    built.build_host()
    host = hostlib.HostLib()
    flat = abi.fin_tree([5.0, 1.0, 0.3], [2.0, 20.0])
    path = tmp_path / "fins.dna"
    path.write_bytes(host.fintree_from_flat(flat).serialize())
    exe = os.path.join(ROOT, "remy_b200", "host", "sender-runner")
    out = subprocess.run([exe, "sender=poisson", f"if={path}", "nsrc=3", "link=1.5", "rtt=120", "on=2000", "off=1000", "buf=40",
                          "seed=11", "ticks=4000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "Running poisson sender" in out.stderr
    cfg = abi.make_configs([dict(link_ppt=1.5, delay=120, num_senders=3, mean_on_duration=2000, mean_off_duration=1000,
                                 buffer_size=40, simulation_ticks=40000.0)])
    sims, snd, _ = oracle.score_fish(flat, cfg, np.array([11], dtype=np.uint32))
    assert ("score = %f" % sims[0]["utility"]) in out.stdout
    senders = re.findall(r"sender: \[tp=([0-9.eE+-]+), del=([0-9.eE+-]+)\]", out.stdout)
    want = []
    for s in range(3):
        r = snd[0][s]
        tp = 0.0 if r["tick_share_sending"] == 0 else r["packets_received"] / r["tick_share_sending"]
        dl = 0.0 if r["packets_received"] == 0 else r["total_delay"] / r["packets_received"]
        want.append(("%f" % (tp / 1.5), "%f" % (dl / 120.0)))
    assert senders == want
    assert "Rules: [{(lo=< rttd=0.000000 >, hi=< rttd=2.000000 >)} gen=0 usage=" in out.stdout and "(lambda=5.000000)]" in out.stdout


def test_remy_poisson_cli_writes_checkpoints(built, device, tmp_path):
    """`remy-poisson cf=... of=... runs=1` (src/remy-poisson.cc): banner, per-config report and a checkpoint `<of>.0`
    that is a RemyBuffers.FinTree carrying config, optimizer and configvector, readable by python-protobuf and by
    `sender-runner sender=poisson`."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import pb_schema
    built.build_host()
    cfg_path = tmp_path / "range.cfg"
    cfg_path.write_bytes(hostlib.encode_config_range(link=(1, 2, 1), rtt=(100, 100, 0), nsrc=(2, 4, 2), buf=(40, 40, 0),
                                                     off=(0, 0, 0), on=(1e9, 1e9, 0), ticks=3000))
    exe = os.path.join(ROOT, "remy_b200", "host", "remy-poisson")
    of = tmp_path / "fish"
    out = subprocess.run([exe, f"cf={cfg_path}", f"of={of}", "runs=1", "seed=5"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Evaluator simulations will run for 3000 ticks" in out.stdout
    assert "Optimizing for link packets_per_ms in [1.000000, 2.000000]" in out.stdout
    assert "Optimizing for buffer_size in [40.000000, 40.000000]" in out.stdout
    assert re.search(r"^run = 0, score = -?[0-9.]+$", out.stdout, re.M) and "Alternatives: lambda" in out.stdout
    assert out.stdout.count("config: mean_on=1000000000.000000") == 4 and out.stdout.count("sender: [tp=") == 2 * (2 + 4)
    raw = (tmp_path / "fish.0").read_bytes()
    msg = pb_schema.classes()["FinTree"]()
    msg.ParseFromString(raw)
    assert msg.SerializeToString() == raw
    assert msg.config.simulation_ticks == 3000 and msg.config.num_senders.high == 4
    assert getattr(msg.optimizer, "lambda").max_value == 30 and len(msg.configvector.config) == 4
    leaves = [msg.leaf] if msg.HasField("leaf") else [c.leaf for c in msg.children]
    assert len(leaves) in (1, 2) and "Bisecting ===" in out.stdout          # one signal (rtt_diff): a split makes two fins
    for leaf in leaves:
        assert 0.01 <= getattr(leaf, "lambda") <= 30
    sr = subprocess.run([os.path.join(ROOT, "remy_b200", "host", "sender-runner"), "sender=poisson", f"if={of}.0", "nsrc=2", "link=1",
                         "rtt=100", "seed=1", "ticks=1000"], capture_output=True, text=True, timeout=300)
    assert sr.returncode == 0 and sr.stdout.startswith("Prior assumptions:\nlink_packets_per_ms {\n  low: 1\n  high: 2\n  incr: 1\n}")
    assert "lambda {" in sr.stdout and "score = " in sr.stdout
