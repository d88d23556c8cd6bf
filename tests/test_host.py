"""C++ host mirror of the reference API (remy_b200/host): wire formats, candidate generation,
ConfigRange expansion.  CPU only; the Evaluator calls that need the GPU are in test_gpu_host.py."""
import glob
import os

import numpy as np
import pytest

from remy_b200 import abi, dna, hostlib, workloads

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(built):
    built.build_cuda()
    built.build_host()
    return hostlib.HostLib()


def test_default_tree(host):
    t = host.tree_default()
    f = t.flatten()
    d = abi.default_tree()
    assert f.nodes.tobytes() == d.nodes.tobytes() and f.leaves.tobytes() == d.leaves.tobytes()
    # WhiskerTree::str of the default tree (src/whisker.cc:83-89 format; 0/0 uses prints as -nan with glibc)
    assert t.str().endswith("gen=0 usage=-nan => (win=1 + 1.000000 * win, intersend=3.000000)]")
    assert host.tree_parse(t.serialize()).flatten().nodes.tobytes() == d.nodes.tobytes()


@pytest.mark.parametrize("name", ["RemyCC-2014-100x", "RemyCC-2013-delta1", "RemyCC-2013-delta0.1", "RemyCC-2013-delta10"])
def test_golden_trees_round_trip(host, name):
    """flat -> WhiskerTree -> DNA bytes -> WhiskerTree -> flat is the identity, and the python
    wire reader (remy_b200/dna.py) agrees with the C++ one on those bytes."""
    flat = workloads.load_golden_tree(name)
    t = host.tree_from_flat(flat)
    raw = t.serialize()
    back = host.tree_parse(raw).flatten()
    assert back.nodes.tobytes() == flat.nodes.tobytes() and back.leaves.tobytes() == flat.leaves.tobytes()
    py = dna.flatten(dna.parse_whisker_tree(raw)[0])
    assert py.nodes.tobytes() == flat.nodes.tobytes() and py.leaves.tobytes() == flat.leaves.tobytes()


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not present")
def test_reference_dna_files_byte_identical(host):
    for path in sorted(glob.glob(os.path.join(REF, "tests", "*.dna"))):
        raw = open(path, "rb").read()
        t = host.tree_parse(raw)
        assert t.serialize() == raw, path                      # incl. config/optimizer sub-messages and wildcards
        want = workloads.load_golden_tree(os.path.basename(path)[:-4])
        got = t.flatten()
        assert got.nodes.tobytes() == want.nodes.tobytes() and got.leaves.tobytes() == want.leaves.tobytes()


def test_next_generation(host):
    """Whisker::next_generation (src/whisker.cc:46-81) on the default whisker (1, 1.0, 3.0):
    increment {1,2,0,5,17} x multiple {1,.99,.96,.84} x intersend {3,2.95,2.8,2.2}."""
    c = host.tree_default().next_generation(0)
    assert len(c) == 80 and (c["leaf_index"] == 0).all()
    assert (c[0]["window_increment"], c[0]["window_multiple"], c[0]["intersend"]) == (1, 1.0, 3.0)
    assert sorted(set(c["window_increment"].tolist())) == [0, 1, 2, 5, 17]
    rnd = lambda v: (1.0 / 10000.0) * int(10000 * v)      # Whisker::round (src/whisker.cc:91-95)
    assert sorted(set(c["window_multiple"].tolist())) == sorted(rnd(v) for v in (1 - 0.16, 1 - 0.04, 1 - 0.01, 1.0))
    assert sorted(set(c["intersend"].tolist())) == sorted(rnd(v) for v in (3 - 0.8, 3 - 0.2, 3 - 0.05, 3.0))
    assert len(host.tree_default().next_generation(0, True, False, False)) == 5
    # a leaf of a real tree: every candidate maps back to that leaf
    t = host.tree_from_flat(workloads.load_golden_tree("RemyCC-2014-100x"))
    flat = t.flatten()
    ok = [i for i in range(flat.n_leaves) if 0 <= flat.leaves[i]["window_increment"] <= 256 and 0.25 <= flat.leaves[i]["intersend"] <= 3]
    g = t.next_generation(ok[0])
    assert 1 <= len(g) <= 343 and (g["leaf_index"] == ok[0]).all()
    # the reference aborts on values outside the optimizer's range (src/action.hh:64-67)
    bad = [i for i in range(flat.n_leaves) if flat.leaves[i]["window_increment"] < 0]
    if bad:
        with pytest.raises(RuntimeError, match="Ineligible value"):
            t.next_generation(bad[0])


def test_config_range_expansion(host):
    """Evaluator::Evaluator(ConfigRange) (src/evaluator.cc:9-38): BASELINE config 2 -> 800 NetConfigs in
    the reference's nesting order; both wire forms of field 77 are accepted."""
    inf = float(abi.BUFFER_INFINITE)
    args = dict(link=(1, 2, 0.25), rtt=(100, 200, 25), nsrc=(1, 32, 1), buf=(inf, inf, 0), off=(5000, 5000, 0),
                on=(5000, 5000, 0), ticks=100000)
    want = workloads.default_config_range(ticks=1e5)
    for unicorn in (False, True):
        raw = hostlib.encode_config_range(unicorn=unicorn, **args)
        got, dna_bytes = host.expand(raw)
        assert len(got) == 800 and got.tobytes() == want.tobytes()
        assert dna_bytes == hostlib.encode_config_range(unicorn=False, **args)   # ConfigRange::DNA, protoc order
    got, _ = host.expand(hostlib.encode_config_range(unicorn=False, **args), carefulness=0.1)
    assert (got["simulation_ticks"] == 10000.0).all()
    # increments that do not divide the range, and isOne() early exits
    got, _ = host.expand(hostlib.encode_config_range(link=(1, 2, 0.3), rtt=(150, 150, 10), nsrc=(2, 9, 3), buf=(10, 10, 0),
                                                     off=(500, 500, 0), on=(5000, 6000, 0), ticks=1000, loss=(0, 0.05, 0.02)))
    assert len(got) == 4 * 1 * 3 * 1 * 1 * 1 * 3
    assert got["num_senders"][:9:3].tolist() == [2, 5, 8] and got["buffer_size"][0] == 10


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not present")
def test_reference_cfg_files(host):
    """Every file under config/ is a ConfigRangeUnicorn; python and C++ readers agree."""
    paths = sorted(glob.glob(os.path.join(REF, "config", "*.cfg")))
    assert len(paths) > 700
    for path in paths[::7]:
        raw = open(path, "rb").read()
        cr = dna.parse_config_range(raw)
        got, _ = host.expand(raw)
        if all(cr[k]["low"] == cr[k]["high"] or cr[k]["incr"] == 0 for k in ("link_ppt", "rtt", "num_senders", "buffer_size", "mean_off_duration", "mean_on_duration")):
            assert len(got) == 1, path
            g = got[0]
            assert g["link_ppt"] == cr["link_ppt"]["low"] and g["delay"] == cr["rtt"]["low"]
            assert g["num_senders"] == int(cr["num_senders"]["low"]) and g["simulation_ticks"] == float(int(cr["simulation_ticks"]["low"]))
    raw = open(os.path.join(REF, "config", "2_2_really_small_buffer_0.cfg"), "rb").read()
    g = host.expand(raw)[0][0]
    assert (g["link_ppt"], g["delay"], g["num_senders"], g["buffer_size"], g["mean_off_duration"], g["simulation_ticks"]) == (2.0, 50.0, 2, 10, 500.0, 600000.0)


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not present")
def test_sender_runner_prints_tree_like_protobuf(host):
    """print_tree (src/sender-runner.cc:15-25) shows DebugString() of the tree's config and
    optimizer sub-messages before scoring; without a GPU the run then fails loudly."""
    import subprocess
    import torch
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "remy_b200", "host", "sender-runner")
    out = subprocess.run([exe, "if=" + os.path.join(REF, "tests", "RemyCC-2014-100x.dna"), "nsrc=2", "link=0.993", "ticks=10"],
                         capture_output=True, text=True, timeout=120)
    assert out.stdout.startswith("Prior assumptions:\nlink_packets_per_ms {\n  low: 0.316227766016838\n  high: 31.6227766016838\n}\nrtt {\n  low: 150\n  high: 150\n}\n")
    assert "Remy optimization settings:\nwindow_increment {\n  min_value: 0\n  max_value: 256\n  min_change: 1\n  max_change: 32\n  multiplier: 4\n  default_value: 1\n}\n" in out.stdout
    assert "Setting num_senders to 2" in out.stderr
    if not torch.cuda.is_available():
        assert out.returncode == 1 and "no CUDA device" in out.stderr


def test_wire_formats_against_python_protobuf(host):
    """Problem / Answer / WhiskerTree / ConfigRange bytes written by the hand-rolled wire code equal what a
    real protobuf implementation writes for the same messages (python-protobuf, schemas rebuilt in
    tests/pb_schema.py), and parse back to the same values."""
    pb = pytest.importorskip("google.protobuf")
    import pb_schema
    cls = pb_schema.classes()
    flat = workloads.load_golden_tree("RemyCC-2013-delta0.1")
    tree = host.tree_from_flat(flat)
    raw_tree = tree.serialize()
    msg = cls["WhiskerTree"]()
    msg.ParseFromString(raw_tree)
    assert msg.SerializeToString() == raw_tree
    inf = float(abi.BUFFER_INFINITE)
    raw_range = hostlib.encode_config_range(link=(1, 2, 0.5), rtt=(100, 150, 50), nsrc=(1, 3, 2), buf=(inf, inf, 0),
                                            off=(5000, 5000, 0), on=(5000, 5000, 0), ticks=1234)
    rng = cls["ConfigRange"]()
    rng.ParseFromString(raw_range)
    assert rng.SerializeToString() == raw_range and rng.simulation_ticks == 1234 and rng.rtt.incr == 50
    # Evaluator::DNA(tree) -> ProblemBuffers.Problem
    problem = tree.problem_dna(raw_range, 4242)
    p = cls["Problem"]()
    p.ParseFromString(problem)
    assert p.SerializeToString() == problem
    assert p.settings.prng_seed == 4242 and p.settings.tick_count == 1234 and len(p.configs) == 3 * 2 * 2
    assert p.whiskers.SerializeToString() == raw_tree
    c = p.configs[5]
    cfgs, _ = host.expand(raw_range)
    assert (c.link_ppt, c.delay, c.num_senders, c.buffer_size) == (cfgs[5]["link_ppt"], cfgs[5]["delay"], cfgs[5]["num_senders"], cfgs[5]["buffer_size"])
    # AnswerBuffers.Outcome written by protobuf -> Outcome(dna) -> Outcome::DNA()
    a = cls["Outcome"]()
    a.score = -12.5
    for k in range(3):
        td = a.throughputs_delays.add()
        td.config.CopyFrom(p.configs[k])
        for s in range(int(p.configs[k].num_senders)):
            r = td.results.add()
            r.throughput, r.delay = 0.5 + s + k, 151.25 * (s + 1)
    answer = a.SerializeToString()
    assert host.answer_roundtrip(answer) == answer


def test_per_config_seed_rule(host):
    """One rule in three places: the C++ host mirror (Evaluator::seeds), the Python workloads (seeds_for) and the
    binding shown in INTEGRATION.md - config 0 runs on the Evaluator's seed itself (what the reference's loop
    starts with, src/evaluator.cc:84, so that a single-config score stays bit-identical to the reference's),
    config k >= 1 on the k-th output of minstd_rand0(seed)."""
    for seed in (0, 1, 12345, 2147483647, 2147483648, 4294967295):
        got = host.seeds(seed, 9)
        assert (got == workloads.seeds_for(seed, 9)).all(), seed
        x = seed % 2147483647 or 1      # what `PRNG seeder(prng_seed)` holds (bits/random.tcc:116-126)
        want = [seed]
        for _ in range(8):
            x = x * 16807 % 2147483647
            want.append(x)               # seeder() in the INTEGRATION.md snippet
        assert got.tolist() == want, seed
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert "seeds.empty() ? prng_seed : (uint32_t)seeder()" in text
