#!/usr/bin/env python3
"""Generate tests/golden/ from the reference itself (run in the build container only).

  trees/<name>.npz   the reference's test fixtures tests/RemyCC-*.dna, flattened to the
                     POD layout of include/remy_b200.h (inputs; /root/reference does not
                     exist on the GPU box)
  vectors.json       expected outputs of oracle/_ref/libremy_ref.so (the reference's own
                     sources, oracle/build_ref.sh) for a fixed list of (tree, candidates,
                     configs, seeds) cases.  Doubles are stored as C99 hex strings.

The cases are defined below; re-running this script must reproduce the committed files.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remy_b200 import abi, dna  # noqa: E402
from oracle import bindings as checkers  # noqa: E402

REF_TESTS = "/root/reference/tests"
OUT = os.path.join(ROOT, "tests", "golden")
TREES = ["RemyCC-2014-100x", "RemyCC-2013-delta1", "RemyCC-2013-delta0.1", "RemyCC-2013-delta10"]
INF = abi.BUFFER_INFINITE

CASES = [
    # SURVEY.md §8c sample vectors (default tree, PRNG(12345))
    dict(name="survey-default", tree="default", seeds=[12345, 12345, 12345], cands=None, configs=[
        dict(link_ppt=1, delay=100, num_senders=1, simulation_ticks=1e6),
        dict(link_ppt=2, delay=50, num_senders=2, buffer_size=10, simulation_ticks=6e5),
        dict(link_ppt=2, delay=50, num_senders=16, buffer_size=10, stochastic_loss_rate=0.01, simulation_ticks=6e5)]),
    # BASELINE config 1 (config/2_2_really_small_buffer_0.cfg decoded, on -> 5000)
    dict(name="config1", tree="default", seeds=[1, 2, 3, 4], cands=None, configs=[
        dict(link_ppt=2, delay=50, num_senders=2, buffer_size=10, mean_off_duration=500, simulation_ticks=6e5)] * 4),
    # the reference's verify-2014 / maintain-2013 test shapes, shortened
    dict(name="verify-2014", tree="RemyCC-2014-100x", seeds=[11, 12, 13, 14], cands=None, configs=[
        dict(link_ppt=l, delay=150, num_senders=2, mean_on_duration=1000, mean_off_duration=1000, simulation_ticks=1e5)
        for l in (0.221, 0.993, 2.941, 18.504)]),
    dict(name="maintain-2013", tree=None, seeds=[21, 22], cands=None, configs=[
        dict(link_ppt=1.5, delay=150, num_senders=8, mean_on_duration=10000, mean_off_duration=10000, simulation_ticks=1e5)] * 2),
    # corners of the default ConfigRange (BASELINE config 2)
    dict(name="config2-corners", tree="RemyCC-2014-100x", seeds=[31, 32, 33, 34, 35, 36], cands=None, configs=[
        dict(link_ppt=l, delay=r, num_senders=n, simulation_ticks=5e4)
        for (l, r, n) in ((1, 100, 1), (2, 200, 32), (1.25, 125, 7), (1.75, 175, 16), (1.5, 150, 31), (2, 100, 3))]),
    # single-leaf candidates (BASELINE config 3): common random numbers across candidates
    dict(name="candidates", tree="RemyCC-2014-100x", seeds=[41, 42, 43],
         cands=[(0, 0, 0, abi.NO_OVERRIDE), (0.9, 0.5, 3, 5), (0.5, 0.0, 0, 17), (1.0, 2.5, 16, 40)], configs=[
        dict(link_ppt=1.5, delay=100, num_senders=4, simulation_ticks=4e4),
        dict(link_ppt=2, delay=200, num_senders=12, simulation_ticks=4e4),
        dict(link_ppt=1, delay=150, num_senders=24, buffer_size=100, simulation_ticks=4e4)]),
    # 16-sender config/*.cfg shapes (BASELINE config 4): small buffers, stochastic loss
    dict(name="config4-16senders", tree="RemyCC-2013-delta1", seeds=[51, 52, 53, 54, 55], cands=None, configs=[
        dict(link_ppt=l, delay=r, num_senders=16, buffer_size=b, stochastic_loss_rate=p, mean_off_duration=500,
             mean_on_duration=on, simulation_ticks=3e4)
        for (l, r, b, p, on) in ((0.5, 150, 10, 0, 5000), (2, 50, 10, 0.01, 5000), (7.5, 30, 57, 0.05, 5000),
                                 (10, 25, 23, 0, 5000), (4, 50, 1000000, 0.01, 80))]),
    # edge cases: no buffer, off = 0 (immediate re-start), delay 0, bursts (intersend 0), zero windows
    dict(name="edges", tree="default", seeds=[61, 62, 63, 64, 65, 66],
         cands=[(0, 0, 0, abi.NO_OVERRIDE), (1.0, 0.0, 5, 0), (0.5, 0.0, 0, 0)], configs=[
        dict(link_ppt=1, delay=50, num_senders=3, buffer_size=0, simulation_ticks=2e4),
        dict(link_ppt=1, delay=50, num_senders=8, buffer_size=5, mean_off_duration=0, mean_on_duration=50, simulation_ticks=5e3),
        dict(link_ppt=1, delay=0, num_senders=2, buffer_size=INF, mean_on_duration=50, mean_off_duration=20, simulation_ticks=3e3),
        dict(link_ppt=3.3, delay=1, num_senders=32, buffer_size=2, stochastic_loss_rate=0.5, mean_on_duration=10, mean_off_duration=3, simulation_ticks=3e3),
        dict(link_ppt=0.1, delay=77.7, num_senders=31, buffer_size=INF, simulation_ticks=12345.678),
        dict(link_ppt=2, delay=100, num_senders=1, buffer_size=1, simulation_ticks=1)]),
]


def hexf(x):
    return float(x).hex()


def main():
    os.makedirs(os.path.join(OUT, "trees"), exist_ok=True)
    trees = {"default": abi.default_tree()}
    for name in TREES:
        t = dna.load_dna(os.path.join(REF_TESTS, name + ".dna"))
        np.savez_compressed(os.path.join(OUT, "trees", name + ".npz"), nodes=t.nodes, leaves=t.leaves)
        trees[name] = t
    ref = checkers.Reference()
    out = []
    for case in CASES:
        tree_names = [case["tree"]] if case["tree"] else TREES[1:]
        for tn in tree_names:
            tree = trees[tn]
            cfgs = abi.make_configs(case["configs"])
            seeds = np.array(case["seeds"], dtype=np.uint32)
            cands = None
            if case["cands"]:
                cands = np.zeros(len(case["cands"]), dtype=abi.leaf_override_dtype)
                for i, c in enumerate(case["cands"]):
                    cands[i] = c
            sims, snd, cnt = ref.score(tree, cfgs, seeds, cands=cands, want_use_counts=True, threads=8)
            rec = dict(name=case["name"] + ("" if case["tree"] else "/" + tn), tree=tn,
                       configs=[{k: (hexf(v) if isinstance(v, float) else int(v)) for k, v in
                                 zip(cfgs.dtype.names, map(lambda x: x.item(), row))} for row in cfgs],
                       seeds=[int(s) for s in seeds],
                       cands=None if cands is None else [[hexf(c[0]), hexf(c[1]), int(c[2]), int(c[3])] for c in cands],
                       sims=[dict(utility=hexf(s["utility"]), last_tick=hexf(s["last_tick"]),
                                  event_ticks=int(s["event_ticks"]), packets_sent=int(s["packets_sent"]),
                                  packets_received=int(s["packets_received"]), link_queued=int(s["link_queued"]),
                                  delay_inflight=int(s["delay_inflight"]), prng_state=int(s["prng_state"]))
                             for s in sims],
                       senders=[[[hexf(x["tick_share_sending"]), hexf(x["total_delay"]), int(x["packets_received"]),
                                  int(x["packets_sent"])] for x in row[:int(cfgs[i % len(cfgs)]["num_senders"])]]
                                for i, row in enumerate(snd)],
                       use_counts=[[int(v) for v in row] for row in cnt])
            out.append(rec)
            print(rec["name"], "sims", len(sims), "event ticks", int(sims["event_ticks"].sum()))
    with open(os.path.join(OUT, "vectors.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", os.path.join(OUT, "vectors.json"), os.path.getsize(os.path.join(OUT, "vectors.json")), "bytes")


if __name__ == "__main__":
    main()
