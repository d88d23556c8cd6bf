#!/usr/bin/env python3
"""Quick device-side throughput probe (not the bench): config-2 shaped batch."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from remy_b200 import abi, workloads

ticks = float(sys.argv[1]) if len(sys.argv) > 1 else 1e4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
treename = sys.argv[3] if len(sys.argv) > 3 else "default"
nmax = int(sys.argv[4]) if len(sys.argv) > 4 else 32
kind = sys.argv[5] if len(sys.argv) > 5 else "rat"   # rat | fish | flows
dev = abi.Device(0, path=os.environ.get("REMY_LIB"))
tree = abi.default_tree()
if treename != "default":
    tree = workloads.load_golden_tree(treename)
if kind == "fish":
    tree = abi.fin_tree([0.02, 0.01, 0.004], [2.0, 20.0])   # bench.py's FinTree
cfgs = np.tile(workloads.default_config_range(ticks=ticks, nsrc=(1, nmax, 1)), reps)
seeds = workloads.seeds_for(1, len(cfgs))
if kind == "flows":
    cfgs["mean_on_duration"] = 80.0; cfgs["mean_off_duration"] = 500.0   # bench.py's flows-pass
b = dev.create_batch(tree, cfgs, seeds, fish=kind == "fish", flows=kind == "flows")
for it in range(3):
    t0 = time.time(); ms = b.run(timed=True); t1 = time.time()
    sims, _, _ = b.fetch()
    ev = int(sims["event_ticks"].sum()); pk = int((sims["packets_sent"] + sims["packets_received"]).sum())
    print(f"run {it}: {len(sims)} sims kernel {ms:.1f} ms wall {1e3*(t1-t0):.1f} ms  ticks {ev:.3e} -> {ev/ms/1e3:.1f} M ticks/s  "
          f"pkt-events {pk:.3e} -> {pk/ms/1e3:.1f} M/s  errs {np.unique(sims['error'])} launches {b.launches()}", flush=True)
