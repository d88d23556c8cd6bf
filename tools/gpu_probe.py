#!/usr/bin/env python3
"""Quick device-side throughput probe (not the bench): config-2 shaped batch."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from remy_b200 import abi, workloads

ticks = float(sys.argv[1]) if len(sys.argv) > 1 else 1e4
ncand = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = abi.Device(0)
tree = abi.default_tree()
cfgs = workloads.default_config_range(ticks=ticks)
seeds = workloads.seeds_for(1, len(cfgs))
cands = workloads.candidate_grid(0, n=ncand)
b = dev.create_batch(tree, cfgs, seeds, cands=cands)
for it in range(3):
    t0 = time.time(); ms = b.run(timed=True); t1 = time.time()
    sims, _, _ = b.fetch()
    ev = int(sims["event_ticks"].sum()); pk = int((sims["packets_sent"] + sims["packets_received"]).sum())
    print(f"run {it}: {len(sims)} sims kernel {ms:.1f} ms wall {1e3*(t1-t0):.1f} ms  ticks {ev:.3e} -> {ev/ms/1e3:.1f} M ticks/s  "
          f"pkt-events {pk:.3e} -> {pk/ms/1e3:.1f} M/s  errs {np.unique(sims['error'])} launches {b.launches()}", flush=True)
