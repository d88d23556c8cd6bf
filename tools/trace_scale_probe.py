#!/usr/bin/env python3
"""Scale check of trace == true scoring (not a test): Breeder::apply_best_split over the default
800-config ConfigRange on the GPU path; prints wall times and the size of the lookup log."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from remy_b200 import abi, hostlib, workloads

ticks = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000
dev = abi.Device(0)
host = hostlib.HostLib()
flat = workloads.load_golden_tree("RemyCC-2014-100x")
inf = float(abi.BUFFER_INFINITE)
raw = hostlib.encode_config_range(link=(1, 2, 0.25), rtt=(100, 200, 25), nsrc=(1, 32, 1), buf=(inf, inf, 0),
                                  off=(5000, 5000, 0), on=(5000, 5000, 0), ticks=ticks)
cfgs, _ = host.expand(raw)
seeds = workloads.seeds_for(7, len(cfgs))
t0 = time.time()
sims, snd, counts, recs = None, None, None, None
plain = dev.score(flat, cfgs, seeds, want_use_counts=True)
t1 = time.time()
print(f"{len(cfgs)} configs x {ticks} ticks: plain scoring {t1 - t0:.2f} s, lookups {int(plain[2].sum()):,} "
      f"(log {plain[2].sum() * 40 / 1e9:.2f} GB)", flush=True)
tree = host.tree_from_flat(flat)
t0 = time.time()
tree.apply_best_split(raw, 99, 0)
t1 = time.time()
print(f"apply_best_split (counting run + writing run + D2H + P^2 feed on one host thread + bisect): {t1 - t0:.2f} s; "
      f"leaves {flat.n_leaves} -> {tree.flatten().n_leaves}", flush=True)
