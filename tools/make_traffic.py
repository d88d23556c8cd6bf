#!/usr/bin/env python3
"""profiles/traffic.json from ncu output of the bench command, tagged with the hash of the kernel build it was
captured on (bench.py refuses a capture of another build or workload).

usage: make_traffic.py <launches.csv> <workload_key> [full.ncu-rep]
  launches.csv : ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
                 --csv --log-file launches.csv python bench.py --steps 1 --warmup 3 [...]
  workload_key : the key bench.py prints to stderr ("traffic key: ...") for the same arguments
  full.ncu-rep : optional ncu --set full capture of the simulation kernel of the SAME build: issue slots, lanes,
                 long-scoreboard stall and cache hit rates are copied into "counters"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

path, wkey = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path)) if r and r[0].strip('"').isdigit()]
hdr = next(r for r in csv.reader(open(path)) if r and r[0].strip('"') == "ID")
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows:
    if "remy_sim_kernel" not in r[ix["Kernel Name"]]:
        continue
    d = launches.setdefault(r[ix["ID"]], {})
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9,
             "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
    d[r[ix["Metric Name"]]] = val * scale
big = max(launches.values(), key=lambda d: d.get("gpu__time_duration.sum", 0))   # the whole-batch launch of a step
out = {"build_hash": bench.build_hash(), "workload_key": wkey,
       "dram_bytes_per_launch": big["dram__bytes_read.sum"] + big["dram__bytes_write.sum"],
       "read": big["dram__bytes_read.sum"], "write": big["dram__bytes_write.sum"],
       "kernel_ns_under_ncu": big["gpu__time_duration.sum"],
       "what": "dram__bytes_read.sum + dram__bytes_write.sum of the ONE remy_sim_kernel launch of a bench step (ncu --metrics, same command as the bench line)",
       "source": f"profiles/{os.path.basename(path)}"}
if len(sys.argv) > 3:
    raw = subprocess.run(["ncu", "-i", sys.argv[3], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    t = list(csv.reader(io.StringIO(raw)))
    h, row = t[0], t[2]
    want = {"smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
            "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "long_scoreboard_cycles_per_issue",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "wait_cycles_per_issue",
            "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
            "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
            "launch__registers_per_thread": "registers_per_thread", "launch__grid_size": "grid"}
    out["counters"] = {v: float(row[h.index(k)].replace(",", "")) for k, v in want.items() if k in h}
    out["counters"]["source"] = f"profiles/{os.path.basename(sys.argv[3])} (ncu --set full of the whole-batch launch, probe workload)"
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
