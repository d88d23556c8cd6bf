#!/usr/bin/env python3
"""BASELINE configs[3] ("config 4"): every nsrc = 16 file under the reference's
config/ (99 `16_*.cfg` + 6 `figure4_16_*.cfg`) decoded to NetConfigs with the .cfg -> NetConfig
mapping of SURVEY.md Appendix A (the C++ host reader: on <= 0 -> 5000 ms).  /root/reference does not
travel to the GPU box, so the decoded values are committed as tests/golden/config4.json."""
import glob, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from remy_b200 import hostlib

REF = os.environ.get("REMY_REFERENCE", "/root/reference")
host = hostlib.HostLib()
rows = []
paths = sorted(glob.glob(os.path.join(REF, "config", "16_*.cfg")) + glob.glob(os.path.join(REF, "config", "figure4_16_*.cfg")))
for p in paths:
    cfgs, _ = host.expand(open(p, "rb").read())
    assert len(cfgs) == 1, p
    c = cfgs[0]
    on = float(c["mean_on_duration"])
    rows.append(dict(file=os.path.basename(p), num_senders=int(c["num_senders"]), link_ppt=float(c["link_ppt"]), delay=float(c["delay"]),
                     buffer_size=int(c["buffer_size"]), stochastic_loss_rate=float(c["stochastic_loss_rate"]),
                     mean_on_duration=on if on > 0 else 5000.0, mean_off_duration=float(c["mean_off_duration"]),
                     simulation_ticks=float(c["simulation_ticks"])))
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "config4.json")
json.dump(rows, open(out, "w"), indent=0)
print(len(rows), "configs ->", out)
