#!/usr/bin/env python3
"""Write profiles/traffic.json (DRAM bytes per launch of the dominant kernel) from an ncu --set full capture.
    python tools/ncu_traffic.py rep.ncu-rep <envs_per_gpu> <scenario>"""
import csv
import json
import os
import subprocess
import sys

rep, envs, scn = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return float(vals[i]) * mult
d = {"kernel": vals[hdr.index("Kernel Name")], "envs_per_gpu": envs, "scenario": scn,
     "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"), "source": os.path.basename(rep)}
d["dram_bytes_per_launch"] = d["dram_bytes_read"] + d["dram_bytes_write"]
# issue-slot roofline inputs (bench.py issue_roofline): warp instructions per launch and active lanes per instruction
d["warp_insts_per_launch"] = float(vals[hdr.index("smsp__inst_executed.sum")])
d["lanes_per_inst"] = float(vals[hdr.index("smsp__thread_inst_executed_per_inst_executed.ratio")])
d["issue_active_pct"] = float(vals[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")])
d["sm_count"] = 148
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(d, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(d)
