#!/usr/bin/env python3
"""Instruction / sample share per phase of embb_step_smem from an ncu capture (phase = source-line range
between the '// =====' / '// ----' markers of csrc/embb_smem.cu).   python tools/ncu_regions.py rep.ncu-rep"""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "network-slicing_b200", "csrc", "embb_smem.cu")).read().split("\n")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = hdr = None
res = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit() or r[2] != "-":
        continue
    res[(cur, int(r[0]))] = (int(r[ii]), int(r[it]), int(r[isamp]))


def find(s):
    for i, l in enumerate(src):
        if s in l:
            return i + 1
    raise KeyError(s)


m = [("kernel prologue + gather", find("__global__ void __launch_bounds__(SM_THREADS,")), ("RAN events check", find("arrivals / departures only on event slots")),
     ("per-UE pass (traffic, walk, mean)", find("per-UE traffic + SNR estimate")), ("PF allocate", find("// ================= scheduling + reception")),
     ("MI loop", find("// ---- MI sums of the served")), ("reception + tx step", find("// ---- per-UE reception")),
     ("unscheduled + update_info", find("nothing touched: stale bits") - 1), ("epilogue", find("if (pad) return;")), ("end", 10 ** 6)]
tot = sum(v[0] for v in res.values()); tots = sum(v[2] for v in res.values())


def agg(f, lo, hi):
    i = t = s = 0
    for (ff, l), v in res.items():
        if ff == f and lo <= l < hi:
            i += v[0]; t += v[1]; s += v[2]
    return i, t, s


print("total warp-inst %.3e  avg lanes %.2f  samples %d" % (tot, sum(v[1] for v in res.values()) / tot, tots))
for (name, lo), (_, hi) in zip(m[:-1], m[1:]):
    i, t, s = agg("embb_smem.cu", lo, hi)
    print("%-36s inst %5.1f%% lanes %5.1f samp %5.1f%%" % (name, 100 * i / tot, t / max(i, 1), 100 * s / tots))
first = m[0][1]
fm = open(os.path.join(root, "network-slicing_b200", "csrc", "embb_fastmath.cuh")).read().split("\n")
def fm_find(t):
    for i, l in enumerate(fm):
        if t in l:
            return i + 1
    raise KeyError(t)
ws_lo, ws_hi = fm_find("// Exact integer sum of the window"), fm_find("// exact fp64 window mean")
for name, (f, lo, hi) in [("ran_events_smem + vbr_step (helpers)", ("embb_smem.cu", 1, first)), ("window_sum_fix", ("embb_fastmath.cuh", ws_lo, ws_hi)),
                          ("fastmath exact paths", ("embb_fastmath.cuh", ws_hi, 400)), ("philox", ("philox.cuh", 1, 200)), ("embb_device.cuh", ("embb_device.cuh", 1, 400)),
                          ("__syncwarp", ("sm_30_intrinsics.hpp", 1, 1000))]:
    i, t, s = agg(f, lo, hi)
    print("%-36s inst %5.1f%% lanes %5.1f samp %5.1f%%" % (name, 100 * i / tot, t / max(i, 1), 100 * s / tots))
