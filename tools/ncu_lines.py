#!/usr/bin/env python3
"""Per-source-line hot spots from an .ncu-rep captured with --import-source on (-lineinfo build).
    python tools/ncu_lines.py rep.ncu-rep [top_n]
Prints, per CUDA source line: warp instructions executed, avg active threads, stall samples."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
lines = {}
tot_inst = tot_thr = tot_samp = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_thr = hdr.index("Thread Instructions Executed")
        i_samp = hdr.index("# Samples")
        i_noinst = hdr.index("stall_no_inst")
        i_lsb = hdr.index("stall_long_sb")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    if r[2] != "-":       # SASS row under a source line; the source row already aggregates
        continue
    key = (cur_file, int(r[0]))
    inst, thr, samp = int(r[i_inst]), int(r[i_thr]), int(r[i_samp])
    lines[key] = (inst, thr, samp, int(r[i_noinst]), int(r[i_lsb]), r[1].strip()[:90])
    tot_inst += inst; tot_thr += thr; tot_samp += samp
print("total warp-inst %d, thread-inst %d (avg %.2f lanes), samples %d" % (tot_inst, tot_thr, tot_thr / max(tot_inst, 1), tot_samp))
print("%-22s %6s %7s %6s %7s %6s %6s  %s" % ("file:line", "inst%", "lanes", "samp%", "noinst", "longsb", "", "source"))
for key, v in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
    inst, thr, samp, noi, lsb, src = v
    print("%-22s %6.2f %7.2f %6.2f %7d %6d         %s" % ("%s:%d" % key, 100.0 * inst / tot_inst, thr / max(inst, 1),
                                                      100.0 * samp / max(tot_samp, 1), noi, lsb, src))
