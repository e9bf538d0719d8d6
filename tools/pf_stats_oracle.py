#!/usr/bin/env python3
"""PF-loop statistics from the CPU oracle (ORC_PF_STATS build): how the winner of ProportionalFair.allocate's RB loop
(schedulers.py:47-63) moves between UEs.  Used to size the warp kernel's argmax bookkeeping.

    gcc -O2 -fPIC -std=c11 -ffp-contract=off -DORC_PF_STATS -shared -pthread -o gpurun_out/liboracle_pfstats.so \
        oracle/ranslice_oracle.c oracle/kbrl_oracle.c -lm
    RANSLICE_ORACLE_LIB=gpurun_out/liboracle_pfstats.so python tools/pf_stats_oracle.py [--mux]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from ranslice_b200.tables import load_tables  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=48)
ap.add_argument("--burn", type=int, default=600)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--mux", action="store_true")
a = ap.parse_args()
tables = load_tables()
b = ol.OracleBatch(tables, 0, a.envs, 20260000, n_threads=1, l1_mux=a.mux)
b.reset()
rng = np.random.default_rng(5)
stat = (C.c_ulonglong * 12).in_dll(ol.lib(), "orc_pf_stat")
S, n = b.S, b.envs[0].n_prbs
for i in range(a.burn + a.steps):
    if i == a.burn:
        for k in range(12):
            stat[k] = 0
    if a.mux:
        act = rng.integers(60, 200, (a.envs, S))
    else:
        w = rng.random((a.envs, S + 1))
        act = np.floor(n * w[:, :S] / w.sum(1, keepdims=True)).astype(np.int64)
    b.step(act)
chunks, cont, runs, changes, to2, keep3, bsteps, bcuts = [int(stat[k]) for k in range(8)]
print("chunks %d  contended %d (%.1f%%)  runs %d (%.2f chunks/run)  winner changes %d: to old runner-up %.1f%%, top-3 set kept %.1f%%"
      % (chunks, cont, 100.0 * cont / chunks, runs, cont / max(runs, 1), changes, 100.0 * to2 / max(changes, 1), 100.0 * keep3 / max(changes, 1)))
print("batched what-if: %d warp-wide steps (%.2f contended chunks per step; the current loop takes %d), %d of them with a budget cut"
      % (bsteps, cont / max(bsteps, 1), runs, bcuts))
att, empty = int(stat[8]), int(stat[9])
print("   batch attempts %d (%d empty), plain runs %d; cost model (run 1, batch 3, cut +1.5, empty batch +2): %.0f vs %d now"
      % (att, empty, bsteps - att + empty, (bsteps - att + empty) + 3.0 * (att - empty) + 1.5 * bcuts + 2.0 * empty, runs))
