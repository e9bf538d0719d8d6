#!/usr/bin/env python3
"""PF-loop workload statistics from an RS_STATS experiment build (RS_BUILD_TAG=stats RS_NVCC_EXTRA=-DRS_STATS).
    RS_B200_LIB=.../libranslice_b200_stats.so python tools/pf_stats.py --envs 65536 --burn-in 600 --steps 4 [--kbrl]"""
import argparse, ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ranslice_b200 import create_batched_env, _lib  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536); ap.add_argument("--burn-in", type=int, default=600); ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--kbrl", action="store_true")
a = ap.parse_args()
env = create_batched_env(20260000, 0, a.envs); env.reset()
L = _lib.lib(); buf = (ctypes.c_ulonglong * 64)()
rng = np.random.default_rng(0)
if a.kbrl:
    from ranslice_b200.kbrl import create_kbrl_agent
    agent = create_kbrl_agent(np.random.default_rng(0), 0, accuracy_range=(0.97, 0.99), n_envs=a.envs, dict_cap=128, resident=True)
    state = torch.zeros((a.envs, 50), dtype=torch.float32, device="cuda"); action = agent.action; bufs = [None, None]
    for i in range(a.burn_in + a.steps):
        if i == a.burn_in: L.rs_debug_stats(buf, 1)
        out = env.step_device(action, bufs[i & 1]); bufs[i & 1] = out
        agent.update_control(state, action, out["labels"])
        action, _ = agent.select_action(out["obs"], action_out=torch.empty_like(action), adjusted_out=agent.adjusted)
        state = out["obs"]
else:
    acts = []
    for i in range(8):
        w = rng.random((a.envs, 6), dtype=np.float32)
        acts.append(torch.from_numpy(np.floor(200 * w[:, :5] / w.sum(axis=1, keepdims=True)).astype(np.int32)).cuda())
    out = None
    for i in range(a.burn_in + a.steps):
        if i == a.burn_in: L.rs_debug_stats(buf, 1)
        out = env.step_device(acts[i % 8], out)
L.rs_debug_stats(buf, 0)
s = np.array(list(buf), dtype=np.float64)
tt = s[0]
print("unit-TTIs %.3e  scheduled %.1f%%" % (tt, 100 * s[1] / tt))
print("phase-1 iterations per unit-TTI %.2f  (same UE as previous chunk: %.1f%%, drained-by-chunk: %.1f%%, tie path: %.2f%%)" % (s[5] / tt, 100 * s[6] / max(s[5], 1), 100 * s[4] / max(s[5], 1), 100 * s[2] / max(s[5], 1)))
print("avg n_ues scanned per iteration %.2f ; TTIs ending in single-UE closed form %.1f%%" % (s[3] / max(s[5], 1), 100 * s[7] / tt))
print("n_backlog at TTI start 0..7+ (%%):", np.round(100 * s[8:16] / tt, 1))
print("n_ues 0..15+ (%%):", np.round(100 * s[16:32] / tt, 1))
print("phase-1 iterations per scheduled TTI, buckets 0,1,2-3,4-7,8-15,16-31,32-63,64+ (%% of TTIs):", np.round(100 * s[32:40] / max(s[1], 1), 1))
print("   share of all iterations in each bucket (%%):", np.round(100 * s[40:48] / max(s[5], 1), 1))
