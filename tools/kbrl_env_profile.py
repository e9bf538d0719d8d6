#!/usr/bin/env python3
"""Where the env step goes under the KBRL policy (BASELINE configs[2]): per-kernel times of the serialised profiling pass
(lane-per-unit kernel | heavy list on the warp kernel + general kernel) and the live step time, for several heavy-list
thresholds, after --steps steps of the device-resident control loop.  One JSON line per threshold.

    python tools/kbrl_env_profile.py --envs 16384 --steps 300 --thresholds 0,600,1000,1500
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ranslice_b200 import _lib, create_batched_env  # noqa: E402
from ranslice_b200.kbrl import create_kbrl_agent  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--thresholds", default="0,600,1000,1500")
a = ap.parse_args()
env = create_batched_env(20260000, 0, a.envs)
env.set_heavy_threshold(1000)
agent = create_kbrl_agent(np.random.default_rng(0), 0, accuracy_range=(0.97, 0.99), n_envs=a.envs, dict_cap=2048, resident=True)
env.reset()
state = torch.zeros((a.envs, env.n_variables), dtype=torch.float32, device=agent.device)
action = agent.action
bufs = [None, None]
nxt = [torch.empty_like(action), torch.empty_like(action)]
hits = torch.empty((a.envs, 5), dtype=torch.int32, device=agent.device)
it = [0]


def loop(n, timed=False):
    global state, action
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n)]
    for k in range(n):
        i = it[0]; it[0] += 1
        ev[2 * k].record()
        out = env.step_device(action, bufs[i & 1]); bufs[i & 1] = out
        ev[2 * k + 1].record()
        agent.update_control(state, action, out["labels"], hits_out=hits)
        action, _ = agent.select_action(out["obs"], action_out=nxt[i & 1], adjusted_out=agent.adjusted)
        state = out["obs"]
    torch.cuda.synchronize()
    return sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(n)) / n


loop(a.steps)
L = _lib.lib()
for thr in [int(x) for x in a.thresholds.split(",")]:
    env.set_heavy_threshold(thr)
    loop(5)
    live = loop(10)
    _lib.check(L.rs_set_profiling(env._h, 1))
    loop(10)
    ms, n = (C.c_double * 6)(), C.c_uint64()
    _lib.check(L.rs_get_profile(env._h, ms, C.byref(n)))
    _lib.check(L.rs_set_profiling(env._h, 0))
    k = max(int(n.value), 1)
    print(json.dumps({"heavy_threshold": thr, "after_steps": it[0], "env_ms_live": live, "sort_ms": ms[0] / k, "lane_per_unit_ms": ms[1] / k,
                      "heavy_warp_plus_general_ms": ms[2] / k, "routes": env.routes(), "mean_action": float(action.float().mean().item()),
                      "live_ues_per_slice": float(env.n_ues().mean())}), flush=True)
