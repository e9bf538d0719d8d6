#!/usr/bin/env python3
"""BASELINE config 3 shape: scenario_0 with the KBRL controller in the loop (env step + update_control +
select_action every step).  Prints env-steps/s and dictionary statistics.

    python tools/kbrl_loop.py --envs 16384 --steps 30 --warm 20 --dict-cap 128             # host-side controller (numpy mirror)
    python tools/kbrl_loop.py --envs 16384 --steps 30 --warm 20 --dict-cap 128 --resident  # controller state in HBM, no host round trip
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ranslice_b200 import create_batched_env  # noqa: E402
from ranslice_b200.kbrl import create_kbrl_agent  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--warm", type=int, default=20)
ap.add_argument("--dict-cap", type=int, default=128)
ap.add_argument("--resident", action="store_true")
a = ap.parse_args()
env = create_batched_env(20260000, 0, a.envs)
agent = create_kbrl_agent(np.random.default_rng(0), 0, accuracy_range=(0.97, 0.99), n_envs=a.envs, dict_cap=a.dict_cap,
                          resident=a.resident)
res = {"workload": "scenario_0 + KBRL in the loop", "envs": a.envs, "steps": a.steps, "resident": a.resident}
if a.resident:
    import torch
    dev = agent.device
    env.reset()
    state = torch.zeros((a.envs, env.n_variables), dtype=torch.float32, device=dev)
    action = agent.action
    bufs = [None, None]
    hits = torch.empty((a.envs, 5), dtype=torch.int32, device=dev)
    nxt = [torch.empty_like(action), torch.empty_like(action)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_env = t_upd = t_sel = 0.0
    t_wall0 = None
    for i in range(a.warm + a.steps):
        if i == a.warm:
            torch.cuda.synchronize()
            t_wall0 = time.perf_counter()
        ev[0].record()
        out = env.step_device(action, bufs[i & 1]); bufs[i & 1] = out
        ev[1].record()
        agent.update_control(state, action, out["labels"], hits_out=hits)
        ev[2].record()
        action, _ = agent.select_action(out["obs"], action_out=nxt[i & 1], adjusted_out=agent.adjusted)
        ev[3].record()
        state = out["obs"]
        if i >= a.warm and (i - a.warm) % 5 == 0:           # sample the per-phase device times (sync) every 5th step
            torch.cuda.synchronize()
            t_env += ev[0].elapsed_time(ev[1]); t_upd += ev[1].elapsed_time(ev[2]); t_sel += ev[2].elapsed_time(ev[3])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall0
    k = len(range(0, a.steps, 5))
    res.update({"env_steps_per_s": a.envs * a.steps / wall, "ms_per_step_wall": 1e3 * wall / a.steps, "ms_env": t_env / k,
                "ms_update_control": t_upd / k, "ms_select_action": t_sel / k})
else:
    state = env.reset()
    action = agent.action
    t_env = t_upd = t_sel = 0.0
    for i in range(a.warm + a.steps):
        t0 = time.perf_counter()
        new_state, reward, _, info = env.step(action)
        t1 = time.perf_counter()
        agent.update_control(state, action, info["SLA_labels"])
        t2 = time.perf_counter()
        action, agent.adjusted = agent.select_action(new_state)
        t3 = time.perf_counter()
        state = new_state
        if i >= a.warm:
            t_env += t1 - t0; t_upd += t2 - t1; t_sel += t3 - t2
    tot = t_env + t_upd + t_sel
    res.update({"env_steps_per_s": a.envs * a.steps / tot, "ms_env": 1e3 * t_env / a.steps,
                "ms_update_control": 1e3 * t_upd / a.steps, "ms_select_action": 1e3 * t_sel / a.steps})
sizes, flags = agent.learners.sizes()
import hashlib
res["digest_sizes"] = hashlib.sha1(np.ascontiguousarray(sizes).tobytes()).hexdigest()[:12]
res.update({"dict_mean": float(sizes.mean()), "dict_max": int(sizes.max()), "cap_hits": int((flags & 1).sum()),
            "updates_last_step": agent.learners.counters()[1], "after_steps": a.warm + a.steps})
try:                                    # KB_CHECK builds only: fast-path validation counters
    import ctypes
    from ranslice_b200 import _lib
    dbg = (ctypes.c_ulonglong * 4)()
    if _lib.lib().kb_debug_counters(dbg) == 0:
        res["kb_check"] = {"accepted": dbg[0], "sign_mismatch": dbg[1], "max_err_over_G": dbg[2] * 1e-9, "guard_hits": dbg[3]}
except AttributeError:
    pass
print(json.dumps(res))
