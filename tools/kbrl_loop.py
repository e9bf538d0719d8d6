#!/usr/bin/env python3
"""BASELINE config 3: scenario_0 with the KBRL controller in the loop (env step + update_control + select_action every
step), as SURVEY 8d specifies it: 16 384 envs, 2000 steps, dictionary-size statistics and cap hits.  Prints ONE JSON
line: throughput over the whole run and, at every checkpoint (--report 200,1000,2000), the per-phase device times of
the steps just before it (env / update_control / select_action, CUDA events) with the dictionary statistics there.

    python tools/kbrl_loop.py --envs 16384 --steps 2000 --resident                 # controller state in HBM, no host round trip
    python tools/kbrl_loop.py --envs 16384 --steps 30 --warm 20                    # host-side controller (numpy mirror)
"""
import argparse
import hashlib
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
ap.add_argument("--steps", type=int, default=30, help="timed steps (after --warm untimed ones)")
ap.add_argument("--warm", type=int, default=20)
ap.add_argument("--dict-cap", type=int, default=1024)
ap.add_argument("--pool-mb", type=int, default=0)
ap.add_argument("--resident", action="store_true")
ap.add_argument("--heavy", type=int, default=1800, help="heavy-list threshold of the env under the KBRL policy (ranslice_b200.kbrl.KBRL_HEAVY_THRESHOLD); -1: library default")
ap.add_argument("--report", default="", help="comma-separated step counts at which per-phase times and dictionary statistics are sampled")
ap.add_argument("--window", type=int, default=20, help="steps averaged ahead of each checkpoint")
a = ap.parse_args()
env = create_batched_env(20260000, 0, a.envs)
if a.heavy >= 0:
    env.set_heavy_threshold(a.heavy)
agent = create_kbrl_agent(np.random.default_rng(0), 0, accuracy_range=(0.97, 0.99), n_envs=a.envs, dict_cap=a.dict_cap,
                          resident=a.resident, pool_mb=a.pool_mb)
res = {"workload": "scenario_0 + KBRL in the loop (BASELINE configs[2])", "envs": a.envs, "steps": a.steps, "warm": a.warm,
       "resident": a.resident, "dict_cap": a.dict_cap}


def dict_stats():
    sizes, flags = agent.learners.sizes()
    pool = agent.learners.pool()
    return {"dict_mean": float(sizes.mean()), "dict_p50": float(np.percentile(sizes, 50)), "dict_p99": float(np.percentile(sizes, 99)),
            "dict_max": int(sizes.max()), "cap_hits": int((flags & 1).sum()), "pool_hits": int((flags & 2).sum()),
            "pool_used_gb": pool["used_bytes"] / 1e9, "pool_total_gb": pool["total_bytes"] / 1e9, "tie_breaks": pool["tie_breaks"],
            "updates_last_step": agent.learners.counters()[1],
            "digest_sizes": hashlib.sha1(np.ascontiguousarray(sizes).tobytes()).hexdigest()[:12]}


checkpoints = sorted(int(x) for x in a.report.split(",") if x)
total = a.warm + a.steps
if a.resident:
    import torch
    dev = agent.device
    env.reset()
    state = torch.zeros((a.envs, env.n_variables), dtype=torch.float32, device=dev)
    action = agent.action
    bufs = [None, None]
    hits = torch.empty((a.envs, 5), dtype=torch.int32, device=dev)
    nxt = [torch.empty_like(action), torch.empty_like(action)]
    viol = torch.zeros((), dtype=torch.int64, device=dev)
    res_sum = torch.zeros((), dtype=torch.int64, device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(a.window)]
    reports = []
    t_wall0 = None
    for i in range(total):
        if i == a.warm:
            torch.cuda.synchronize()
            t_wall0 = time.perf_counter()
        e = ev[i % a.window]
        e[0].record()
        out = env.step_device(action, bufs[i & 1]); bufs[i & 1] = out
        e[1].record()
        agent.update_control(state, action, out["labels"], hits_out=hits)
        e[2].record()
        action, _ = agent.select_action(out["obs"], action_out=nxt[i & 1], adjusted_out=agent.adjusted)
        e[3].record()
        state = out["obs"]
        viol += out["violations"].sum()
        res_sum += action.sum()
        if (i + 1) in checkpoints or i + 1 == total:            # sample the last `window` steps (one sync per checkpoint)
            torch.cuda.synchronize()
            t_pause = time.perf_counter()
            k = min(a.window, i + 1)
            te = sum(x[0].elapsed_time(x[1]) for x in ev[:k]) / k
            tu = sum(x[1].elapsed_time(x[2]) for x in ev[:k]) / k
            ts = sum(x[2].elapsed_time(x[3]) for x in ev[:k]) / k
            r = {"after_steps": i + 1, "ms_env": te, "ms_update_control": tu, "ms_select_action": ts,
                 "env_steps_per_s_here": a.envs / ((te + tu + ts) * 1e-3)}
            r.update(dict_stats())
            reports.append(r)
            if t_wall0 is not None:
                t_wall0 += time.perf_counter() - t_pause           # the statistics read-back is not part of the loop
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall0
    res.update({"env_steps_per_s": a.envs * a.steps / wall, "ms_per_step_wall": 1e3 * wall / a.steps,
                "violations_per_env_step": float(viol.item()) / (a.envs * total),
                "mean_resources": float(res_sum.item()) / (a.envs * total), "checkpoints": reports})
    res.update({k: v for k, v in reports[-1].items() if k.startswith(("ms_", "dict_", "cap_", "pool_", "tie_", "updates_", "digest_"))})
else:
    state = env.reset()
    action = agent.action
    t_env = t_upd = t_sel = 0.0
    for i in range(total):
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
    res.update(dict_stats())
res["after_steps"] = total
try:                                    # KB_CHECK builds only: fast-path validation counters
    import ctypes
    from ranslice_b200 import _lib
    dbg = (ctypes.c_ulonglong * 4)()
    if _lib.lib().kb_debug_counters(dbg) == 0:
        res["kb_check"] = {"accepted": dbg[0], "sign_mismatch": dbg[1], "max_err_over_G": dbg[2] * 1e-9, "guard_hits": dbg[3]}
except AttributeError:
    pass
print(json.dumps(res))
