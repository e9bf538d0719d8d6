#!/usr/bin/env python3
"""BASELINE config 3 shape: scenario_0 with the KBRL controller in the loop (env step + kb_update + kb_predict
every step).  Prints env-steps/s and dictionary statistics.
    python tools/kbrl_loop.py --envs 16384 --steps 30 --warm 20 --dict-cap 128"""
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
a = ap.parse_args()
env = create_batched_env(20260000, 0, a.envs)
agent = create_kbrl_agent(np.random.default_rng(0), 0, accuracy_range=(0.97, 0.99), n_envs=a.envs, dict_cap=a.dict_cap)
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
sizes, flags = agent.learners.sizes()
tot = t_env + t_upd + t_sel
print(json.dumps({"workload": "scenario_0 + KBRL in the loop", "envs": a.envs, "steps": a.steps, "env_steps_per_s": a.envs * a.steps / tot,
                  "ms_env": 1e3 * t_env / a.steps, "ms_update_control": 1e3 * t_upd / a.steps, "ms_select_action": 1e3 * t_sel / a.steps,
                  "dict_mean": float(sizes.mean()), "dict_max": int(sizes.max()), "cap_hits": int((flags & 1).sum()),
                  "updates_last_step": agent.learners.counters()[1], "after_steps": a.warm + a.steps}))
