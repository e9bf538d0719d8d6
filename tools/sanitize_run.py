#!/usr/bin/env python3
"""Smoke-size workload for compute-sanitizer (memcheck / racecheck / initcheck) over K1 (all routes), K2, K3, K4:
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
Small enough for racecheck's ~100x slowdown; shrunk route limits make the pair-of-lanes and abort-and-replay paths run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ranslice_b200 import create_batched_env  # noqa: E402
from ranslice_b200.kbrl import create_kbrl_agent  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(0)
for scn, S, n_prbs, kw in ((0, 5, 200, {}), (0, 5, 200, {"routes": (2, 2, 5, 6)}), (3, 2, 70, {}), (0, 1, 200, {"L1_level": False})):
    N = 96
    routes = kw.pop("routes", None)
    env = create_batched_env(7, scn, N, **kw)
    if routes:
        env.set_route_limits(*routes)
    env.reset()
    for t in range(steps):
        w = rng.random((N, S + 1))
        env.step(np.floor(n_prbs * w[:, :S] / w.sum(axis=1, keepdims=True)).astype(np.int32))
    print("scenario", scn, kw, routes, "routes of the last step", env.routes(), "live UEs", int(env.n_ues().sum()))
    env.close()
env = create_batched_env(11, 0, 32)
for resident in (False, True):
    agent = create_kbrl_agent(np.random.default_rng(1), 0, accuracy_range=(0.97, 0.99), n_envs=32, resident=resident, dict_cap=64)
    agent.run(env, steps)
    print("kbrl resident" if resident else "kbrl host", "sizes max", int(agent.learners.sizes()[0].max()), agent.learners.pool())
env.close()
print("sanitize_run done")
