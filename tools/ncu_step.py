#!/usr/bin/env python3
"""Tiny driver for ncu captures: burn in, then run a few device-path steps.
    ncu ... python tools/ncu_step.py --envs 16384 --burn-in 100 --steps 3 [--variant V] [--scenario S]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ranslice_b200 import create_batched_env  # noqa: E402

SCN = {0: (5, 200), 1: (5, 150), 2: (5, 100), 3: (2, 70)}
ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--burn-in", type=int, default=100)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--scenario", type=int, default=0)
a = ap.parse_args()
S, n_prbs = SCN[a.scenario]
env = create_batched_env(20260000, a.scenario, a.envs, kernel_variant=a.variant)
env.reset()
rng = np.random.default_rng(0)
acts = []
for i in range(8):
    w = rng.random((a.envs, S + 1), dtype=np.float32)
    acts.append(torch.from_numpy(np.floor(n_prbs * w[:, :S] / w.sum(axis=1, keepdims=True)).astype(np.int32)).cuda())
out = None
for i in range(a.burn_in + a.steps):
    out = env.step_device(acts[i % 8], out)
torch.cuda.synchronize()
print("done", float(out["reward"].mean()), env.n_ues().mean())
