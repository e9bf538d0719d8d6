"""CPU restatement (numpy) of the reference's wrapper arithmetic -- TEST INFRASTRUCTURE ONLY.

Oracle for the batched wrappers (include/wrapper_b200.h, csrc/wrappers.cu).  Only tests/ may import this module;
the product never does.  Pinned bit-exact against tests/golden/W_wrappers.npz, which tools/make_golden_wrapper.py
generated from the UNMODIFIED reference classes (wrapper.py ReportWrapper / DQNWrapper).
"""
from itertools import product

import numpy as np


def map_action(action, n_prbs):
    """wrapper.py:77-82 for one env: ``action`` float array [S+1] -> int64 [S]."""
    n_slices = len(action) - 1
    action = abs(action)
    t_action = action.sum()
    if t_action == 0:
        t_action = 1
    return np.array([np.floor(n_prbs * action[i] / t_action) for i in range(n_slices)], dtype=np.int64)


def normalize_obs(obs):
    """wrapper.py:88-90."""
    obs = np.clip(obs, -0.5, 1.5)
    return obs - 0.5


def dqn_table(n_prbs, n_slices=2, g_eMBB=2, max_eMBB=51):
    """wrapper.py:141-149 (the reference hard-codes the product over two slices)."""
    a = list(range(0, max_eMBB, g_eMBB))
    return np.array([c for c in product(a, repeat=n_slices) if sum(c) <= n_prbs], dtype=np.int64)


def histories(violations, reward, prbs):
    """wrapper.py:101-106 over T steps of one env: violations [T,S], reward [T], prbs [T,S]."""
    return (violations.sum(axis=1).astype(np.int16), reward.astype(np.float64), prbs.sum(axis=1).astype(np.int16))
