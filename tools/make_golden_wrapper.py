#!/usr/bin/env python3
"""Golden fixture for the batched wrappers (tests/golden/W_wrappers.npz) from the UNMODIFIED reference wrapper.py.
Build-container tool.  ReportWrapper / DQNWrapper are driven over a scripted stub env (the wrappers' arithmetic does
not depend on what is inside the env): recorded per step are the float action an agent would emit, the PRB action
the wrapper handed to env.step (captured by the stub), the raw obs / reward / violations the stub returned, the
normalised obs the wrapper returned, and at the end the wrapper's three history buffers.  Several independent
"envs" (rows) are recorded so that the batched implementation can be checked row by row."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "W_wrappers.npz")


class StubEnv:
    """Scripted env with the attribute surface ReportWrapper reads (wrapper.py:39-42,77,82)."""

    def __init__(self, n_slices, n_prbs, n_variables, obs, reward, violations):
        self.n_slices, self.n_prbs, self.n_variables = n_slices, n_prbs, n_variables
        self._obs, self._reward, self._viol = obs, reward, violations
        self.t = 0
        self.seen = []

    def reset(self):
        self.t = 0
        return np.zeros(self.n_variables, np.float32)

    def step(self, action):
        self.seen.append(np.array(action, np.int64))
        o, r, v = self._obs[self.t], self._reward[self.t], self._viol[self.t]
        self.t += 1
        return o, float(r), False, {"total_violations": int(v.sum()), "violations": v}


def main():
    import refharness as rh
    ref = rh.load_reference()
    W = ref.wrapper
    rng = np.random.default_rng(4242)
    N, T, S, n_prbs, V = 24, 40, 5, 200, 50
    out = {"n_prbs": n_prbs, "numpy_version": np.__version__}
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        act = rng.random((N, T, S + 1)).astype(dt)
        act[:, ::7] *= -1                                    # negative entries: abs()
        act[0, 3] = 0                                        # all-zero action: t_action = 1
        act[1, 5, :S] = 0                                    # only the slack weight non-zero
        act[2, :, :] = (rng.random((T, S + 1)) * 1e-3).astype(dt)
        obs = (rng.random((N, T, V)) * 3 - 1).astype(np.float32)   # spans both clip bounds
        viol = (rng.random((N, T, S)) < 0.2).astype(np.int64)
        reward = np.where(viol.sum(axis=2) > 0, -100.0 * viol.sum(axis=2), rng.integers(0, 200, (N, T))).astype(np.float32)
        prbs = np.zeros((N, T, S), np.int64)
        nobs = np.zeros((N, T, V), np.float32)
        vh = np.zeros((N, T), np.int16); rh_ = np.zeros((N, T)); ah = np.zeros((N, T), np.int16)
        for e in range(N):
            stub = StubEnv(S, n_prbs, V, obs[e], reward[e], viol[e])
            w = W.ReportWrapper(stub, steps=T, control_steps=10 ** 9, env_id=e, path="/tmp/")
            w.reset()
            for t in range(T):
                o, r, d, info = w.step(act[e, t].copy())
                nobs[e, t] = o
                assert r == float(reward[e, t]) and d is False and info == {0: 0}
            prbs[e] = np.stack(stub.seen)
            vh[e], rh_[e], ah[e] = w.violation_history, w.reward_history, w.action_history
        out.update({"act_" + name: act, "obs_" + name: obs, "viol_" + name: viol, "reward_" + name: reward,
                    "prbs_" + name: prbs, "nobs_" + name: nobs, "vh_" + name: vh, "rh_" + name: rh_, "ah_" + name: ah})
    # DQNWrapper: two slices (wrapper.py:141-149 hard-codes pairs), scenario_3 shape
    S2, n2, V2, T2 = 2, 70, 13, 30
    obs = (rng.random((T2, V2)) * 3 - 1).astype(np.float32)
    viol = (rng.random((T2, S2)) < 0.2).astype(np.int64)
    reward = rng.integers(-200, 70, T2).astype(np.float32)
    stub = StubEnv(S2, n2, V2, obs, reward, viol)
    w = W.DQNWrapper(stub, steps=T2, control_steps=10 ** 9, env_id=0, path="/tmp/")
    w.reset()
    idx = rng.integers(0, len(w.actions), T2)
    for t in range(T2):
        w.step(int(idx[t]))
    out.update({"dqn_table": np.stack(w.actions).astype(np.int64), "dqn_index": idx, "dqn_prbs": np.stack(stub.seen),
                "dqn_resources": w.action_history.copy(), "dqn_n_prbs": n2})
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
