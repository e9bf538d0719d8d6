#!/usr/bin/env python3
"""Golden fixtures for kernel #2 (KBRL: GaussianKernel + Projectron + KBRL_Control) from the UNMODIFIED
reference.  Build-container tool.  For each case it runs the reference controller on the reference env
(Philox-injected, tests/refharness.py) and records, per step, everything the controller saw and decided,
so that the K3 tests can replay update_control / select_action without the env:
  state[t], action[t], labels[t], new_state[t], next_action[t], adjusted[t], hits[t], sizes[t] (dictionary
  size per learner after the step), security_factors[t], margins[t]; final landmarks/coeff/Kinv per learner.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {"K_scn0": (0, 9000, 400, [0.97, 0.99]), "K_scn1": (1, 9100, 300, [0.99, 0.999]),
         # the reference's horizon: dictionaries of several hundred landmarks (experiments_kbrl.py runs 50 400 steps)
         "K_long0": (0, 9200, 3000, [0.97, 0.99]),
         # scripted states (no env) that put whole dictionaries out of reach: f == 0 exactly, so the random tie-break of
         # GaussianKernel.predict (kernel.py:26-27) fires in update_control and in the select_action scan
         "K_tie": (0, 9300, 80, [0.97, 0.99]),
         # the learners of create_kbrl_agent swapped for ProjectronPlus (algorithms/projectron.py:66-107; never instantiated by
         # the reference's own factories): scripted states
         "K_plus": (0, 9400, 150, [0.97, 0.99])}
FULL_KINV_MAX_D = 160      # K_long: K^-1 is stored in full only for dictionaries up to this size (diag + row sums for all)


class _Ctx:
    learner = 0


def _route_tie_break(ref, agent, seed, tie_calls):
    """np.random.choice([-1, 1]) (kernel.py:27) -> Philox stream (seed, learner, STREAM_KBRL) of the learner predicting."""
    import types
    from ranslice_b200 import philox as px
    streams = [px.PhiloxStream(seed, s, px.STREAM_KBRL) for s in range(len(agent.learners))]

    def choice(seq):
        tie_calls.append(_Ctx.learner)
        return seq[streams[_Ctx.learner].integers(len(seq))]

    ref.kernel.np = types.SimpleNamespace(exp=np.exp, ndim=np.ndim, array=np.array, float32=np.float32, sign=np.sign,
                                          random=types.SimpleNamespace(choice=choice))
    for s, h in enumerate(agent.learners):
        orig = h.algorithm.kernel.predict

        def predict(x, orig=orig, s=s):
            _Ctx.learner = s
            return orig(x)

        h.algorithm.kernel.predict = predict


class _ScriptedEnv:
    """Stand-in for the env in K_tie: scripted observations / labels; every 7th step far away from everything seen."""

    def __init__(self, seed, V, S):
        self.rng, self.V, self.S, self.t = np.random.default_rng(seed), V, S, 0

    def reset(self):
        return np.zeros(self.V, np.float32)

    def step(self, action):
        self.t += 1
        obs = self.rng.random(self.V).astype(np.float32)
        if self.t % 7 in (3, 4):
            obs = (obs + np.float32(40.0 + self.t)).astype(np.float32)
        labels = np.where(self.rng.random(self.S) < 0.5, 1, -1)
        return obs, 0.0, False, {"SLA_labels": labels, "total_violations": int((labels < 0).sum())}


def gen(name):
    import refharness as rh
    scn, seed, steps, a_range = CASES[name]
    ref = rh.load_reference()
    tie_calls = []
    plain = ref.scenario_creator.Projectron
    if name == "K_plus":
        ref.scenario_creator.Projectron = ref.projectron.ProjectronPlus
    agent = ref.scenario_creator.create_kbrl_agent(np.random.default_rng(seed), scn, accuracy_range=a_range)
    ref.scenario_creator.Projectron = plain
    S = agent.n_slices
    _route_tie_break(ref, agent, seed, tie_calls)
    env = _ScriptedEnv(seed, 10 * S, S) if name in ("K_tie", "K_plus") else rh.make_env_philox(seed, scn)[0]
    init_action = agent.action.copy()
    init_sec = agent.security_factors.copy()
    rec = {k: [] for k in ("state", "action", "labels", "new_state", "next_action", "adjusted", "hits", "sizes",
                           "security_factors", "margins", "reward", "violations")}
    action = agent.action
    state = env.reset()
    for i in range(steps):
        new_state, reward, _, info = env.step(action)
        labels = info["SLA_labels"]
        hits = agent.update_control(state, action, labels)
        rec["state"].append(np.array(state, np.float32)); rec["action"].append(np.array(action, np.int64))
        rec["labels"].append(np.array(labels, np.int64)); rec["new_state"].append(np.array(new_state, np.float32))
        rec["hits"].append(np.array(hits, np.int64))
        rec["sizes"].append(np.array([h.algorithm.counter for h in agent.learners], np.int64))
        action, agent.adjusted = agent.select_action(new_state)
        rec["next_action"].append(np.array(action, np.int64)); rec["adjusted"].append(int(agent.adjusted))
        rec["security_factors"].append(agent.security_factors.astype(np.int64).copy())
        rec["margins"].append(np.asarray(agent.margins, np.int64).copy())
        rec["reward"].append(float(reward)); rec["violations"].append(int(info["total_violations"]))
        state = new_state
    out = {k: np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v) for k, v in rec.items()}
    Dmax = int(out["sizes"].max())
    dims = [h.indexes.stop - h.indexes.start + 1 for h in agent.learners]
    lm = np.zeros((S, Dmax, max(dims)))
    cf = np.zeros((S, Dmax))
    ki = np.zeros((S, Dmax, Dmax))
    for s, h in enumerate(agent.learners):
        D = h.algorithm.counter
        if D == 0:
            continue
        L = np.atleast_2d(np.asarray(h.algorithm.sv.landmarks, np.float64))
        lm[s, :D, :dims[s]] = L
        cf[s, :D] = np.asarray(h.algorithm.sv.coeff, np.float64)
        ki[s, :D, :D] = np.atleast_2d(np.asarray(h.algorithm.Kinv, np.float64))
    extra = {}
    if Dmax > FULL_KINV_MAX_D:         # keep the fixture small: full K^-1 only for the small dictionaries
        extra["kinv_diag"] = np.stack([np.diag(ki[s]) for s in range(S)])
        extra["kinv_rowsum"] = ki.sum(axis=2)
        small = [s for s in range(S) if out["sizes"][-1][s] <= FULL_KINV_MAX_D]
        extra["kinv_full_learners"] = np.array(small, np.int64)
        ki = ki[small][:, :FULL_KINV_MAX_D, :FULL_KINV_MAX_D] if small else np.zeros((0, 1, 1))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), scenario=scn, seed=seed, accuracy_range=np.array(a_range),
                        init_action=init_action.astype(np.int64), init_sec=init_sec.astype(np.int64), dims=np.array(dims),
                        n_prbs=agent.n_prbs, alfa=agent.alfa, tie_calls=len(tie_calls), final_landmarks=lm,
                        final_coeff=cf, final_kinv=ki, accuracies=agent.accuracies, tie_learners=np.array(tie_calls, np.int64),
                        numpy_version=np.__version__, **extra, **out)
    return name, out["sizes"][-1].tolist(), int(out["violations"].sum()), len(tie_calls)


if __name__ == "__main__":
    from concurrent.futures import ProcessPoolExecutor
    names = sys.argv[1:] or list(CASES)
    with ProcessPoolExecutor(4) as ex:
        for r in ex.map(gen, names):
            print(r, flush=True)
