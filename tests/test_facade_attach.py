"""The reference's own callers attach UNCHANGED to the product's single-env facade:
``wrapper.ReportWrapper`` / ``DQNWrapper`` and ``KBRL_Control.run``.  The reference tree only exists in
the build container (no GPU there), so the facade is exercised on a CPU stand-in backend with the
``BatchedRanSlice`` interface (the oracle); the same facade class wraps the CUDA batch on the GPU box
(tests/test_gpu_parity.py::test_single_env_facade_matches_reference_surface)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.needs_reference


class OracleBackend:
    """Duck-typed BatchedRanSlice (n_envs = 1) on the CPU oracle."""

    def __init__(self, tables, scenario, seed, penalty=100):
        self.env = ol.OracleEnv(tables, scenario, seed, penalty=penalty)
        self.n_envs, self.n_prbs, self.n_slices, self.n_variables = 1, self.env.n_prbs, self.env.S, self.env.V
        self.n_embb = {0: 5, 1: 3, 2: 1, 3: 1}[scenario]
        self.slots_per_step, self.penalty = 50, penalty
        self._acc = None
        self._prbs = None

    def reset(self):
        return self.env.reset()[None]

    def step(self, action):
        obs, rew, lab, vio, acc, flags = self.env.step(np.asarray(action).reshape(-1))
        self._acc, self._prbs = acc, np.asarray(action).reshape(-1).astype(np.int32)
        info = {"SLA_labels": lab[None], "violations": vio[None], "total_violations": vio.sum()[None],
                "flags": np.array([flags], np.uint32)}
        return obs[None], np.array([rew], np.float32), False, info

    def get_info(self, env=0):
        return self._acc, self._prbs


def test_report_wrapper_and_kbrl_run_attach_unchanged(tables, tmp_path):
    import refharness as rh
    from ranslice_b200.ran_slice import RanSlice
    ref = rh.load_reference()
    env = RanSlice(OracleBackend(tables, 0, 321))
    assert (env.n_prbs, env.n_slices, env.n_variables) == (200, 5, 50)
    # wrapper.ReportWrapper: float simplex action -> integer PRBs, obs clip/shift, history buffers
    w = ref.wrapper.ReportWrapper(env, steps=20, control_steps=10, env_id=1, path=str(tmp_path) + "/", verbose=False)
    obs = w.reset()
    rng = np.random.default_rng(0)
    for _ in range(20):
        obs, reward, done, info = w.step(rng.random(6))
        assert obs.shape == (50,) and obs.min() >= -1.0 and obs.max() <= 1.0 and done is False
    assert w.action_history[:20].max() <= 200
    # KBRL_Control.run on a fresh facade
    env2 = RanSlice(OracleBackend(tables, 1, 654))
    agent = ref.scenario_creator.create_kbrl_agent(np.random.default_rng(1), 1)
    out = agent.run(env2, 15)
    assert set(out) == {"reward", "resources", "hits", "adjusted", "SLA", "violation"}
    assert out["hits"].shape == (5, 15)
