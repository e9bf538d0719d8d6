"""Batched wrappers: numpy oracle vs the fixture recorded from the unmodified reference wrapper.py
(CPU), CUDA kernels vs fixture and oracle through the C ABI (GPU), result-file schema."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import wrapper_oracle as wo  # noqa: E402


# ------------------------------------------------------------------------------------------------ CPU: oracle vs reference fixture
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_oracle_matches_reference_wrapper_fixture(golden, dt):
    g = golden("W_wrappers")
    n_prbs = int(g["n_prbs"])
    act, prbs, obs, nobs = g["act_" + dt], g["prbs_" + dt], g["obs_" + dt], g["nobs_" + dt]
    N, T = act.shape[:2]
    for e in range(N):
        for t in range(T):
            assert np.array_equal(wo.map_action(act[e, t].copy(), n_prbs), prbs[e, t]), (e, t)
        assert np.array_equal(wo.normalize_obs(obs[e]), nobs[e])
        vh, rh, ah = wo.histories(g["viol_" + dt][e], g["reward_" + dt][e], prbs[e])
        assert np.array_equal(vh, g["vh_" + dt][e]) and np.array_equal(rh, g["rh_" + dt][e]) and np.array_equal(ah, g["ah_" + dt][e])
    assert (prbs.sum(axis=2) <= n_prbs).all()
    assert np.array_equal(wo.dqn_table(int(g["dqn_n_prbs"])), g["dqn_table"])
    assert np.array_equal(g["dqn_table"][g["dqn_index"]], g["dqn_prbs"])


def test_kbrl_result_files_have_the_reference_schema(tmp_path):
    from ranslice_b200.wrapper import save_kbrl_results
    N, S, T = 3, 5, 7
    res = {"reward": np.arange(N * T, dtype=float).reshape(N, T), "resources": np.ones((N, T), np.int16),
           "hits": np.ones((N, S, T), np.int16), "adjusted": np.zeros((N, T), np.int16), "SLA": np.zeros((N, T), np.int16),
           "violation": np.zeros((N, T), np.int16)}
    files = save_kbrl_results(res, str(tmp_path), first_run=4)
    assert [os.path.basename(f) for f in files] == ["results_4.npz", "results_5.npz", "results_6.npz"]
    h = np.load(files[1])
    assert set(h.files) == {"reward", "resources", "hits", "adjusted", "SLA", "violation"}      # kbrl_control.py:148-155
    assert h["hits"].shape == (S, T) and h["violation"].shape == (T,) and np.array_equal(h["reward"], res["reward"][1])
    assert np.mean(h["hits"], axis=0).shape == (T,)                                             # plot_results.py:73


# ------------------------------------------------------------------------------------------------ GPU
class _FakeEnv:
    def __init__(self, N, S, n_prbs, V):
        self.n_envs, self.n_slices, self.n_prbs, self.n_variables, self.device = N, S, n_prbs, V, 0


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_cuda_wrappers_match_reference_fixture(golden, dt):
    import ctypes as C
    import torch
    from ranslice_b200 import _lib
    from ranslice_b200.wrapper import BatchedReportWrapper, _bind, _ptr
    g = golden("W_wrappers")
    n_prbs = int(g["n_prbs"])
    act, prbs, obs, nobs = g["act_" + dt], g["prbs_" + dt], g["obs_" + dt], g["nobs_" + dt]
    viol, reward = g["viol_" + dt], g["reward_" + dt]
    N, T, S1 = act.shape
    S, V = S1 - 1, obs.shape[2]
    w = BatchedReportWrapper(_FakeEnv(N, S, n_prbs, V), steps=T)
    L = _bind(_lib.lib())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for t in range(T):
        p = w._to_prbs(act[:, t].copy())                                  # numpy in, dtype kept (float32 / float64)
        assert np.array_equal(p.cpu().numpy(), prbs[:, t]), t
        o = torch.from_numpy(obs[:, t].copy()).cuda()
        out = torch.empty_like(o)
        _lib.check(L.rs_wrap_obs_device(_ptr(o), _ptr(out), N * V, st))
        assert np.array_equal(out.cpu().numpy(), nobs[:, t]), t
        v = torch.from_numpy(viol[:, t].astype(np.int32)).cuda()
        r = torch.from_numpy(reward[:, t].copy()).cuda()
        _lib.check(L.rs_wrap_record_device(_ptr(v), _ptr(r), _ptr(p), N, S, t, _ptr(w._violation), _ptr(w._reward),
                                           _ptr(w._action), st))
    assert np.array_equal(w.violation_history, g["vh_" + dt])
    assert np.array_equal(w.reward_history, g["rh_" + dt])
    assert np.array_equal(w.action_history, g["ah_" + dt])


@pytest.mark.gpu
def test_cuda_action_mapping_matches_oracle_at_scale_and_edges():
    from ranslice_b200.wrapper import BatchedReportWrapper
    rng = np.random.default_rng(9)
    for S, n_prbs in ((5, 200), (2, 70), (8, 150)):                        # 8 slices: numpy's 8-way pairwise sum path
        N = 20000
        for dt in (np.float32, np.float64):
            a = rng.random((N, S + 1)).astype(dt)
            a[::11] *= -1; a[5] = 0; a[7, :S] = 0; a[9] = 1e-30; a[13] = dt(1) / dt(3)
            w = BatchedReportWrapper(_FakeEnv(N, S, n_prbs, 10 * S), steps=1)
            got = w._to_prbs(a.copy()).cpu().numpy()
            want = np.stack([wo.map_action(a[e].copy(), n_prbs) for e in range(N)])
            assert np.array_equal(got, want), (S, dt)
            assert (got.sum(axis=1) <= n_prbs).all() and (got >= 0).all()


@pytest.mark.gpu
def test_dqn_table_and_vecenv_over_the_real_env(golden, tmp_path):
    """DQN table == the reference's; wrappers + VecEnvAdapter drive the native batched env; per-env history files
    carry the reference's keys (wrapper.py:120-123) and equal what the env reported."""
    from ranslice_b200 import create_batched_env
    from ranslice_b200.wrapper import BatchedDQNWrapper, BatchedReportWrapper, VecEnvAdapter
    g = golden("W_wrappers")
    N = 16
    env3 = create_batched_env(5, 3, N)
    dq = BatchedDQNWrapper(env3, steps=12, control_steps=6, path=str(tmp_path) + "/dqn/")
    assert np.array_equal(np.stack(dq.actions).astype(np.int64), g["dqn_table"])
    dq.reset()
    rng = np.random.default_rng(0)
    for t in range(12):
        idx = rng.integers(0, len(dq.actions), N)
        obs, rew, done, info = dq.step(idx)
        assert np.array_equal(dq._prbs.cpu().numpy(), g["dqn_table"][idx])
        assert float(obs.min()) >= -1.0 and float(obs.max()) <= 1.0 and done is False and info[0] == 0 and not bool(info["flags"].any())
    h = np.load(str(tmp_path) + "/dqn/history_%d.npz" % (dq.env_id + 3))
    assert set(h.files) == {"violation", "reward", "resources"} and h["violation"].shape == (12,)
    assert np.array_equal(h["resources"], dq.action_history[3])
    env3.close()

    env0 = create_batched_env(6, 0, N)
    ref_env = create_batched_env(6, 0, N)                                    # same seeds, driven without the wrapper
    rw = BatchedReportWrapper(env0, steps=10, control_steps=1000, path=str(tmp_path) + "/rw/")
    venv = VecEnvAdapter(rw)
    assert venv.num_envs == N and venv.reset().shape == (N, 50)
    ref_env.reset()
    for t in range(10):
        a = rng.random((N, 6)).astype(np.float32)
        obs, rew, dones, infos = venv.step(a)
        prbs = np.stack([wo.map_action(a[e].copy(), 200) for e in range(N)])
        o2, r2, _, i2 = ref_env.step(prbs)
        assert np.array_equal(obs, wo.normalize_obs(o2)) and np.array_equal(rew, r2) and not dones.any()
        assert np.array_equal(rw.violation_history[:, t], i2["total_violations"].astype(np.int16))
        assert np.array_equal(rw.action_history[:, t], prbs.sum(axis=1).astype(np.int16))
    env0.close(); ref_env.close()
