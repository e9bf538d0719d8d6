"""Batched equivalents of the reference's gym wrappers (``wrapper.py``) for :class:`BatchedRanSlice`.

* :class:`BatchedReportWrapper` -- ``ReportWrapper`` (wrapper.py:26-134): continuous action mapping
  (``floor(n_prbs * |a_i| / sum|a|)``), observation normalisation ``clip(obs, -0.5, 1.5) - 0.5``, the
  ``violation / reward / resources`` history buffers and ``save_results`` / ``set_evaluation``.
* :class:`BatchedDQNWrapper` -- ``DQNWrapper`` (wrapper.py:136-154): the same with the discrete action table.
* :class:`BatchedTimerWrapper` -- ``TimerWrapper`` (wrapper.py:156-217): accumulates the simulation time.
* :class:`VecEnvAdapter` -- the ``VecEnv`` surface (``num_envs``, ``reset``, ``step_async`` / ``step_wait``,
  ``step``) model-free libraries drive, with numpy in / out (the reference wraps ONE env with
  ``make_vec_env(lambda: env, n_envs=1)``, experiments_rl.py:95).

All arithmetic runs on the GPU through ``include/wrapper_b200.h`` (``rs_wrap_*``): ``step`` takes a torch CUDA
tensor (or a numpy array, copied once) and returns CUDA tensors; nothing per env happens in Python.
``save_results`` writes the reference's ``.npz`` schema (keys ``violation, reward, resources``,
wrapper.py:120-123) -- one ``history_<env_id>.npz`` per env, so that the reference's ``plot_results.py`` reads the
files unchanged, or a single batched file with a leading env axis.
"""
import ctypes as C
import os
from itertools import product

import numpy as np

from . import _lib


def _bind(L):
    if getattr(L, "_rw_bound", False):
        return L
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.rs_wrap_action_device.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.rs_wrap_dqn_action_device.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    L.rs_wrap_obs_device.argtypes = [vp, vp, i64, vp]
    L.rs_wrap_record_device.argtypes = [vp, vp, vp, i32, i32, i64, vp, vp, vp, vp]
    for n in ("rs_wrap_action_device", "rs_wrap_dqn_action_device", "rs_wrap_obs_device", "rs_wrap_record_device"):
        getattr(L, n).restype = C.c_int
    L._rw_bound = True
    return L


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class _Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class _Discrete:
    def __init__(self, n):
        self.n = n


class BatchedReportWrapper:
    """``ReportWrapper`` for N envs.  ``step(action)``: action float32/float64 ``[N, S+1]`` (simplex weights, mapped
    like wrapper.py:77-82) or integer ``[N, S]`` (PRBs, passed through) -> ``(obs, reward, done, info)`` with
    ``obs`` normalised, as CUDA tensors; ``info`` is ``{0: 0}`` like the reference's (wrapper.py:117)."""

    def __init__(self, env, steps=2000, control_steps=500, env_id=1, extra_samples=10, path='./logs/', verbose=False,
                 per_env_files=True):
        import torch
        self.env = env
        self._L = _bind(_lib.lib())
        self.n_envs, self.n_slices, self.n_prbs, self.n_variables = env.n_envs, env.n_slices, env.n_prbs, env.n_variables
        self.device = torch.device("cuda", env.device)
        self.action_space = _Box(0, 1, (self.n_slices + 1,), np.float64)                 # wrapper.py:39-40
        self.observation_space = _Box(-1, 1, (self.n_variables,), np.float64)            # wrapper.py:41-42
        self.steps, self.step_counter, self.control_steps = steps, 0, control_steps
        self.env_id, self.verbose, self.path, self.extra_samples = env_id, verbose, path, extra_samples
        self.per_env_files = per_env_files
        self.file_prefix = 'history'
        self._out = None
        self._prbs = torch.zeros((self.n_envs, self.n_slices), dtype=torch.int32, device=self.device)
        self.obs = torch.zeros((self.n_envs, self.n_variables), dtype=torch.float32, device=self.device)
        self._flags = torch.zeros(self.n_envs, dtype=torch.int32, device=self.device)   # OR of the env's RS_FLAG_* bits over all steps
        self.reset_history()

    # ------------------------------------------------------------------ histories (wrapper.py:57-60)
    def reset_history(self):
        import torch
        n, dev = self.n_envs, self.device
        self._violation = torch.zeros((self.steps, n), dtype=torch.int16, device=dev)
        self._reward = torch.zeros((self.steps, n), dtype=torch.float64, device=dev)
        self._action = torch.zeros((self.steps, n), dtype=torch.int16, device=dev)

    @property
    def violation_history(self):
        return self._violation.transpose(0, 1).cpu().numpy()

    @property
    def reward_history(self):
        return self._reward.transpose(0, 1).cpu().numpy()

    @property
    def action_history(self):
        return self._action.transpose(0, 1).cpu().numpy()

    @property
    def flags_seen(self):
        """OR of the per-env deviation flags (RS_FLAG_* in include/ranslice_b200.h: UE / burst / mMTC backlog caps, clamped
        actions) over every step so far, uint32 [N].  Non-zero means that env left the reference's trajectory at a cap."""
        return self._flags.cpu().numpy().view(np.uint32)

    # ------------------------------------------------------------------ gym-like API
    def reset(self):
        self.step_counter = 0
        self.env.reset()
        self.obs.zero_()
        return self.obs

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _to_prbs(self, action):
        """wrapper.py:77-82: weights [N, S+1] -> PRBs [N, S]; integer [N, S] actions pass through."""
        import torch
        if not torch.is_tensor(action):
            action = torch.from_numpy(np.ascontiguousarray(action)).to(self.device)
        if action.shape[-1] > self.n_slices:
            if action.dtype not in (torch.float32, torch.float64):
                action = action.to(torch.float64)
            action = action.contiguous()
            assert tuple(action.shape) == (self.n_envs, self.n_slices + 1)
            _lib.check(self._L.rs_wrap_action_device(_ptr(action), int(action.dtype == torch.float64), self.n_envs,
                                                     self.n_slices, self.n_prbs, _ptr(self._prbs), self._stream()))
            return self._prbs
        assert tuple(action.shape) == (self.n_envs, self.n_slices)
        self._prbs.copy_(action)
        return self._prbs

    def step(self, action):
        prbs = self._to_prbs(action)
        out = self._out = self.env.step_device(prbs, self._out)
        st = self._stream()
        _lib.check(self._L.rs_wrap_obs_device(_ptr(out['obs']), _ptr(self.obs), self.n_envs * self.n_variables, st))
        if self.step_counter < self.steps:                                               # wrapper.py:103-106
            _lib.check(self._L.rs_wrap_record_device(_ptr(out['violations']), _ptr(out['reward']), _ptr(prbs), self.n_envs,
                                                     self.n_slices, self.step_counter, _ptr(self._violation),
                                                     _ptr(self._reward), _ptr(self._action), st))
        self._flags |= out['flags']
        self.step_counter += 1
        if self.step_counter % self.control_steps == 0:
            self.save_results()
        return self.obs, out['reward'], False, {0: 0, 'flags': out['flags']}        # (the reference's info is {0: 0})

    # ------------------------------------------------------------------ result files (wrapper.py:119-134)
    def _file(self, e):
        return '{}{}_{}.npz'.format(self.path, self.file_prefix, self.env_id + e)

    def save_results(self):
        """``np.savez(history_<id>.npz, violation=, reward=, resources=)`` per env (ids ``env_id .. env_id+N-1``), or
        one ``history_<id>_batch.npz`` with ``[N, steps]`` arrays when ``per_env_files`` is False."""
        os.makedirs(self.path, exist_ok=True)
        v, r, a = self.violation_history, self.reward_history, self.action_history
        if self.per_env_files:
            for e in range(self.n_envs):
                np.savez(self._file(e), violation=v[e], reward=r[e], resources=a[e])
        else:
            np.savez('{}{}_{}_batch.npz'.format(self.path, self.file_prefix, self.env_id), violation=v, reward=r, resources=a)

    def set_evaluation(self, eval_steps, new_path=None, change_name=False):
        import torch
        self.step_counter = self.steps
        self.steps += eval_steps
        pad = lambda t: torch.cat([t, torch.zeros((eval_steps, self.n_envs), dtype=t.dtype, device=t.device)])
        self._violation, self._reward, self._action = pad(self._violation), pad(self._reward), pad(self._action)
        if new_path:
            self.path = new_path
        if change_name:
            self.file_prefix = 'evaluation'


class BatchedDQNWrapper(BatchedReportWrapper):
    """``DQNWrapper`` (wrapper.py:136-154): ``step(index)`` with ``index`` int ``[N]`` into the action table built
    like wrapper.py:141-149 (granularity 2, at most 50 PRBs per slice, ``sum <= n_prbs``; the reference hard-codes
    two slices, here the product runs over ``n_slices``)."""

    def __init__(self, env, g_eMBB=2, max_eMBB=51, **kw):
        import torch
        super().__init__(env, **kw)
        a = list(range(0, max_eMBB, g_eMBB))
        self.actions = [np.array(c, dtype=np.int16) for c in product(a, repeat=self.n_slices) if sum(c) <= self.n_prbs]
        self.action_space = _Discrete(len(self.actions))
        self._table = torch.from_numpy(np.asarray(self.actions, np.int32)).to(self.device).contiguous()
        self._index = torch.zeros(self.n_envs, dtype=torch.int32, device=self.device)
        self.flags = torch.zeros(self.n_envs, dtype=torch.int32, device=self.device)

    def step(self, action):
        import torch
        if not torch.is_tensor(action):
            action = torch.from_numpy(np.ascontiguousarray(action))
        self._index.copy_(action.reshape(self.n_envs))
        _lib.check(self._L.rs_wrap_dqn_action_device(_ptr(self._index), _ptr(self._table), self.n_envs, self.n_slices,
                                                     len(self.actions), _ptr(self._prbs), _ptr(self.flags), self._stream()))
        return super().step(self._prbs)


class BatchedTimerWrapper(BatchedReportWrapper):
    """``TimerWrapper`` (wrapper.py:156-217): same mapping / normalisation, no histories; ``get_simtime`` returns the
    accumulated device time of the env steps in seconds (CUDA events; the reference's wall-clock difference is
    accumulated with the wrong sign, wrapper.py:199-201 -- not replicated)."""

    def __init__(self, env, steps=2000):
        super().__init__(env, steps=steps, control_steps=1 << 62)
        self.simtime = 0.0
        self._ev = []

    def reset(self):
        self.simtime, self._ev = 0.0, []
        return super().reset()

    def step(self, action):
        import torch
        prbs = self._to_prbs(action)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = self._out = self.env.step_device(prbs, self._out)
        e1.record()
        self._ev.append((e0, e1))
        _lib.check(self._L.rs_wrap_obs_device(_ptr(out['obs']), _ptr(self.obs), self.n_envs * self.n_variables, self._stream()))
        self.step_counter += 1
        return self.obs, out['reward'], False, {0: 0}

    def get_simtime(self):
        import torch
        torch.cuda.synchronize(self.device)
        self.simtime += sum(a.elapsed_time(b) for a, b in self._ev) * 1e-3
        self._ev = []
        return self.simtime


class VecEnvAdapter:
    """Vectorised-env surface over a batched wrapper: numpy in / out, ``dones`` all False (infinite horizon,
    ran_slice.py:50), one ``info`` dict per env."""

    def __init__(self, wrapped):
        self.venv = wrapped
        self.num_envs = wrapped.n_envs
        self.observation_space, self.action_space = wrapped.observation_space, wrapped.action_space
        self._pending = None

    def reset(self):
        return self.venv.reset().cpu().numpy()

    def step_async(self, actions):
        self._pending = self.venv.step(np.asarray(actions))

    def step_wait(self):
        obs, rew, _, _ = self._pending
        self._pending = None
        return obs.cpu().numpy(), rew.cpu().numpy(), np.zeros(self.num_envs, bool), [{} for _ in range(self.num_envs)]

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.venv.env.close()


def save_kbrl_results(results, path, first_run=0):
    """``KBRL_Control.run`` histories with a leading env axis (ranslice_b200.kbrl) -> one ``results_<run>.npz`` per env
    with the reference's keys and shapes (kbrl_control.py:148-155, experiments_kbrl.py:60-61: ``reward, resources,
    adjusted, SLA, violation`` ``[steps]``, ``hits`` ``[n_slices, steps]``), readable by plot_results.py:59-80."""
    os.makedirs(path, exist_ok=True)
    n = len(results['reward'])
    files = []
    for e in range(n):
        f = os.path.join(path, 'results_{}.npz'.format(first_run + e))
        np.savez(f, **{k: np.asarray(v[e]) for k, v in results.items()})
        files.append(f)
    return files
