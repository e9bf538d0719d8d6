"""Host-side mirror of the reference's KBRL controller on top of kernel #2 (include/kbrl_b200.h).

* :class:`BatchedProjectron` -- L = n_envs * n_slices Projectron learners resident on the GPU; the two
  batched calls replace the per-sample ``Projectron.predict / update`` loops of ``KBRL_Control``
  (kbrl_control.py:54-61 and :88-89,103-112).
* :class:`KBRLControl` -- ``KBRL_Control`` (kbrl_control.py:23-157) vectorised over the env axis with
  numpy: E-learner accuracies, security factors, margins, ``adjust_action`` and the ``run`` loop keep the
  reference's names, arithmetic and result keys (``reward, resources, hits, adjusted, SLA, violation``).
* :func:`create_kbrl_agent` -- ``scenario_creator.create_kbrl_agent`` (scenario_creator.py:197-238).
"""
import ctypes as C
import warnings

import numpy as np

from . import _lib
from .scenario_creator import scenarios, state_variables_embb, state_variables_mmtc

alfa = 0.05                 # scenario_creator.py:187
KBRL_HEAVY_THRESHOLD = 1800   # env routing under a KBRL policy (BatchedRanSlice.set_heavy_threshold; measured)
embb_sec, embb_a = (2, 8), (4, 20)        # :190-191
mmtc_sec, mmtc_a = (1, 4), (2, 10)        # :192-193


class KbConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n_envs", C.c_int32), ("n_slices", C.c_int32),
                ("n_prbs", C.c_int32), ("n_variables", C.c_int32), ("dict_cap", C.c_int32), ("pool_mb", C.c_int32),
                ("gamma", C.c_double), ("eta", C.c_double), ("tie_seed", C.c_uint64), ("first_env_id", C.c_uint64),
                ("algorithm", C.c_int32), ("reserved", C.c_int32)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bind(L):
    if getattr(L, "_kb_bound", False):
        return L
    vp = C.c_void_p
    L.kb_create.argtypes = [C.POINTER(KbConfig), vp, vp, C.POINTER(vp)]
    L.kb_destroy.argtypes = [vp]
    L.kb_reset.argtypes = [vp]
    L.kb_update.argtypes = [vp] * 5
    L.kb_predict.argtypes = [vp] * 3
    L.kb_update_device.argtypes = [vp] * 6
    L.kb_predict_device.argtypes = [vp] * 4
    L.kb_control_init.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.c_double]
    L.kb_control_update_device.argtypes = [vp] * 6
    L.kb_control_select_device.argtypes = [vp] * 5
    L.kb_control_get.argtypes = [vp] * 6
    L.kb_set_exact.argtypes = [vp, C.c_int32]
    L.kb_get_sizes.argtypes = [vp, vp, vp]
    L.kb_get_learner.argtypes = [vp, C.c_int32, vp, vp, vp, C.POINTER(C.c_int32)]
    L.kb_get_counters.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.kb_state_size.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.kb_get_state.argtypes = [vp, vp, C.c_size_t]
    L.kb_set_state.argtypes = [vp, vp, C.c_size_t]
    L.kb_get_pool.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
    for n in ("kb_create", "kb_destroy", "kb_reset", "kb_update", "kb_predict", "kb_update_device", "kb_predict_device",
              "kb_get_sizes", "kb_get_learner", "kb_get_counters", "kb_control_init", "kb_control_update_device",
              "kb_control_select_device", "kb_control_get", "kb_set_exact", "kb_get_pool", "kb_state_size", "kb_get_state", "kb_set_state"):
        getattr(L, n).restype = C.c_int
    L._kb_bound = True
    return L


def _report_flags(owner, env_flags, learners):
    """Deviation flags are never silent: the OR over the run is kept on the agent (``env_flags`` uint32 [N] of the env,
    ``learner_flags`` uint32 [N,S] of the dictionaries) and a warning names what fired."""
    owner.env_flags = np.asarray(env_flags).astype(np.uint32)
    owner.learner_flags = learners.sizes()[1]
    names = {1: 'UE cap', 2: 'burst cap', 4: 'action clamped', 8: 'same-slot departure', 16: 'mMTC backlog / arrival cap'}
    hit = [n for b, n in names.items() if (owner.env_flags & b).any()]
    if (owner.learner_flags & 1).any():
        hit.append('KBRL dictionary cap')
    if (owner.learner_flags & 2).any():
        hit.append('KBRL dictionary pool exhausted')
    if hit:
        warnings.warn('ranslice_b200: deviation flags raised during run(): %s (see agent.env_flags / agent.learner_flags)'
                      % ', '.join(hit))


class BatchedProjectron:
    """One ``Projectron(GaussianKernel(SVvariable(), gamma), eta)`` per (env, slice), on the GPU."""

    def __init__(self, scenario, n_envs, dict_cap=1024, device=0, gamma=1.0, eta=0.1, tie_seed=0, first_env_id=0,
                 pool_mb=0, algorithm="projectron"):
        """``dict_cap`` only bounds the size ONE dictionary may reach (the reference's are unbounded, flagged when hit);
        memory comes from a device pool of ``pool_mb`` MiB (0: automatic) as the dictionaries grow.  ``tie_seed`` keys
        the Philox stream behind the random tie-break of ``GaussianKernel.predict`` (kernel.py:26-27);
        ``algorithm``: "projectron" (projectron.py:23-60) or "projectron_plus" (:66-107)."""
        sc = scenarios[scenario] if isinstance(scenario, int) else scenario
        self.n_envs, self.n_prbs, self.device = n_envs, sc['n_prbs'], device
        n_embb, n_mmtc = sc['n_embb'], sc['n_mmtc']
        self.n_slices = n_embb + n_mmtc
        self.dims = np.array([len(state_variables_embb) + 1] * n_embb + [len(state_variables_mmtc) + 1] * n_mmtc, np.int32)
        self.offsets = np.concatenate([[0], np.cumsum(self.dims - 1)[:-1]]).astype(np.int32)
        self.n_variables = int((self.dims - 1).sum())
        L = _bind(_lib.lib())
        algo = {"projectron": 0, "projectron_plus": 1}[algorithm]
        cfg = KbConfig(_lib.RS_ABI_VERSION, device, n_envs, self.n_slices, self.n_prbs, self.n_variables, dict_cap, pool_mb,
                       gamma, eta, int(tie_seed) & (2 ** 64 - 1), first_env_id, algo, 0)
        h = C.c_void_p()
        _lib.check(L.kb_create(C.byref(cfg), _p(self.dims), _p(self.offsets), C.byref(h)))
        self._h, self._L = h, L

    def close(self):
        if getattr(self, "_h", None):
            self._L.kb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, state, action, labels):
        """-> y_pred int32 [N,S] (prediction at (state, action) before the augmentation updates)."""
        st = np.ascontiguousarray(state, np.float32).reshape(self.n_envs, self.n_variables)
        a = np.ascontiguousarray(action, np.int32).reshape(self.n_envs, self.n_slices)
        lab = np.ascontiguousarray(labels, np.int32).reshape(self.n_envs, self.n_slices)
        out = np.empty((self.n_envs, self.n_slices), np.int32)
        _lib.check(self._L.kb_update(self._h, _p(st), _p(a), _p(lab), _p(out)))
        return out

    def predict(self, state):
        """-> first_pos int32 [N,S]: smallest l1_prbs whose prediction is +1, -1 if none."""
        st = np.ascontiguousarray(state, np.float32).reshape(self.n_envs, self.n_variables)
        out = np.empty((self.n_envs, self.n_slices), np.int32)
        _lib.check(self._L.kb_predict(self._h, _p(st), _p(out)))
        return out

    def update_device(self, state, action, labels, y_pred):
        import torch
        stream = torch.cuda.current_stream(state.device).cuda_stream
        _lib.check(self._L.kb_update_device(self._h, C.c_void_p(state.data_ptr()), C.c_void_p(action.data_ptr()),
                                            C.c_void_p(labels.data_ptr()), C.c_void_p(y_pred.data_ptr()), C.c_void_p(stream)))
        return y_pred

    def predict_device(self, state, first_pos):
        import torch
        stream = torch.cuda.current_stream(state.device).cuda_stream
        _lib.check(self._L.kb_predict_device(self._h, C.c_void_p(state.data_ptr()), C.c_void_p(first_pos.data_ptr()),
                                             C.c_void_p(stream)))
        return first_pos

    def set_exact(self, on=True):
        """Validation: evaluate every f(x) in fp64 like the reference instead of the guarded fp32 fast path."""
        _lib.check(self._L.kb_set_exact(self._h, int(bool(on))))

    def sizes(self):
        s = np.empty((self.n_envs, self.n_slices), np.int32)
        f = np.empty((self.n_envs, self.n_slices), np.uint32)
        _lib.check(self._L.kb_get_sizes(self._h, _p(s), _p(f)))
        return s, f

    def learner(self, env, s):
        l = env * self.n_slices + s
        D = int(self.sizes()[0][env, s])
        d = int(self.dims[s])
        lm = np.zeros((max(D, 1), d)); cf = np.zeros(max(D, 1)); ki = np.zeros((max(D, 1), max(D, 1)))
        Dout = C.c_int32()
        _lib.check(self._L.kb_get_learner(self._h, l, _p(lm), _p(cf), _p(ki), C.byref(Dout)))
        return lm[:D], cf[:D], ki[:D, :D]

    def get_state(self):
        """Checkpoint of the learners (and of the device-resident controller, if any) as a uint8 array."""
        n = C.c_size_t()
        _lib.check(self._L.kb_state_size(self._h, C.byref(n)))
        blob = np.empty(n.value, np.uint8)
        _lib.check(self._L.kb_get_state(self._h, _p(blob), n))
        return blob

    def set_state(self, blob):
        blob = np.ascontiguousarray(blob, np.uint8)
        _lib.check(self._L.kb_set_state(self._h, _p(blob), C.c_size_t(blob.size)))

    def pool(self):
        """dict(used_bytes, total_bytes, max_dictionary, tie_breaks) of the dictionary pool."""
        u, t, m, tb = C.c_uint64(), C.c_uint64(), C.c_int32(), C.c_uint64()
        _lib.check(self._L.kb_get_pool(self._h, C.byref(u), C.byref(t), C.byref(m), C.byref(tb)))
        return {"used_bytes": int(u.value), "total_bytes": int(t.value), "max_dictionary": int(m.value),
                "tie_breaks": int(tb.value)}

    def counters(self):
        k, u = C.c_uint64(), C.c_uint64()
        _lib.check(self._L.kb_get_counters(self._h, C.byref(k), C.byref(u)))
        return int(k.value), int(u.value)


class KBRLControl:
    """``KBRL_Control`` (kbrl_control.py:23-157) for N envs at once; learners = :class:`BatchedProjectron`."""

    def __init__(self, learners, n_prbs, initial_action, security_factor, alfa=0.05, accuracy_range=(0.99, 0.999)):
        self.learners = learners
        self.accuracy_range = list(accuracy_range)
        self.n_envs, self.n_slices, self.n_prbs, self.alfa = learners.n_envs, learners.n_slices, n_prbs, alfa
        N, S = self.n_envs, self.n_slices
        self.adjusted = np.zeros(N, np.int64)
        self.action = np.broadcast_to(np.asarray(initial_action, np.int64), (N, S)).copy()
        self.security_factors = np.broadcast_to(np.asarray(security_factor, np.int64), (N, S)).copy()
        self.margins = np.zeros((N, S), np.int64)
        self.accuracies = np.full((N, S, n_prbs), (self.accuracy_range[0] + self.accuracy_range[1]) / 2, float)   # :38-39

    def select_action(self, state):
        """kbrl_control.py:41-78 -> (action int64 [N,S], adjusted int64 [N])."""
        n = self.n_prbs
        first = self.learners.predict(state).astype(np.int64)
        found = first >= 0
        a = np.minimum(n, first + self.security_factors)                     # :58
        action = np.where(found, a, n)                                       # loop ran out: l1_prbs = n_prbs
        margins = np.where(found, a - first, 0)
        assigned = action.sum(axis=1)
        adjusted = assigned > n                                              # :65
        rel = action / np.maximum(assigned, 1)[:, None]                      # adjust_action, :75-78
        new_action = np.floor(n * rel).astype(np.int64)
        self.margins = np.where(adjusted[:, None], margins - (action - new_action), margins)
        self.action = np.where(adjusted[:, None], new_action, action)
        return self.action.copy(), adjusted.astype(np.int64)

    def update_control(self, state, action, reward):
        """kbrl_control.py:80-114; ``reward`` = SLA labels (+1 / -1) [N,S] -> hits [N,S]."""
        action = np.asarray(action, np.int64)
        y = np.asarray(reward, np.int64)
        y_pred = self.learners.update(state, action, y).astype(np.int64)      # predict + sample augmentation on the GPU
        hit = y == y_pred
        margin = np.maximum(0, self.margins)
        idx = np.arange(self.n_prbs)[None, None, :]
        pos = (y_pred == 1)[:, :, None]
        miss_mask = pos & ~hit[:, :, None] & (idx <= margin[:, :, None])      # :93-94
        hit_mask = pos & hit[:, :, None] & (idx >= margin[:, :, None])        # :95-96
        acc = self.accuracies
        acc = np.where(miss_mask, (1 - self.alfa) * acc, acc)
        acc = np.where(hit_mask, (1 - self.alfa) * acc + self.alfa, acc)
        self.accuracies = acc
        sf = np.argmax(acc > self.accuracy_range[0], axis=2)                  # :98-99 (0 when none)
        free = (self.adjusted == 0)[:, None]
        self.security_factors = np.where(free, sf, self.security_factors)
        return hit.astype(np.int64)

    def run(self, system, steps, learning_time=-1):
        """kbrl_control.py:116-157 against a :class:`BatchedRanSlice`; histories get a leading env axis."""
        N, S = self.n_envs, self.n_slices
        action = self.action
        SLA_history = np.zeros((N, steps), np.int16)
        reward_history = np.zeros((N, steps), float)
        violation_history = np.zeros((N, steps), np.int16)
        adjusted_actions = np.zeros((N, steps), np.int16)
        resources_history = np.zeros((N, steps), np.int16)
        hits_history = np.zeros((N, S, steps), np.int16)
        state = system.reset()
        flags = np.zeros(N, np.uint32)
        for i in range(steps):
            new_state, reward, _, info = system.step(action)
            flags |= info['flags']
            SLA_labels = info['SLA_labels']
            hits = self.update_control(state, action, SLA_labels)
            action, self.adjusted = self.select_action(new_state)
            state = new_state
            SLA_history[:, i] = SLA_labels.sum(axis=1)
            reward_history[:, i] = reward
            violation_history[:, i] = info['total_violations']
            resources_history[:, i] = action.sum(axis=1)
            adjusted_actions[:, i] = self.adjusted
            hits_history[:, :, i] = hits
        _report_flags(self, flags, self.learners)
        return {'reward': reward_history, 'resources': resources_history, 'hits': hits_history,
                'adjusted': adjusted_actions, 'SLA': SLA_history, 'violation': violation_history}


class DeviceKBRLControl:
    """``KBRL_Control`` with ALL controller state resident on the GPU (kb_control_* in include/kbrl_b200.h):
    ``update_control`` / ``select_action`` take and return torch CUDA tensors and enqueue kernels on the current
    stream; ``run`` drives a :class:`BatchedRanSlice` through ``step_device`` with no host round trip per step."""

    def __init__(self, learners, n_prbs, initial_action, security_factor, alfa=0.05, accuracy_range=(0.99, 0.999)):
        import torch
        self.learners = learners
        self.accuracy_range = list(accuracy_range)
        self.n_envs, self.n_slices, self.n_prbs, self.alfa = learners.n_envs, learners.n_slices, n_prbs, alfa
        N, S = self.n_envs, self.n_slices
        ia = np.ascontiguousarray(np.broadcast_to(np.asarray(initial_action, np.int32), (N, S)))
        sec = np.ascontiguousarray(np.broadcast_to(np.asarray(security_factor, np.int32), (N, S)))
        self._L, self._h = learners._L, learners._h
        _lib.check(self._L.kb_control_init(self._h, _p(ia), _p(sec), float(alfa), float(accuracy_range[0]),
                                           float(accuracy_range[1])))
        self.device = torch.device("cuda", learners.device)
        self.action = torch.from_numpy(ia.copy()).to(self.device)
        self.adjusted = torch.zeros(N, dtype=torch.int32, device=self.device)

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def select_action(self, state, action_out=None, adjusted_out=None):
        """state float32 CUDA [N,V] -> (action int32 CUDA [N,S], adjusted int32 CUDA [N]) (kbrl_control.py:41-78)."""
        import torch
        assert state.is_cuda and state.dtype == torch.float32 and state.is_contiguous()
        a = action_out if action_out is not None else torch.empty((self.n_envs, self.n_slices), dtype=torch.int32, device=self.device)
        adj = adjusted_out if adjusted_out is not None else torch.empty(self.n_envs, dtype=torch.int32, device=self.device)
        _lib.check(self._L.kb_control_select_device(self._h, C.c_void_p(state.data_ptr()), C.c_void_p(a.data_ptr()),
                                                    C.c_void_p(adj.data_ptr()), self._stream()))
        self.action, self.adjusted = a, adj
        return a, adj

    def update_control(self, state, action, reward, hits_out=None):
        """(state f32 [N,V], action i32 [N,S], SLA labels i32 [N,S]) CUDA -> hits int32 CUDA [N,S] (kbrl_control.py:80-114)."""
        import torch
        for t, dt in ((state, torch.float32), (action, torch.int32), (reward, torch.int32)):
            assert t.is_cuda and t.dtype == dt and t.is_contiguous()
        hits = hits_out if hits_out is not None else torch.empty((self.n_envs, self.n_slices), dtype=torch.int32, device=self.device)
        _lib.check(self._L.kb_control_update_device(self._h, C.c_void_p(state.data_ptr()), C.c_void_p(action.data_ptr()),
                                                    C.c_void_p(reward.data_ptr()), C.c_void_p(hits.data_ptr()), self._stream()))
        return hits

    def control_state(self):
        """Host copies: dict(action, security_factors, margins [N,S], adjusted [N], accuracies [N,S,n_prbs])."""
        N, S = self.n_envs, self.n_slices
        out = dict(action=np.empty((N, S), np.int32), security_factors=np.empty((N, S), np.int32),
                   margins=np.empty((N, S), np.int32), adjusted=np.empty(N, np.int32),
                   accuracies=np.empty((N, S, self.n_prbs), np.float64))
        _lib.check(self._L.kb_control_get(self._h, _p(out["action"]), _p(out["security_factors"]), _p(out["margins"]),
                                          _p(out["adjusted"]), _p(out["accuracies"])))
        return out

    def run(self, system, steps, learning_time=-1):
        """kbrl_control.py:116-157 against a :class:`BatchedRanSlice`, everything on the device; the histories
        (leading env axis) are accumulated in HBM and copied to the host once at the end."""
        import torch
        N, S, dev = self.n_envs, self.n_slices, self.device
        i16 = dict(dtype=torch.int16, device=dev)
        SLA_history = torch.zeros((steps, N), **i16)
        reward_history = torch.zeros((steps, N), dtype=torch.float64, device=dev)
        violation_history = torch.zeros((steps, N), **i16)
        adjusted_actions = torch.zeros((steps, N), **i16)
        resources_history = torch.zeros((steps, N), **i16)
        hits_history = torch.zeros((steps, N, S), **i16)
        system.reset()
        if hasattr(system, 'set_heavy_threshold'):      # KBRL allocates just enough PRBs: many saturated slices with long PF loops
            system.set_heavy_threshold(KBRL_HEAVY_THRESHOLD)
        state = torch.zeros((N, system.n_variables), dtype=torch.float32, device=dev)      # reset() returns zeros
        bufs = [None, None]                                                                 # two output sets (state / new_state)
        action = self.action
        hits = torch.zeros((N, S), dtype=torch.int32, device=dev)
        flags = torch.zeros(N, dtype=torch.int32, device=dev)
        for i in range(steps):
            out = system.step_device(action, bufs[i & 1])
            bufs[i & 1] = out
            flags |= out['flags']
            if learning_time < steps:
                self.update_control(state, action, out['labels'], hits_out=hits)
            nxt = torch.empty_like(action)
            action, _ = self.select_action(out['obs'], action_out=nxt, adjusted_out=self.adjusted)
            state = out['obs']
            SLA_history[i] = out['labels'].sum(dim=1)
            reward_history[i] = out['reward']
            violation_history[i] = out['violations'].sum(dim=1)
            resources_history[i] = action.sum(dim=1)
            adjusted_actions[i] = self.adjusted
            hits_history[i] = hits
        torch.cuda.synchronize(dev)
        _report_flags(self, flags.cpu().numpy().view(np.uint32), self.learners)
        t = lambda x: x.transpose(0, 1).cpu().numpy()
        return {'reward': t(reward_history), 'resources': t(resources_history), 'hits': hits_history.permute(1, 2, 0).cpu().numpy(),
                'adjusted': t(adjusted_actions), 'SLA': t(SLA_history), 'violation': t(violation_history)}


def create_kbrl_agent(rng, n, accuracy_range=(0.99, 0.999), n_envs=1, dict_cap=1024, device=0, resident=False,
                      tie_seed=None, first_env_id=0, pool_mb=0, algorithm="projectron"):
    """``scenario_creator.create_kbrl_agent`` (scenario_creator.py:197-238): random initial action and security
    factor per learner (drawn per env from ``rng`` in the reference's order), gamma = 1, eta = 0.1."""
    sc = scenarios[n]
    n_embb, n_mmtc = sc['n_embb'], sc['n_mmtc']
    ia = np.zeros((n_envs, n_embb + n_mmtc), np.int64)
    sec = np.zeros_like(ia)
    for e in range(n_envs):
        for s in range(n_embb):
            ia[e, s] = rng.integers(embb_a[0], embb_a[1]); sec[e, s] = rng.integers(embb_sec[0], embb_sec[1])
        for s in range(n_embb, n_embb + n_mmtc):
            ia[e, s] = rng.integers(mmtc_a[0], mmtc_a[1]); sec[e, s] = rng.integers(mmtc_sec[0], mmtc_sec[1])
    if tie_seed is None:           # the reference's tie-break uses the global np.random; here: one more draw from rng
        tie_seed = int(rng.integers(0, 2 ** 63 - 1))
    learners = BatchedProjectron(n, n_envs, dict_cap=dict_cap, device=device, tie_seed=tie_seed, first_env_id=first_env_id,
                                 pool_mb=pool_mb, algorithm=algorithm)
    if resident:        # controller state on the GPU, torch tensors in / out (no host round trip per step)
        return DeviceKBRLControl(learners, sc['n_prbs'], ia, sec, alfa=alfa, accuracy_range=accuracy_range)
    return KBRLControl(learners, sc['n_prbs'], ia, sec, alfa=alfa, accuracy_range=accuracy_range)
