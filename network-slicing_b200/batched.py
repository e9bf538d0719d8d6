"""BatchedRanSlice: N independent RAN-slicing envs advanced in lockstep on one B200.

Mirrors ``RanSlice.reset()/step()`` (gym-ran_slice/gym_ran_slice/ran_slice.py:30-54) with a
leading env axis.  Two call paths, both through the C ABI (include/ranslice_b200.h):

* ``step(action)``          host numpy in / host numpy out (``rs_step``: H2D + kernels + D2H);
* ``step_device(action)``   torch CUDA tensors in / out, asynchronous on torch's current stream
                            (``rs_step_device``); no host round trip.
"""
import ctypes as C

import numpy as np

from . import _lib
from .scenario_creator import PROPAGATION, scenarios
from .tables import load_tables


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class BatchedRanSlice:
    def __init__(self, scenario=0, n_envs=1, base_seed=0, slots_per_step=50,
                 propagation_type='macro_cell_urban_2GHz', penalty=100, device=0, first_env_id=0,
                 max_ues=0, max_bursts=0, mtc_queue_cap=0, kernel_variant=0, tables=None, l1_level=True):
        sc = scenarios[scenario] if isinstance(scenario, int) else scenario
        self.n_prbs, self.n_embb, self.n_mmtc = sc['n_prbs'], sc['n_embb'], sc['n_mmtc']
        # create_env(L1_level=False) (scenario_creator.py:168-177): the eMBB RAN slices are multiplexed in ONE L1 slice, which
        # takes one action entry and reports one label; the observation keeps the 10 variables of every RAN slice
        self.l1_level = bool(l1_level)
        self.n_l1_embb = self.n_embb if self.l1_level else int(self.n_embb > 0)
        self.n_l1_mmtc = self.n_mmtc if self.l1_level else int(self.n_mmtc > 0)     # ... and so are the mMTC RAN slices (:173-176)
        self.n_slices = self.n_l1_embb + self.n_l1_mmtc
        self.n_ran = self.n_embb + self.n_mmtc
        self.n_variables = 10 * self.n_embb + 3 * self.n_mmtc
        self.n_envs, self.device, self.penalty = n_envs, device, penalty
        self.slots_per_step = slots_per_step
        A, B = PROPAGATION[propagation_type]
        L = _lib.lib()
        t = tables or load_tables()
        self._tables = t
        self._cfg = _lib.RsConfig(_lib.RS_ABI_VERSION, device, n_envs, self.n_prbs, self.n_embb, self.n_mmtc,
                                  slots_per_step, max_ues, max_bursts, mtc_queue_cap, kernel_variant, int(not self.l1_level),
                                  float(penalty), A, B, base_seed & (2 ** 64 - 1), first_env_id)
        tb = _lib.RsTables(_p(t.trace), _p(t.mcs_rate), _p(t.mcs_snr), _p(t.mcs_order), _p(t.mcs_mod))
        h = C.c_void_p()
        _lib.check(L.rs_create(C.byref(self._cfg), C.byref(tb), C.byref(h)))
        self._h = h
        N, S, V = n_envs, self.n_slices, self.n_variables
        self._pin = None
        self._alloc_host(N, S, V)

    # ---------------------------------------------------------------- host buffers (pinned when torch+CUDA)
    def _alloc_host(self, N, S, V):
        try:
            import torch
            pin = torch.cuda.is_available()
            mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pin)
            self._pin = dict(action=mk((N, S), torch.int32), obs=mk((N, V), torch.float32),
                             reward=mk((N,), torch.float32), labels=mk((N, S), torch.int32),
                             violations=mk((N, S), torch.int32), flags=mk((N,), torch.int32))
            self._hb = {k: v.numpy() for k, v in self._pin.items()}
            self._hb['flags'] = self._hb['flags'].view(np.uint32)
        except ImportError:
            self._hb = dict(action=np.empty((N, S), np.int32), obs=np.empty((N, V), np.float32),
                            reward=np.empty(N, np.float32), labels=np.empty((N, S), np.int32),
                            violations=np.empty((N, S), np.int32), flags=np.empty(N, np.uint32))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().rs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- gym-like API
    def reset(self):
        """-> float32 zeros [N, V] (ran_slice.py:30-36)."""
        _lib.check(_lib.lib().rs_reset(self._h, _p(self._hb['obs'])))
        return self._hb['obs'].copy()

    def step(self, action):
        """action int [N, S] -> (obs f32 [N,V], reward f32 [N], done False, info dict of arrays)."""
        a = np.asarray(action)
        if a.shape != (self.n_envs, self.n_slices):
            raise ValueError("The action must contain as many elements as slices! expected %s got %s"
                             % ((self.n_envs, self.n_slices), a.shape))     # node_b.py:66-68 prints; we raise
        hb = self._hb
        np.copyto(hb['action'], a, casting='unsafe')
        _lib.check(_lib.lib().rs_step(self._h, _p(hb['action']), _p(hb['obs']), _p(hb['reward']), _p(hb['labels']),
                                      _p(hb['violations']), _p(hb['flags'])))
        info = {'SLA_labels': hb['labels'].copy(), 'violations': hb['violations'].copy(),
                'total_violations': hb['violations'].sum(axis=1), 'flags': hb['flags'].copy()}
        return hb['obs'].copy(), hb['reward'].copy(), False, info

    def step_host_inplace(self, action_i32):
        """Zero-copy variant for throughput loops: ``action_i32`` is int32 [N,S] (ideally pinned);
        results land in the env's pinned host buffers (returned without copying)."""
        hb = self._hb
        _lib.check(_lib.lib().rs_step(self._h, _p(action_i32), _p(hb['obs']), _p(hb['reward']), _p(hb['labels']),
                                      _p(hb['violations']), _p(hb['flags'])))
        return hb

    def alloc_host_buffers(self):
        """A fresh set of (pinned) host output buffers for :meth:`step_host_async`: dict of numpy arrays."""
        N, S, V = self.n_envs, self.n_slices, self.n_variables
        try:
            import torch
            pin = torch.cuda.is_available()
            t = dict(obs=torch.empty((N, V), dtype=torch.float32, pin_memory=pin), reward=torch.empty((N,), dtype=torch.float32, pin_memory=pin),
                     labels=torch.empty((N, S), dtype=torch.int32, pin_memory=pin), violations=torch.empty((N, S), dtype=torch.int32, pin_memory=pin),
                     flags=torch.empty((N,), dtype=torch.int32, pin_memory=pin))
            hb = {k: v.numpy() for k, v in t.items()}
            hb['flags'] = hb['flags'].view(np.uint32)
            hb['_pin'] = t
            return hb
        except ImportError:
            return dict(obs=np.empty((N, V), np.float32), reward=np.empty(N, np.float32), labels=np.empty((N, S), np.int32),
                        violations=np.empty((N, S), np.int32), flags=np.empty(N, np.uint32))

    def step_host_async(self, action_i32, hb):
        """Pipelined ``step``: enqueue H2D + kernels + D2H and return a ticket at once; ``wait(ticket)`` blocks until
        the results are in ``hb`` (from :meth:`alloc_host_buffers`).  Two steps may be in flight, so alternate
        between two buffer sets; ``action_i32`` (int32 [N,S], ideally pinned) and ``hb`` must stay untouched until
        the wait.  Same results as ``step``."""
        t = C.c_int32()
        _lib.check(_lib.lib().rs_step_async(self._h, _p(action_i32), _p(hb['obs']), _p(hb['reward']), _p(hb['labels']),
                                            _p(hb['violations']), _p(hb['flags']), C.byref(t)))
        return int(t.value)

    def wait(self, ticket):
        _lib.check(_lib.lib().rs_wait(self._h, int(ticket)))

    def step_device(self, action, out=None):
        """torch int32 CUDA tensor [N,S] -> dict of CUDA tensors; async on the current torch stream."""
        import torch
        assert action.is_cuda and action.dtype == torch.int32 and action.is_contiguous()
        assert tuple(action.shape) == (self.n_envs, self.n_slices)
        if out is None:
            dev = action.device
            N, S, V = self.n_envs, self.n_slices, self.n_variables
            out = dict(obs=torch.empty((N, V), dtype=torch.float32, device=dev),
                       reward=torch.empty((N,), dtype=torch.float32, device=dev),
                       labels=torch.empty((N, S), dtype=torch.int32, device=dev),
                       violations=torch.empty((N, S), dtype=torch.int32, device=dev),
                       flags=torch.empty((N,), dtype=torch.int32, device=dev))
        stream = torch.cuda.current_stream(action.device).cuda_stream
        _lib.check(_lib.lib().rs_step_device(
            self._h, C.c_void_p(action.data_ptr()), C.c_void_p(out['obs'].data_ptr()),
            C.c_void_p(out['reward'].data_ptr()), C.c_void_p(out['labels'].data_ptr()),
            C.c_void_p(out['violations'].data_ptr()), C.c_void_p(out['flags'].data_ptr()), C.c_void_p(stream)))
        return out

    # ---------------------------------------------------------------- introspection
    def get_info(self, env=0):
        acc = np.zeros((self.n_ran, 10), np.float64)          # one row per RAN slice (L1-major), == per L1 unless multiplexed
        prbs = np.zeros(self.n_slices, np.int32)
        _lib.check(_lib.lib().rs_get_info(self._h, env, _p(acc), _p(prbs)))
        return acc, prbs

    def n_ues(self):
        out = np.zeros((self.n_envs, max(self.n_l1_embb, 1)), np.int32)
        _lib.check(_lib.lib().rs_get_n_ues(self._h, _p(out)))
        return out[:, :self.n_l1_embb]

    def counters(self):
        k, t = C.c_uint64(), C.c_uint64()
        _lib.check(_lib.lib().rs_get_counters(self._h, C.byref(k), C.byref(t)))
        return int(k.value), int(t.value)

    def state_bytes(self):
        n = C.c_size_t()
        _lib.check(_lib.lib().rs_state_size(self._h, C.byref(n)))
        return int(n.value)

    VARIANT_NAMES = {1: "embb_step_unit_thread (all fp64)", 2: "embb_step_fast (PRB-sorted units, guarded fast math)",
                     3: "embb_step_warp (one warp per unit: lanes over UEs / PRB quads, guarded fast math)",
                     4: "embb_step_smem (PRB-sorted units, UE table in shared memory, guarded fast math)",
                     5: "embb_step_mux_thread (multiplexed L1, all fp64)"}

    def active_variant(self):
        """The kernel variant that steps the eMBB slices (``kernel_variant`` 0 resolved by the batch size)."""
        return int(_lib.lib().rs_active_variant(self._h))

    def kernel_variant_name(self):
        v = self.active_variant()
        return self.VARIANT_NAMES.get(v, str(v))

    def profile_steps(self, actions, out=None):
        """Runs step_device over ``actions`` with per-kernel CUDA events (inside the library, on the
        launching stream); returns average per-launch durations in ms."""
        L = _lib.lib()
        _lib.check(L.rs_set_profiling(self._h, 1))
        for a in actions:
            out = self.step_device(a, out)
        ms, n = (C.c_double * 6)(), C.c_uint64()
        _lib.check(L.rs_get_profile(self._h, ms, C.byref(n)))
        _lib.check(L.rs_set_profiling(self._h, 0))
        k = max(int(n.value), 1)
        names = ("sort_ms", "dominant_ms", "embb_rest_ms", "mmtc_scan_ms", "mmtc_step_ms", "reward_ms")
        out = {nm: ms[i] / k for i, nm in enumerate(names)}
        out.update({"embb_ms": (ms[0] + ms[1] + ms[2]) / k, "mmtc_ms": (ms[3] + ms[4]) / k, "steps": int(n.value),
                    "kernel": self.kernel_variant_name()})
        return out

    def set_route_limits(self, single_start_max=6, single_slots=8, pair_start_max=14, pair_slots=16):
        """Routing limits of the default eMBB kernel (tests; results never depend on them), see rs_set_route_limits."""
        _lib.check(_lib.lib().rs_set_route_limits(self._h, single_start_max, single_slots, pair_start_max, pair_slots))

    def set_heavy_threshold(self, contended_chunks_per_step, max_units=0):
        """Workload knob of the lane-per-unit route (see rs_set_heavy_threshold); results never depend on it."""
        _lib.check(_lib.lib().rs_set_heavy_threshold(self._h, int(contended_chunks_per_step), int(max_units)))

    def routes(self):
        """Units of the last step by route: dict(single, pair, general, aborted, warp)."""
        out = (C.c_uint64 * 5)()
        _lib.check(_lib.lib().rs_get_routes(self._h, out))
        return {"single": int(out[0]), "pair": int(out[1]), "general": int(out[2]), "aborted": int(out[3]), "warp": int(out[4])}

    def set_debug_check(self, on=True):
        _lib.check(_lib.lib().rs_set_debug_check(self._h, int(bool(on))))

    def diag(self):
        """Guard-band diagnostics of the default kernel (see rs_get_diag in include/ranslice_b200.h)."""
        out = (C.c_double * 6)()
        _lib.check(_lib.lib().rs_get_diag(self._h, out, 6))
        return {"max_p_err_over_eps": out[0], "max_mean_err_over_guard": out[1], "decision_mismatches": int(out[2]),
                "slow_snr_last_step": int(out[3]), "slow_rx_last_step": int(out[4]), "pf_batched_chunks_last_step": int(out[5])}

    def get_state(self):
        n = C.c_size_t()
        _lib.check(_lib.lib().rs_state_size(self._h, C.byref(n)))
        blob = np.empty(n.value, np.uint8)
        _lib.check(_lib.lib().rs_get_state(self._h, _p(blob), n))
        return blob

    def set_state(self, blob):
        blob = np.ascontiguousarray(blob, np.uint8)
        _lib.check(_lib.lib().rs_set_state(self._h, _p(blob), C.c_size_t(blob.size)))
