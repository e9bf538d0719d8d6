"""ctypes binding of the CPU oracle (oracle/libranslice_oracle.so).  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.environ.get("RANSLICE_ORACLE_LIB") or os.path.join(_ROOT, "oracle", "libranslice_oracle.so")   # RANSLICE_ORACLE_LIB: sanitizer builds

PROPAGATION = {"macro_cell_urban_2GHz": (128.1, 37.6), "macro_cell_urban_900MHz": (120.9, 37.6),
               "macro_cell_rural": (95.5, 34.1)}
SCENARIOS = [(200, 5, 0), (150, 3, 2), (100, 1, 4), (70, 1, 1)]   # n_prbs, n_embb, n_mmtc


class OrcConfig(C.Structure):
    _fields_ = [("n_prbs", C.c_int32), ("n_embb", C.c_int32), ("n_mmtc", C.c_int32),
                ("slots_per_step", C.c_int32), ("penalty", C.c_double), ("prop_A", C.c_double),
                ("prop_B", C.c_double), ("l1_mux", C.c_int32), ("reserved", C.c_int32)]


class OrcTables(C.Structure):
    _fields_ = [("trace", C.c_void_p), ("mcs_rate", C.c_void_p), ("mcs_snr", C.c_void_p),
                ("mcs_order", C.c_void_p), ("mcs_mod", C.c_void_p)]


F_RANDOM = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int, C.c_int)
F_EXP = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_double)
F_INT = C.CFUNCTYPE(C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int64)
F_NORMAL = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double)
F_RANDOM2 = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double))


class OrcRng(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("random", F_RANDOM), ("exponential", F_EXP), ("integers", F_INT),
                ("choice", F_INT), ("normal", F_NORMAL), ("random2", F_RANDOM2)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcTables), C.c_uint64, C.c_uint32]
        L.orc_set_rng.argtypes = [C.c_void_p, C.POINTER(OrcRng)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_n_variables.argtypes = [C.c_void_p]
        L.orc_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_step.restype = C.c_uint32
        L.orc_step.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_step_batch.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.orc_mcs_lut.argtypes = [C.POINTER(OrcTables), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                  C.POINTER(C.c_int)]
        L.orc_response.restype = C.c_double
        L.orc_response.argtypes = [C.POINTER(OrcTables), C.c_int, C.c_void_p, C.c_int]
        L.orc_macro_cell.restype = C.c_double
        L.orc_macro_cell.argtypes = [C.c_double] * 5
        L.orc_n_ues.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_acc_ran.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def c_tables(tables):
    t = OrcTables(_ptr(tables.trace), _ptr(tables.mcs_rate), _ptr(tables.mcs_snr), _ptr(tables.mcs_order),
                  _ptr(tables.mcs_mod))
    t._keep = tables
    return t


class NumpyRng:
    """orc_rng vtable backed by a numpy Generator + the legacy global np.random (VBR stream),
    i.e. exactly the reference's native RNG usage."""

    def __init__(self, rng):
        self.rng = rng
        g = rng

        def random(ctx, sl, st):
            return float(g.random())

        def exponential(ctx, sl, st, scale):
            if st == 3:                                   # ORC_STREAM_VBR -> legacy global stream
                return float(np.random.exponential(scale))
            return float(g.exponential(scale))

        def integers(ctx, sl, st, n):
            return int(g.integers(n))

        def choice(ctx, sl, st, n):
            return int(g.choice(n))

        def normal(ctx, sl, st, mu, sigma):
            return float(g.normal(mu, sigma))

        def random2(ctx, sl, st, out):
            v = g.random(2)
            out[0] = v[0]
            out[1] = v[1]

        self.struct = OrcRng(None, F_RANDOM(random), F_EXP(exponential), F_INT(integers), F_INT(choice),
                             F_NORMAL(normal), F_RANDOM2(random2))


class OracleEnv:
    """One oracle environment (scenario index like scenario_creator.create_env)."""

    def __init__(self, tables, scenario, seed, slots_per_step=50, penalty=100.0,
                 propagation="macro_cell_urban_2GHz", numpy_rng=None, l1_mux=False, env_id=0):
        n_prbs, n_embb, n_mmtc = SCENARIOS[scenario]
        A, B = PROPAGATION[propagation]
        self.cfg = OrcConfig(n_prbs, n_embb, n_mmtc, slots_per_step, penalty, A, B, int(bool(l1_mux)), 0)
        self.tbl = c_tables(tables)
        self.S = (int(n_embb > 0) + int(n_mmtc > 0)) if l1_mux else n_embb + n_mmtc   # L1 slices = action entries (create_env(L1_level=False))
        self.n_ran = n_embb + n_mmtc
        self.V = 10 * n_embb + 3 * n_mmtc
        self.n_prbs = n_prbs
        self.h = lib().orc_create(C.byref(self.cfg), C.byref(self.tbl), C.c_uint64(seed), C.c_uint32(env_id))
        self._rng = None
        if numpy_rng is not None:
            self._rng = NumpyRng(numpy_rng)
            lib().orc_set_rng(self.h, C.byref(self._rng.struct))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def reset(self):
        obs = np.zeros(self.V, np.float32)
        lib().orc_reset(self.h, _ptr(obs))
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, np.int64)
        obs = np.zeros(self.V, np.float32)
        rew = np.zeros(1, np.float64)
        lab = np.zeros(self.S, np.int32)
        vio = np.zeros(self.S, np.int32)
        acc = np.zeros((self.S, 10), np.float64)
        flags = lib().orc_step(self.h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(lab), _ptr(vio), _ptr(acc))
        return obs, float(rew[0]), lab, vio, acc, flags

    def acc_ran(self):
        """Raw accumulators of every RAN slice after the last step, L1-major [n_embb + n_mmtc, 10]."""
        a = np.zeros((self.n_ran, 10), np.float64)
        lib().orc_get_acc_ran(self.h, _ptr(a))
        return a


class OracleBatch:
    """N independent oracle envs (Philox key base_seed, env ids first_env + i), stepped by n_threads pthreads."""

    def __init__(self, tables, scenario, n_envs, base_seed, n_threads=1, first_env=0, **kw):
        self.envs = [OracleEnv(tables, scenario, base_seed, env_id=first_env + i, **kw) for i in range(n_envs)]
        self.N, self.S, self.V = n_envs, self.envs[0].S, self.envs[0].V
        self.n_threads = n_threads
        self.handles = (C.c_void_p * n_envs)(*[e.h for e in self.envs])

    def reset(self):
        return np.stack([e.reset() for e in self.envs])

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.int64).reshape(self.N, self.S)
        obs = np.zeros((self.N, self.V), np.float32)
        rew = np.zeros(self.N, np.float64)
        lab = np.zeros((self.N, self.S), np.int32)
        vio = np.zeros((self.N, self.S), np.int32)
        flags = np.zeros(self.N, np.uint32)
        lib().orc_step_batch(self.handles, self.N, self.n_threads, _ptr(a), _ptr(obs), _ptr(rew), _ptr(lab),
                             _ptr(vio), _ptr(flags))
        return obs, rew, lab, vio, flags


# ----------------------------------------------------------------------------- KBRL oracle (oracle/kbrl_oracle.c)
def _kb_lib():
    L = lib()
    if not getattr(L, "_kb_ready", False):
        L.orc_kb_create.restype = C.c_void_p
        L.orc_kb_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                    C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.orc_kb_destroy.argtypes = [C.c_void_p]
        L.orc_kb_update_control.argtypes = [C.c_void_p] * 5
        L.orc_kb_select_action.argtypes = [C.c_void_p] * 4
        L.orc_kb_get_control.argtypes = [C.c_void_p] * 6
        L.orc_kb_get_learner.restype = C.c_int
        L.orc_kb_get_learner.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_kb_predict.restype = C.c_double
        L.orc_kb_predict.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_kb_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int32]
        L.orc_kb_set_tie_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.orc_kb_set_algorithm.argtypes = [C.c_void_p, C.c_int]
        L._kb_ready = True
    return L


class OracleKBRL:
    """One KBRL_Control (kbrl_control.py) with S Projectron learners, CPU oracle."""

    def __init__(self, dims, n_prbs, init_action, init_sec, accuracy_range=(0.99, 0.999), alfa=0.05, gamma=1.0, eta=0.1,
                 tie_seed=None, env_id=0, plus=False):
        self.S = len(dims)
        self.dims = np.ascontiguousarray(dims, np.int32)
        self.offsets = np.ascontiguousarray(np.concatenate([[0], np.cumsum(self.dims - 1)[:-1]]), np.int32)
        self.n_prbs = n_prbs
        ia = np.ascontiguousarray(init_action, np.int64)
        isec = np.ascontiguousarray(init_sec, np.int64)
        self.h = _kb_lib().orc_kb_create(self.S, _ptr(self.dims), _ptr(self.offsets), n_prbs, alfa, accuracy_range[0],
                                         accuracy_range[1], _ptr(ia), _ptr(isec), gamma, eta)
        if plus:                       # ProjectronPlus (algorithms/projectron.py:66-107)
            _kb_lib().orc_kb_set_algorithm(self.h, 1)
        if tie_seed is not None:       # f == 0 tie-break of GaussianKernel.predict (kernel.py:26-27) from the KBRL Philox stream
            _kb_lib().orc_kb_set_tie_stream(self.h, C.c_uint64(tie_seed), C.c_uint32(env_id))

    def __del__(self):
        if getattr(self, "h", None):
            _kb_lib().orc_kb_destroy(self.h)
            self.h = None

    def update_control(self, state, action, labels):
        st = np.ascontiguousarray(state, np.float32)
        a = np.ascontiguousarray(action, np.int64)
        lab = np.ascontiguousarray(labels, np.int64)
        hits = np.zeros(self.S, np.int64)
        _kb_lib().orc_kb_update_control(self.h, _ptr(st), _ptr(a), _ptr(lab), _ptr(hits))
        return hits

    def select_action(self, state):
        st = np.ascontiguousarray(state, np.float32)
        a = np.zeros(self.S, np.int64)
        adj = np.zeros(1, np.int32)
        _kb_lib().orc_kb_select_action(self.h, _ptr(st), _ptr(a), _ptr(adj))
        return a, int(adj[0])

    def control(self):
        sec = np.zeros(self.S, np.int64); mar = np.zeros(self.S, np.int64); sizes = np.zeros(self.S, np.int64)
        acc = np.zeros((self.S, self.n_prbs), np.float64)
        tb = C.c_long()
        _kb_lib().orc_kb_get_control(self.h, _ptr(sec), _ptr(mar), _ptr(acc), _ptr(sizes), C.byref(tb))
        return dict(security_factors=sec, margins=mar, accuracies=acc, sizes=sizes, tie_breaks=tb.value)

    def learner(self, s):
        D = int(self.control()["sizes"][s])
        d = int(self.dims[s])
        lm = np.zeros((max(D, 1), d)); cf = np.zeros(max(D, 1)); ki = np.zeros((max(D, 1), max(D, 1)))
        _kb_lib().orc_kb_get_learner(self.h, s, _ptr(lm), _ptr(cf), _ptr(ki))
        return lm[:D], cf[:D], ki[:D, :D]
