"""Harness that runs the UNMODIFIED reference (``/root/reference``) in this container.

Test infrastructure only.  It is used (a) by ``tools/make_golden.py`` to generate the fixtures
committed under ``tests/golden/`` and (b) by the ``needs_reference`` tests, which are skipped
when ``/root/reference`` is absent (the GPU box).  Nothing here is imported by the product.

Import-time shims: ``np.int``/``np.float`` aliases, stub ``gym`` and
``matplotlib``, ``sys.path`` + cwd = reference root (dataset paths are relative,
``channel_models.py:29-33,260``), memoised ``pandas.read_csv``.

Two seeding modes:

* native  -- ``default_rng(seed)`` + ``np.random.seed(seed)`` (golden trace A);
* philox  -- every RNG holder of every slice gets a :class:`PhiloxStream`
  (``slice_l1.py:133``, ``slice_ran.py:78,159``, ``channel_models.py:116,137``) and the legacy
  global ``np.random.exponential`` used by ``VbrSource`` (``traffic_generators.py:66,96-97``)
  is routed to the VBR stream of the slice whose ``slot()`` is executing (lock-step trace B).
"""
import copy
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("RANSLICE_REFERENCE", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from ranslice_b200 import philox as px  # noqa: E402


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "node_b.py"))


_loaded = None


def _install_stubs():
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        mp = types.ModuleType("matplotlib.pyplot")
        m.pyplot = mp
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = mp
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:
            pass

        class Wrapper(Env):
            def __init__(self, env):
                self.env = env

            def __getattr__(self, name):
                if name.startswith("_"):
                    raise AttributeError(name)
                return getattr(self.env, name)

        class Box:
            def __init__(self, low, high, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        class Discrete:
            def __init__(self, n):
                self.n = n

        spaces = types.ModuleType("gym.spaces")
        spaces.Box, spaces.Discrete = Box, Discrete
        envs = types.ModuleType("gym.envs")
        reg = types.ModuleType("gym.envs.registration")
        registry = {}

        def register(id, entry_point):
            registry[id] = entry_point

        def make(id, **kw):
            from gym_ran_slice.ran_slice import RanSlice
            return RanSlice(**kw)

        reg.register = register
        envs.registration = reg
        gym.Env, gym.Wrapper, gym.spaces, gym.envs, gym.make = Env, Wrapper, spaces, envs, make
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                            "gym.envs.registration": reg})


def load_reference():
    """Import the reference modules (once); returns a namespace of modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    _install_stubs()
    for p in (REF_ROOT, os.path.join(REF_ROOT, "gym-ran_slice")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.chdir(REF_ROOT)
    import pandas as pd
    _orig_read = pd.read_csv
    _cache = {}

    def cached_read_csv(fn, *a, **kw):
        key = (os.path.abspath(fn), repr(a), repr(sorted(kw.items())))
        if key not in _cache:
            _cache[key] = _orig_read(fn, *a, **kw)
        return _cache[key].copy()

    import channel_models
    channel_models.pd = types.SimpleNamespace(read_csv=cached_read_csv)
    import traffic_generators
    import scenario_creator
    import node_b, slice_l1, slice_ran, schedulers, kbrl_control, wrapper
    import algorithms.kernel as ref_kernel
    import algorithms.projectron as ref_projectron
    import gym_ran_slice  # noqa: F401  (registers the id)
    _loaded = types.SimpleNamespace(
        channel_models=channel_models, traffic_generators=traffic_generators,
        scenario_creator=scenario_creator, node_b=node_b, slice_l1=slice_l1, slice_ran=slice_ran,
        schedulers=schedulers, kbrl_control=kbrl_control, wrapper=wrapper,
        kernel=ref_kernel, projectron=ref_projectron)
    return _loaded


# ----------------------------------------------------------------------------- native seeding
def make_env_native(seed, scenario, **kw):
    """Reference env seeded the reference's own way (both RNGs)."""
    ref = load_reference()
    ref.traffic_generators.np = np  # undo a possible philox routing
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env = ref.scenario_creator.create_env(rng, scenario, **kw)
    return env, rng


# ----------------------------------------------------------------------------- philox injection
class _Ctx:
    current_vbr = None


def _route_global_exponential(scale=1.0):
    return _Ctx.current_vbr.exponential(scale)


def make_env_philox(seed, scenario, env_id=0, **kw):
    """Reference env whose every draw comes from per-slice Philox streams (RNG contract of the
    native env, ``ranslice_b200/philox.py``).  Returns (env, streams) with all counters at 0
    right before the caller's ``env.reset()``."""
    ref = load_reference()
    tg = ref.traffic_generators
    tg.np = types.SimpleNamespace(
        rint=np.rint, random=types.SimpleNamespace(exponential=_route_global_exponential))
    dummy = np.random.default_rng(12345)  # consumed only by the constructors' resets
    env = ref.scenario_creator.create_env(dummy, scenario, **kw)
    streams = []
    shared_gen = None
    for i, l1 in enumerate(env.node_b.slices_l1):
        st = {}
        if l1.type == "eMBB":
            st["ran"] = px.PhiloxStream(seed, i, px.STREAM_RAN, env=env_id)
            st["chan"] = px.PhiloxStream(seed, i, px.STREAM_CHAN, env=env_id)
            st["l1rx"] = px.PhiloxStream(seed, i, px.STREAM_L1RX, env=env_id)
            st["vbr"] = px.PhiloxStream(seed, i, px.STREAM_VBR, env=env_id)
            l1.rng = st["l1rx"]
            if len(l1.slices_ran) == 1:
                l1.slices_ran[0].rng = st["ran"]
            else:   # L1_level=False: several RAN slices multiplexed in one L1; RAN stream of RAN slice r = (slice r, RAN)
                st["ran_mux"] = [px.PhiloxStream(seed, r, px.STREAM_RAN, env=env_id) for r in range(len(l1.slices_ran))]
                for r, sr in enumerate(l1.slices_ran):
                    sr.rng = st["ran_mux"][r]
            shared_gen = l1.snr_generator if shared_gen is None else shared_gen
            gen = copy.copy(shared_gen)          # shares .samples, private users/rng
            gen.users = {}
            gen.rng = st["chan"]
            gen.nominal_sinr = copy.copy(shared_gen.nominal_sinr)
            gen.nominal_sinr.rng = st["chan"]
            l1.snr_generator = gen
        else:
            # one stream per mMTC RAN slice: slice id = L1 index + position inside the L1 (a multiplexed L1 holds all of them)
            st["mtc_all"] = [px.PhiloxStream(seed, i + m, px.STREAM_MTC, env=env_id) for m in range(len(l1.slices_ran))]
            st["mtc"] = st["mtc_all"][0]
            for m, sr in enumerate(l1.slices_ran):
                sr.rng = st["mtc_all"][m]
        streams.append(st)
        orig_slot = l1.slot

        def slot(orig_slot=orig_slot, st=st):
            _Ctx.current_vbr = st.get("vbr")
            return orig_slot()

        l1.slot = slot
    return env, streams


def simplex_actions(seed, n_slices, n_prbs, steps):
    """Uniform random agent mapped like ``wrapper.py:77-82`` (SURVEY §8d action policy)."""
    rng = np.random.default_rng(1000 + seed)
    w = rng.random((steps, n_slices + 1))
    a = np.floor(n_prbs * w[:, :n_slices] / w.sum(axis=1, keepdims=True))
    return a.astype(np.int64)


EMBB_VARS = ['cbr_traffic', 'cbr_th', 'cbr_prb', 'cbr_queue', 'cbr_snr',
             'vbr_traffic', 'vbr_th', 'vbr_prb', 'vbr_queue', 'vbr_snr']
MMTC_VARS = ['devices', 'avg_rep', 'delay']


def run_trace(env, actions):
    """Steps ``env`` through ``actions`` [T,S]; returns dict of stacked outputs incl. raw
    accumulators (info['l1_info'], ``node_b.py:46-49``) padded to 10 per slice."""
    T, S = actions.shape
    obs0 = env.reset()
    V = len(obs0)
    out = dict(obs=np.zeros((T, V), np.float32), reward=np.zeros(T, np.float64),
               labels=np.zeros((T, S), np.int32), violations=np.zeros((T, S), np.int32),
               acc=np.zeros((T, S, 10), np.float64), obs0=np.asarray(obs0, np.float32))
    for t in range(T):
        o, r, done, info = env.step(actions[t])
        out["obs"][t] = o
        out["reward"][t] = r
        out["labels"][t] = info["SLA_labels"]
        out["violations"][t] = info["violations"]
        for s, l1 in enumerate(info["l1_info"]):
            d = l1[0]
            names = EMBB_VARS if "cbr_th" in d else MMTC_VARS
            for j, nme in enumerate(names):
                out["acc"][t, s, j] = d[nme]
        if any(len(l1) > 1 for l1 in info["l1_info"]):        # multiplexed L1: accumulators of every RAN slice, L1-major
            rows = []
            for l1 in info["l1_info"]:
                for r in sorted(l1):
                    d = l1[r]
                    names = EMBB_VARS if "cbr_th" in d else MMTC_VARS
                    rows.append([d[n] for n in names] + [0.0] * (10 - len(names)))
            out.setdefault("acc_ran", np.zeros((T, len(rows), 10), np.float64))[t] = np.asarray(rows, np.float64)
    return out
