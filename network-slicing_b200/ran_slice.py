"""Single-env facade with exactly the ``RanSlice`` surface (gym-ran_slice/gym_ran_slice/ran_slice.py:15-57)
so that the reference's ``wrapper.py`` (ReportWrapper / DQNWrapper / TimerWrapper) and
``KBRL_Control.run`` attach unchanged.  Backed by a 1-env native batch."""
import numpy as np

from .scenario_creator import state_variables_embb, state_variables_mmtc

try:                                   # gym is optional: subclass gym.Env when it is installed
    import gym as _gym
    _Base = _gym.Env
    _spaces = _gym.spaces
except Exception:                      # pragma: no cover - depends on the environment
    _Base = object
    _spaces = None


class _Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class _NodeBView:
    """Minimal ``node_b`` attribute (ran_slice.py:20-23 read n_prbs / n_slices_l1 from it)."""

    def __init__(self, env):
        self.n_prbs = env.n_prbs
        self.n_slices_l1 = env.n_slices
        self.slots_per_step = env.slots_per_step


class RanSlice(_Base):
    def __init__(self, batched, penalty=None):
        assert batched.n_envs == 1
        self.batched = batched
        self.node_b = _NodeBView(batched)
        self.penalty = batched.penalty if penalty is None else penalty
        self.n_prbs = batched.n_prbs
        self.n_slices = batched.n_slices
        self.n_variables = batched.n_variables
        Box = _spaces.Box if _spaces is not None else _Box
        self.action_space = Box(low=0, high=self.n_prbs, shape=(self.n_slices,), dtype=np.int64)
        self.observation_space = Box(low=-float('inf'), high=+float('inf'), shape=(self.n_variables,),
                                     dtype=np.float64)

    def reset(self):
        return self.batched.reset()[0]

    def step(self, action):
        action = np.asarray(action)
        if len(action) != self.n_slices:
            raise ValueError('The action must contain as many elements as slices!')     # node_b.py:66-68
        obs, reward, _, binfo = self.batched.step(action.reshape(1, -1))
        acc, prbs = self.batched.get_info(0)
        b = self.batched
        l1_info = []
        if getattr(b, 'l1_level', True):
            for s in range(self.n_slices):                                              # node_b.py:46-49
                names = state_variables_embb if s < b.n_embb else state_variables_mmtc
                l1_info.append({0: {n: acc[s, j] for j, n in enumerate(names)}})
        else:                                                                           # slice_l1.py:178-181: {i: slice_ran.info}
            if b.n_embb:
                l1_info.append({r: {n: acc[r, j] for j, n in enumerate(state_variables_embb)} for r in range(b.n_embb)})
            if b.n_mmtc:
                l1_info.append({m: {n: acc[b.n_embb + m, j] for j, n in enumerate(state_variables_mmtc)} for m in range(b.n_mmtc)})
        info = {'l1_info': l1_info, 'SLA_labels': binfo['SLA_labels'][0].astype(np.int64),
                'violations': binfo['violations'][0].astype(np.int64), 'n_prbs': [int(x) for x in prbs],
                'total_violations': int(binfo['total_violations'][0]), 'flags': int(binfo['flags'][0])}
        return obs[0], float(reward[0]), False, info

    def render(self):
        pass
