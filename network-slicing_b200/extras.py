"""Host-side mirrors of the two leaf classes the reference ships but never instantiates:

* :class:`OnOffSource`  -- ``traffic_generators.py:32-54``: a periodic source gated by geometric ON / OFF sojourns;
* :class:`SNRGenerator` -- ``channel_models.py:197-253``: per-user walk over the flat-fading ``srslte_v19.03.csv`` trace.

No factory of the reference builds them (``scenario_creator.py`` wires ``CbrSource`` / ``VbrSource`` and
``SINRSelectiveFading``), and ``SNRGenerator.get_snr`` returns a one-element list that ``SliceL1eMBB.slot`` could not
window, so they have no place on the step path and no kernel.  They are kept here with the reference's semantics, state
names and draw order so that user code written against them keeps working; the random source is any object with the
``numpy.random.Generator`` methods they call (a :class:`ranslice_b200.philox.PhiloxStream` works, which is how the tests
compare them draw for draw with the reference classes).
"""
import numpy as np


class PeriodicSource:                       # traffic_generators.py:18-30
    def __init__(self, packet_size=1000, period=2):
        self.packet_size, self.period, self.steps_to_go = packet_size, period, period

    def step(self):
        self.steps_to_go -= 1
        if self.steps_to_go == 0:
            self.steps_to_go = self.period
            return self.packet_size
        return 0


class OnOffSource:
    """``geometric`` stands for ``np.random.geometric`` (the reference draws from the legacy global RNG); pass any callable
    ``geometric(p) -> int >= 1``."""

    def __init__(self, packet_size=1000, period=2, T_on=500, T_off=1000, initial_state=1, geometric=None):
        self.geometric = geometric or (lambda p: int(np.random.geometric(p=p)))
        self.T_on, self.T_off, self.state = T_on, T_off, initial_state
        self.periodic_source = PeriodicSource(packet_size, period)
        self.time_to_change = self.geometric(1 / T_off)                      # traffic_generators.py:38 (T_off whatever the state)

    def step(self):
        if self.time_to_change == 0:                                          # :41-47
            if self.state == 1:
                self.state = 0
                self.time_to_change = self.geometric(1 / self.T_on)
            else:
                self.state = 1
                self.time_to_change = self.geometric(1 / self.T_off)
        self.time_to_change = max(self.time_to_change - 1, 0)                 # :49
        if self.state == 1:
            return self.periodic_source.step()
        return 0


class SNRGenerator:
    """``norm_snr_array`` = ``mean_snr - txpower`` of the srsLTE trace (channel_models.py:206-208); pass the array, or a
    CSV path to parse like the reference does."""

    def __init__(self, rng, norm_snr_array=None, filename='./datasets/srslte_v19.03.csv', user_ids=None, powers=None):
        self.rng = rng
        if norm_snr_array is None:
            import pandas as pd
            df = pd.read_csv(filename)
            norm_snr_array = df[["mean_snr"]].to_numpy().flatten() - df[["txpower"]].to_numpy().flatten()
        self.norm_snr_array = np.asarray(norm_snr_array, float)
        self.n_samples = len(self.norm_snr_array)
        self.users = {}
        if user_ids:
            self.insert_user_list(user_ids, powers)

    def reset(self):
        self.users = {}

    def get_snr(self, user_id, power=None):
        u = self.users[user_id]
        if power:
            u['power'] = power
        u['index'] += u['step']
        if u['index'] >= self.n_samples or u['index'] < 0:                    # channel_models.py:229-231
            u['index'] = self.rng.integers(self.n_samples)
            u['step'] = self.rng.choice([-1, 1])
        return [self.norm_snr_array[u['index']] + u['power']]

    def insert_user_list(self, user_id_list, powers=None):
        if not powers:
            powers = np.array(len(user_id_list) * [0.0], dtype=float)
        for u_id, power in zip(user_id_list, powers):
            self.insert_user(u_id, power)

    def insert_user(self, user_id, power=None):
        if not power:
            power = 0.0
        index = self.rng.integers(self.n_samples)
        step = self.rng.choice([-1, 1])
        self.users[user_id] = {'index': index, 'step': step, 'power': power}

    def extract_user(self, user_id):
        self.users.pop(user_id)
