"""Host-side equivalents of the two leaf classes the reference ships but never instantiates:

* :class:`OnOffSource`  -- behaviour of ``traffic_generators.py:32-54``: a periodic source gated by geometric ON / OFF sojourns;
* :class:`SNRGenerator` -- behaviour of ``channel_models.py:197-253``: per-user walk over the flat-fading ``srslte_v19.03.csv`` trace.

No factory of the reference builds them (``scenario_creator.py`` wires ``CbrSource`` / ``VbrSource`` and
``SINRSelectiveFading``), and ``SNRGenerator.get_snr`` returns a one-element list that ``SliceL1eMBB.slot`` could not
window, so they have no place on the step path and no kernel.  They are provided with the reference's constructor
arguments, method names, observable state and DRAW ORDER, so that user code written against them keeps working; the
random source is any object with the ``numpy.random.Generator`` methods they call (a
:class:`ranslice_b200.philox.PhiloxStream` works, which is how ``tests/test_extras.py`` compares them draw for draw with
the reference classes).
"""
from dataclasses import dataclass

import numpy as np


class PeriodicSource:
    """``packet_size`` bits every ``period``-th call (traffic_generators.py:18-30)."""

    def __init__(self, packet_size=1000, period=2):
        self.packet_size, self.period, self.steps_to_go = packet_size, period, period

    def step(self):
        self.steps_to_go -= 1
        if self.steps_to_go:
            return 0
        self.steps_to_go = self.period
        return self.packet_size


class OnOffSource:
    """Two-state gate in front of a :class:`PeriodicSource`.  ``geometric(p) -> int >= 1`` stands for ``np.random.geometric``
    (the reference draws from the legacy global RNG).  Quirks kept: the first sojourn is drawn with ``1 / T_off`` whatever
    the initial state (traffic_generators.py:38), an ON sojourn ENDS with a draw of mean ``T_on`` for the OFF period that
    follows and vice versa (:41-47), and the periodic source only advances while ON."""

    def __init__(self, packet_size=1000, period=2, T_on=500, T_off=1000, initial_state=1, geometric=None):
        self.geometric = geometric or (lambda p: int(np.random.geometric(p=p)))
        self.T_on, self.T_off, self.state = T_on, T_off, initial_state
        self.periodic_source = PeriodicSource(packet_size, period)
        self.time_to_change = self.geometric(1 / T_off)

    def step(self):
        if self.time_to_change == 0:
            leaving_on = self.state == 1
            self.state = 0 if leaving_on else 1
            self.time_to_change = self.geometric(1 / (self.T_on if leaving_on else self.T_off))
        if self.time_to_change > 0:
            self.time_to_change -= 1
        return self.periodic_source.step() if self.state == 1 else 0


@dataclass
class _Walker:
    index: int
    step: int
    power: float


class SNRGenerator:
    """Flat-fading SNR per user: every user walks ``norm_snr_array`` (= ``mean_snr - txpower`` of the srsLTE trace,
    channel_models.py:206-208) one sample per call, in its own direction, and jumps to a random position and direction when
    it leaves the trace (:229-231: position first, then direction).  Pass the array, or a CSV path to parse like the
    reference does.  ``users`` exposes the walkers as the reference's ``{id: {'index', 'step', 'power'}}`` dictionary."""

    def __init__(self, rng, norm_snr_array=None, filename='./datasets/srslte_v19.03.csv', user_ids=None, powers=None):
        self.rng = rng
        if norm_snr_array is None:
            import pandas as pd
            df = pd.read_csv(filename)
            norm_snr_array = df["mean_snr"].to_numpy() - df["txpower"].to_numpy()
        self.norm_snr_array = np.asarray(norm_snr_array, float)
        self.n_samples = len(self.norm_snr_array)
        self._walkers = {}
        if user_ids:
            self.insert_user_list(user_ids, powers)

    @property
    def users(self):
        return {uid: {'index': w.index, 'step': w.step, 'power': w.power} for uid, w in self._walkers.items()}

    def reset(self):
        self._walkers = {}

    def _draw(self):
        index = self.rng.integers(self.n_samples)        # one draw for the position, one for the direction, in this order
        return index, self.rng.choice([-1, 1])

    def insert_user(self, user_id, power=None):
        index, step = self._draw()
        self._walkers[user_id] = _Walker(index, step, power if power else 0.0)

    def insert_user_list(self, user_id_list, powers=None):
        if not powers:
            powers = [0.0] * len(user_id_list)
        for uid, power in zip(user_id_list, powers):
            self.insert_user(uid, power)

    def extract_user(self, user_id):
        del self._walkers[user_id]

    def get_snr(self, user_id, power=None):
        w = self._walkers[user_id]
        if power:
            w.power = power
        w.index += w.step
        if not 0 <= w.index < self.n_samples:
            w.index, w.step = self._draw()
        return [self.norm_snr_array[w.index] + w.power]
