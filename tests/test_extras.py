"""The leaf classes the reference never instantiates (OnOffSource, SNRGenerator), mirrored in ranslice_b200/extras.py, against
the reference classes themselves with the same injected random sources, draw for draw."""
import numpy as np
import pytest

import refharness as rh
from ranslice_b200 import philox as px
from ranslice_b200.extras import OnOffSource, SNRGenerator

pytestmark = pytest.mark.needs_reference


def test_onoff_source_matches_reference():
    ref = rh.load_reference()
    tg = ref.traffic_generators
    saved = tg.np
    try:
        for seed in (1, 2, 3):
            g1, g2 = np.random.RandomState(seed), np.random.RandomState(seed)
            import types
            tg.np = types.SimpleNamespace(random=types.SimpleNamespace(geometric=lambda p: g1.geometric(p=p)))
            a = tg.OnOffSource(packet_size=700, period=3, T_on=20, T_off=35, initial_state=seed % 2)
            b = OnOffSource(packet_size=700, period=3, T_on=20, T_off=35, initial_state=seed % 2, geometric=lambda p: int(g2.geometric(p=p)))
            out_a = [a.step() for _ in range(3000)]
            out_b = [b.step() for _ in range(3000)]
            assert out_a == out_b and sum(out_a) > 0 and 0 in out_a
    finally:
        tg.np = saved


def test_snr_generator_matches_reference():
    ref = rh.load_reference()
    cm = ref.channel_models
    a = cm.SNRGenerator(px.PhiloxStream(77, 0, px.STREAM_CHAN), user_ids=[3, 4], powers=[1.5, 0.0])
    b = SNRGenerator(px.PhiloxStream(77, 0, px.STREAM_CHAN), norm_snr_array=a.norm_snr_array, user_ids=[3, 4], powers=[1.5, 0.0])
    a.insert_user(9); b.insert_user(9)
    for t in range(4000):
        for uid in (3, 4, 9):
            assert a.get_snr(uid) == b.get_snr(uid), (t, uid)
    a.extract_user(4); b.extract_user(4)
    assert sorted(a.users) == sorted(b.users) and a.users[3] == b.users[3]
