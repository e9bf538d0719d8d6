"""The ``gym_ran_slice`` shim keeps the reference's id and constructor kwargs (gym-ran_slice/gym_ran_slice/__init__.py:5-8,
scenario_creator.py:181).  CPU part: import, spec extraction from a NodeB-like object; GPU part: make() -> working env."""
import importlib.util
import os
import types

import numpy as np
import pytest


def _shim():
    """This repo's ``gym_ran_slice`` package, loaded by path: the test harness of the reference-backed tests puts the
    REFERENCE's package of the same name (by design) first on sys.path."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gym_ran_slice_b200_shim", os.path.join(root, "gym_ran_slice", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shim_reads_the_scenario_off_a_nodeb_like_object():
    g = _shim()
    assert g.ENV_ID == "RanSlice-v1"
    l1 = lambda t, n: types.SimpleNamespace(type=t, slices_ran=[object()] * n)
    node = types.SimpleNamespace(n_prbs=150, slots_per_step=25, slices_l1=[l1("eMBB", 1)] * 3 + [l1("mMTC", 1)] * 2)
    spec = g._spec_from_node_b(node)
    assert spec.scenario == {"n_prbs": 150, "n_embb": 3, "n_mmtc": 2} and spec.slots_per_step == 25 and spec.L1_level
    mux = types.SimpleNamespace(n_prbs=70, slices_l1=[l1("eMBB", 1), l1("mMTC", 1)])
    assert g._spec_from_node_b(mux).scenario == {"n_prbs": 70, "n_embb": 1, "n_mmtc": 1}
    assert not g._spec_from_node_b(types.SimpleNamespace(n_prbs=200, slices_l1=[l1("eMBB", 5)])).L1_level
    assert g._spec_from_node_b({"scenario": 3, "seed": 9}).seed == 9
    with pytest.raises(ValueError):
        g.make("Other-v0")
    with pytest.raises(TypeError):
        g._spec_from_node_b(object())


@pytest.mark.gpu
def test_make_returns_the_native_env_with_the_reference_kwargs(golden):
    g = _shim()
    gold = golden("B_scn3")
    env = g.make("gym_ran_slice:RanSlice-v1", node_b=g.NodeBSpec(scenario=3, seed=int(gold["base_seed"])), penalty=100)
    assert np.array_equal(env.reset(), gold["obs0"][0])
    for t in range(20):
        obs, rew, done, info = env.step(gold["actions"][0, t])
        assert np.array_equal(obs, gold["obs"][0, t]) and rew == gold["reward"][0, t] and done is False
    env2 = g.make(node_b={"scenario": 0, "seed": 1}, penalty=1000)
    env2.reset()
    o, r, d, info = env2.step(np.array([1, 1, 1, 1, 1]))
    assert env2.penalty == 1000 and o.shape == (50,)
