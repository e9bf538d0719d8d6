"""Kernel #2 on the GPU (through the C ABI) against the reference fixtures and the pinned oracle."""
import numpy as np
import pytest

import oracle_lib as ol
from test_kbrl_oracle import check_final_dictionaries

pytestmark = pytest.mark.gpu

FIXTURES = ["K_scn0", "K_scn1", "K_tie", "K_plus", "K_long0"]
TOL = {"K_long0": (1e-6, 1e-8, 1e-6)}       # rtol, atol coeff, atol K^-1 (3000 steps, dictionaries of 500 landmarks)


def _learners(g, name, n_envs=1, dict_cap=1024):
    from ranslice_b200.kbrl import BatchedProjectron
    return BatchedProjectron(int(g["scenario"]), n_envs, dict_cap=dict_cap, tie_seed=int(g["seed"]),
                             algorithm="projectron_plus" if name == "K_plus" else "projectron")


def _control(g, name, n_envs=1, dict_cap=1024):
    from ranslice_b200.kbrl import KBRLControl
    lrn = _learners(g, name, n_envs, dict_cap)
    return KBRLControl(lrn, int(g["n_prbs"]), g["init_action"], g["init_sec"], alfa=float(g["alfa"]),
                       accuracy_range=tuple(g["accuracy_range"]))


@pytest.mark.parametrize("name", FIXTURES)
def test_controller_replays_reference_fixture(golden, name):
    """CUDA Projectron + host mirror of KBRL_Control vs the unmodified reference controller, step by step.  K_long0 runs
    the reference's own horizon (3000 steps, dictionaries of ~500 landmarks: pooled growth, both update launches and the
    hand-over between them), K_tie the random tie-break of kernel.py:26-27, K_plus ProjectronPlus."""
    g = golden(name)
    ctl = _control(g, name)
    for t in range(len(g["state"])):
        hits = ctl.update_control(g["state"][t][None], g["action"][t][None], g["labels"][t][None])
        assert np.array_equal(hits[0], g["hits"][t]), t
        assert np.array_equal(ctl.learners.sizes()[0][0], g["sizes"][t]), t
        a, adj = ctl.select_action(g["new_state"][t][None])
        ctl.adjusted = adj                       # run(): action, self.adjusted = self.select_action(new_state)
        assert np.array_equal(a[0], g["next_action"][t]) and adj[0] == g["adjusted"][t], t
        assert np.array_equal(ctl.security_factors[0], g["security_factors"][t]), t
        assert np.array_equal(ctl.margins[0], g["margins"][t]), t
    assert np.allclose(ctl.accuracies[0], g["accuracies"], rtol=0, atol=1e-15)
    check_final_dictionaries(g, lambda s: ctl.learners.learner(0, s), *TOL.get(name, (1e-8, 1e-11, 1e-8)))
    assert not ctl.learners.sizes()[1].any()                  # no dictionary cap / pool flag
    pool = ctl.learners.pool()
    assert pool["tie_breaks"] == int(g["tie_calls"])          # every np.random.choice of the reference, and no other
    assert pool["max_dictionary"] == int(g["sizes"].max())
    rows = (g["sizes"][-1] + 31) // 32                        # bump allocator: exactly the tile rows the dictionaries hold
    assert pool["used_bytes"] == 8 * int(sum(544 * r + 1024 * r * (r + 1) // 2 for r in rows))


def test_batched_learners_match_oracle_on_synthetic_streams():
    """64 envs x 5 learners driven by synthetic (state, action, label) streams: every env's decisions equal
    the oracle's for that env (learners are independent; batching must not mix them)."""
    from ranslice_b200.kbrl import BatchedProjectron, KBRLControl
    N, S, n_prbs, T = 64, 5, 200, 60
    rng = np.random.default_rng(11)
    ia = rng.integers(4, 20, (N, S)); sec = rng.integers(2, 8, (N, S))
    ctl = KBRLControl(BatchedProjectron(0, N, dict_cap=128, tie_seed=5), n_prbs, ia, sec, accuracy_range=(0.97, 0.99))
    orcs = [ol.OracleKBRL([11] * S, n_prbs, ia[e], sec[e], (0.97, 0.99), tie_seed=5, env_id=e) for e in range(N)]
    state = rng.random((N, 50)).astype(np.float32)
    action = ia.copy()
    for t in range(T):
        new_state = np.clip(state + rng.normal(0, 0.05, state.shape), 0, 1.5).astype(np.float32)
        need = (new_state.reshape(N, S, 10)[:, :, 0] * 60).astype(np.int64)       # synthetic SLA rule
        labels = np.where(action >= need, 1, -1)
        hits = ctl.update_control(state, action, labels)
        nxt, adj = ctl.select_action(new_state)
        ctl.adjusted = adj
        for e in range(N):
            h = orcs[e].update_control(state[e], action[e], labels[e])
            a, ad = orcs[e].select_action(new_state[e])
            assert np.array_equal(h, hits[e]) and np.array_equal(a, nxt[e]) and ad == adj[e], (t, e)
        action, state = nxt, new_state
    sizes = ctl.learners.sizes()[0]
    assert np.array_equal(sizes, np.array([o.control()["sizes"] for o in orcs]))
    assert sizes.max() > 3


def test_dictionary_cap_and_pool_exhaustion_are_flagged():
    """dict_cap bounds ONE dictionary (flag 1); an exhausted pool drops the sample and flags 2 -- never a crash."""
    from ranslice_b200.kbrl import BatchedProjectron
    rng = np.random.default_rng(0)
    lrn = BatchedProjectron(0, 2, dict_cap=32)
    for t in range(120):
        st = (rng.random((2, 50)) * 3).astype(np.float32)
        lrn.update(st, rng.integers(0, 200, (2, 5)), rng.choice([-1, 1], (2, 5)))
    sizes, flags = lrn.sizes()
    assert (sizes == 32).all() and (flags & 1).all() and not (flags & 2).any()
    lrn.close()
    lrn = BatchedProjectron(0, 64, dict_cap=1024, pool_mb=1)      # 1 MiB for 320 learners: 12.6 KB per first tile row -> ~83 rows
    for t in range(40):
        st = (rng.random((64, 50)) * 3).astype(np.float32)
        lrn.update(st, rng.integers(0, 200, (64, 5)), rng.choice([-1, 1], (64, 5)))
    sizes, flags = lrn.sizes()
    pool = lrn.pool()
    assert (flags & 2).any() and pool["used_bytes"] <= pool["total_bytes"] == 1 << 20
    lrn.close()


def test_kbrl_in_the_loop_with_the_env():
    """BASELINE config 3 shape: env step + KBRL update/select every step (small batch); dictionaries grow,
    allocations respect sum(action) <= n_prbs, result keys match kbrl_control.py:148-155."""
    from ranslice_b200 import create_batched_env
    from ranslice_b200.kbrl import create_kbrl_agent
    N = 32
    env = create_batched_env(99, 0, N)
    agent = create_kbrl_agent(np.random.default_rng(0), 0, accuracy_range=(0.97, 0.99), n_envs=N, dict_cap=128)
    out = agent.run(env, 40)
    assert set(out) == {"reward", "resources", "hits", "adjusted", "SLA", "violation"}
    assert out["hits"].shape == (N, 5, 40) and (out["resources"] <= 200).all()
    assert agent.learners.sizes()[0].max() > 2
    env.close()


# ------------------------------------------------------------------------------------------------ device-resident controller
def _device_control(g, name, n_envs=1, dict_cap=1024):
    from ranslice_b200.kbrl import DeviceKBRLControl
    lrn = _learners(g, name, n_envs, dict_cap)
    return DeviceKBRLControl(lrn, int(g["n_prbs"]), g["init_action"], g["init_sec"], alfa=float(g["alfa"]),
                             accuracy_range=tuple(g["accuracy_range"]))


@pytest.mark.parametrize("name", FIXTURES)
def test_device_controller_replays_reference_fixture(golden, name):
    """kb_control_update_device / kb_control_select_device (controller state in HBM) vs the unmodified reference
    KBRL_Control, step by step: hits, next action, adjusted flag, security factors, margins, accuracies."""
    import torch
    g = golden(name)
    ctl = _device_control(g, name)
    dev = ctl.device
    tt = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a[None]).astype(dt)).to(dev)
    for t in range(len(g["state"])):
        hits = ctl.update_control(tt(g["state"][t], np.float32), tt(g["action"][t], np.int32), tt(g["labels"][t], np.int32))
        assert np.array_equal(hits.cpu().numpy()[0], g["hits"][t]), t
        a, adj = ctl.select_action(tt(g["new_state"][t], np.float32))
        assert np.array_equal(a.cpu().numpy()[0], g["next_action"][t]) and int(adj[0]) == g["adjusted"][t], t
        if t % 25 == 0 or t == len(g["state"]) - 1:
            cs = ctl.control_state()
            assert np.array_equal(cs["security_factors"][0], g["security_factors"][t]), t
            assert np.array_equal(cs["margins"][0], g["margins"][t]), t
            assert np.array_equal(ctl.learners.sizes()[0][0], g["sizes"][t]), t
    assert np.allclose(ctl.control_state()["accuracies"][0], g["accuracies"], rtol=0, atol=1e-15)
    assert ctl.learners.pool()["tie_breaks"] == int(g["tie_calls"]) and not ctl.learners.sizes()[1].any()


def test_device_controller_matches_host_mirror_in_the_loop():
    """run() of the device-resident controller == run() of the numpy mirror on the same envs and seeds
    (the mirror is pinned to the oracle / reference above): every history array is identical."""
    from ranslice_b200 import create_batched_env
    from ranslice_b200.kbrl import create_kbrl_agent
    N, T = 48, 60
    outs = []
    for resident in (False, True):
        env = create_batched_env(123, 0, N)
        agent = create_kbrl_agent(np.random.default_rng(5), 0, accuracy_range=(0.97, 0.99), n_envs=N, dict_cap=128,
                                  resident=resident)
        outs.append((agent.run(env, T), agent.learners.sizes()[0].copy()))
        env.close()
    (a, sa), (b, sb) = outs
    for k in a:
        assert np.array_equal(np.asarray(a[k], np.float64), np.asarray(b[k], np.float64)), k
    assert np.array_equal(sa, sb) and sa.max() > 2


def test_guarded_fp32_evaluation_never_changes_a_decision():
    """The fp32 fast path of f(x) (csrc/kbrl.cu eval_f_guarded) against all-fp64 evaluation (kb_set_exact) in the
    closed loop with the env, 512 envs x 120 steps (~2.5e8 kernel evaluations): identical actions, hits and
    dictionaries.  A single flipped sign would change a dictionary and diverge the two runs."""
    import torch
    from ranslice_b200 import create_batched_env
    from ranslice_b200.kbrl import create_kbrl_agent
    N, T = 512, 120
    outs = []
    for exact in (False, True):
        env = create_batched_env(777, 0, N)
        agent = create_kbrl_agent(np.random.default_rng(3), 0, accuracy_range=(0.97, 0.99), n_envs=N, dict_cap=128,
                                  resident=True)
        agent.learners.set_exact(exact)
        out = agent.run(env, T)
        lm, cf, _ = agent.learners.learner(7, 2)
        outs.append((out, agent.learners.sizes()[0].copy(), lm.copy(), cf.copy()))
        env.close()
    (a, sa, la, ca), (b, sb, lb, cb) = outs
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(sa, sb) and np.array_equal(la, lb) and np.array_equal(ca, cb)
    assert sa.max() >= 20


def test_checkpoint_restores_learners_and_controller(golden):
    """kb_get_state / kb_set_state: replaying the second half of a reference fixture from a checkpoint taken in the middle,
    on a FRESH handle, gives the reference's decisions (dictionaries in the pool, tie counters, E-learner state)."""
    import torch
    g = golden("K_scn0")
    T = len(g["state"])
    ctl = _device_control(g, "K_scn0")
    dev = ctl.device
    tt = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a[None]).astype(dt)).to(dev)

    def steps(c, lo, hi):
        for t in range(lo, hi):
            hits = c.update_control(tt(g["state"][t], np.float32), tt(g["action"][t], np.int32), tt(g["labels"][t], np.int32))
            assert np.array_equal(hits.cpu().numpy()[0], g["hits"][t]), t
            a, adj = c.select_action(tt(g["new_state"][t], np.float32))
            assert np.array_equal(a.cpu().numpy()[0], g["next_action"][t]) and int(adj[0]) == g["adjusted"][t], t

    steps(ctl, 0, T // 2)
    blob = ctl.learners.get_state()
    fresh = _device_control(g, "K_scn0")
    fresh.learners.set_state(blob)
    assert np.array_equal(fresh.learners.sizes()[0], ctl.learners.sizes()[0])
    steps(fresh, T // 2, T)
    assert np.array_equal(fresh.learners.sizes()[0][0], g["sizes"][-1])
    assert np.allclose(fresh.control_state()["accuracies"][0], g["accuracies"], rtol=0, atol=1e-15)
    with pytest.raises(RuntimeError):
        fresh.learners.set_state(blob[:-8])
