"""Pins the CPU oracle (oracle/ranslice_oracle.c) against fixtures generated from the UNMODIFIED
reference (tools/make_golden.py):  golden A = native numpy seeding replayed through Generator
callbacks, golden B = Philox stream injection, known answers = leaf functions."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol


@pytest.mark.parametrize("name,scn", [("A_scn0", 0), ("A_scn1", 1), ("A_scn3", 3)])
def test_golden_A_native_seeding(tables, golden, name, scn):
    g = golden(name)
    assert str(g["numpy_version"]) == np.__version__, "golden A is tied to the numpy Generator streams"
    seed = int(g["seed"])
    np.random.seed(seed)
    env = ol.OracleEnv(tables, scn, 0, numpy_rng=np.random.default_rng(seed))
    for _ in range(3):      # create_env runs the mMTC reset 3x before the caller's reset
        env.reset()
    assert np.array_equal(env.reset(), g["obs0"])
    for t in range(len(g["actions"])):
        obs, rew, lab, vio, acc, flags = env.step(g["actions"][t])
        assert flags == 0
        assert np.array_equal(obs, g["obs"][t]), "obs diverged at step %d" % t
        assert rew == g["reward"][t]
        assert np.array_equal(lab, g["labels"][t]) and np.array_equal(vio, g["violations"][t])
        assert np.array_equal(acc, g["acc"][t]), "raw accumulators diverged at step %d" % t


@pytest.mark.parametrize("name,scn", [("B_scn0", 0), ("B_scn1", 1), ("B_scn3", 3)])
def test_golden_B_philox_streams(tables, golden, name, scn):
    g = golden(name)
    E, T, S = g["actions"].shape
    b = ol.OracleBatch(tables, scn, E, int(g["base_seed"]), n_threads=2)
    assert np.array_equal(b.reset(), g["obs0"])
    for t in range(T):
        obs, rew, lab, vio, flags = b.step(g["actions"][:, t])
        assert not flags.any()
        assert np.array_equal(obs, g["obs"][:, t]), "obs diverged at step %d" % t
        assert np.array_equal(rew, g["reward"][:, t])
        assert np.array_equal(lab, g["labels"][:, t]) and np.array_equal(vio, g["violations"][:, t])


@pytest.mark.parametrize("name,scn", [("B_mux0", 0), ("B_mux3", 3), ("B_mux1", 1), ("B_mux2", 2)])
def test_golden_multiplexed_l1(tables, golden, name, scn):
    """create_env(L1_level=False) (scenario_creator.py:168-177): the eMBB RAN slices share ONE L1 scheduler.
    Oracle (l1_mux) vs the unmodified reference with injected Philox streams: obs, reward, per-L1 labels / violation
    counts (0..n_embb) and the raw accumulators of every RAN slice, bit-exact.  (The CUDA path does not implement this
    mode yet and rejects it; this pins the semantics for it.)"""
    g = golden(name)
    E, T, S = g["actions"].shape
    for e in range(E):
        env = ol.OracleEnv(tables, scn, int(g["base_seed"]), env_id=e, l1_mux=True)
        assert env.S == S
        assert np.array_equal(env.reset(), g["obs0"][e])
        for t in range(T):
            obs, rew, lab, vio, acc, flags = env.step(g["actions"][e, t])
            assert not flags
            assert np.array_equal(obs, g["obs"][e, t]), "obs diverged at env %d step %d" % (e, t)
            assert rew == g["reward"][e, t]
            assert np.array_equal(lab, g["labels"][e, t]) and np.array_equal(vio, g["violations"][e, t])
            assert np.array_equal(env.acc_ran(), g["acc_ran"][e, t]), (e, t)
    if scn == 0:
        assert g["violations"].max() > 1            # several RAN slices of one L1 violating in the same period
    if scn in (1, 2):                               # several mMTC RAN slices in ONE SliceL1mMTC (one queue, one action entry)
        assert S == 2 and g["violations"][:, :, 1].max() > 1


def test_known_answers_mcs_lut(tables, golden):
    k = golden("known_answers")
    tb = ol.c_tables(tables)
    for e, m, b, r in zip(k["e_snr"], k["lut_mcs"], k["lut_bps"], k["lut_rate"]):
        mcs, bps, rate = C.c_int(), C.c_double(), C.c_int()
        ol.lib().orc_mcs_lut(C.byref(tb), int(e), C.byref(mcs), C.byref(bps), C.byref(rate))
        assert (mcs.value, bps.value, rate.value) == (int(m), float(b), int(r))
    # SURVEY B.2 spot checks (truncation of the CSV's 9-digit rates matters)
    lut = dict(zip(k["e_snr"].tolist(), k["lut_rate"].tolist()))
    assert lut[12] == 526 and lut[15] == 671 and lut[18] == 789 and lut[-3] == 63 and lut[20] == 853


def test_known_answers_response(tables, golden):
    k = golden("known_answers")
    tb = ol.c_tables(tables)
    for snr, m, want in zip(k["resp_in"], k["resp_mcs"], k["resp_out"]):
        n = int(np.isfinite(snr).sum())
        v = np.ascontiguousarray(snr[:n])
        got = ol.lib().orc_response(C.byref(tb), int(m), v.ctypes.data_as(C.c_void_p), n)
        assert got == pytest.approx(want, rel=1e-12, abs=1e-300)


def test_known_answers_nominal_sinr(golden):
    """macro_cell through a Philox stream: oracle transforms + formula vs the reference's formula."""
    from ranslice_b200 import philox as px
    k = golden("known_answers")
    for name, (A, B) in (("nominal_2GHz", (128.1, 37.6)), ("nominal_900MHz", (120.9, 37.6)), ("nominal_rural", (95.5, 34.1))):
        st = px.PhiloxStream(4242, 3, px.STREAM_CHAN)
        lines = [((0, 0.5), (0.25, 0)), ((0.75, 0), (1, 0.5)), ((0, 0.5), (0.25, 1)), ((0.75, 1), (1, .5))]

        def fy(l, x):
            (x1, y1), (x2, y2) = l
            m = (y2 - y1) / (x2 - x1)
            return m * x + (-m * x1 + y1)
        for want in k[name]:
            while True:
                x, y = st.random(), st.random()
                if y > fy(lines[0], x) and y > fy(lines[1], x) and y < fy(lines[2], x) and y < fy(lines[3], x):
                    break
            logf = st.normal(0, 10)
            got = ol.lib().orc_macro_cell(x, y, logf, A, B)
            assert got == pytest.approx(float(want), rel=1e-13)
    assert np.array_equal(k["nominal_2GHz_ctr"][:5], [4, 10, 14, 18, 22])


def test_philox_known_answer():
    """Random123 KAT for philox4x32-10 (oracle C, Python contract)."""
    from ranslice_b200 import philox as px
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert px.philox4x32_10(*ctr, *key) == want
        c = (C.c_uint32 * 4)(*ctr)
        kk = (C.c_uint32 * 2)(*key)
        out = (C.c_uint32 * 4)()
        ol.lib().orc_philox(c, kk, out)
        assert tuple(out) == want


@pytest.mark.needs_reference
def test_reference_live_matches_golden_B_prefix(golden):
    """Re-runs the unmodified reference (Philox injection) for a few steps: the committed fixture is
    reproducible from the reference tree."""
    import refharness as rh
    g = golden("B_scn3")
    env, _ = rh.make_env_philox(int(g["base_seed"]), 3)
    tr = rh.run_trace(env, g["actions"][0, :25])
    assert np.array_equal(tr["obs"], g["obs"][0, :25])
    assert np.array_equal(tr["reward"], g["reward"][0, :25])
