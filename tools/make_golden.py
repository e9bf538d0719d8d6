#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  Usage:
    python tools/make_golden.py [name ...]        # default: all
Fixtures (all compressed .npz, numpy version recorded inside):
  A_scn{0,1,3}  native numpy seeding (default_rng(0) + np.random.seed(0)), single env
  B_scn{0,1,3}  Philox-stream injection (tests/refharness.make_env_philox), several envs
  B_mux{0,3}    the same with create_env(L1_level=False): eMBB RAN slices multiplexed in one L1 scheduler
  known_answers leaf-function tables (MCS LUT, response(), macro_cell, constants)
"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden")

SCN = {0: (5, 200), 1: (5, 150), 2: (5, 100), 3: (2, 70)}   # index -> (S, n_prbs)

A_CASES = {"A_scn0": (0, 0, 2000), "A_scn1": (1, 0, 400), "A_scn3": (3, 0, 2000)}
B_CASES = {"B_scn0": (0, 7000, 8, 300), "B_scn1": (1, 7100, 4, 200), "B_scn3": (3, 7200, 8, 300)}
# create_env(L1_level=False): the eMBB RAN slices multiplexed in one L1 (scenario_creator.py:168-177); name -> (scn, seed, envs, steps)
M_CASES = {"B_mux0": (0, 7300, 4, 200), "B_mux3": (3, 7400, 2, 120),
           # scenarios with several mMTC slices: they too share ONE L1 (one queue, one action entry)
           "B_mux1": (1, 7500, 3, 150), "B_mux2": (2, 7600, 3, 150)}


def gen_A(name):
    import refharness as rh
    scn, seed, steps = A_CASES[name]
    S, n_prbs = SCN[scn]
    env, _ = rh.make_env_native(seed, scn)
    act = rh.simplex_actions(seed, S, n_prbs, steps)
    tr = rh.run_trace(env, act)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), scenario=scn, seed=seed, actions=act,
                        numpy_version=np.__version__, **tr)
    return name, int(tr["violations"].sum()), float(tr["reward"].mean())


def _gen_B_env(args):
    import refharness as rh
    scn, base, e, steps = args
    S, n_prbs = SCN[scn]
    env, _ = rh.make_env_philox(base, scn, env_id=e)                # key = base seed, counter word 3 = env id
    act = rh.simplex_actions(base + e, S, n_prbs, steps)
    tr = rh.run_trace(env, act)
    tr["actions"] = act
    return tr


def _gen_M_env(args):
    import refharness as rh
    scn, base, e, steps = args
    seed = base + e
    S = 1 + (1 if scn in (1, 2, 3) else 0)                   # L1 slices: one multiplexed eMBB L1 (+ one mMTC L1)
    env, _ = rh.make_env_philox(base, scn, env_id=e, L1_level=False)
    assert env.n_slices == S
    act = rh.simplex_actions(seed, S, SCN[scn][1], steps)
    act[::9] = act[::9] // 8                                 # starved periods: deep backlogs, contended PF over all RAN slices
    if scn in (1, 2):
        act[:, 1] = np.minimum(act[:, 1], 3 + (np.arange(steps) % 11))   # few carriers: the shared mMTC queue builds up across slices
    tr = rh.run_trace(env, act)
    tr["actions"] = act
    if "acc_ran" not in tr:                                  # single RAN slice per L1 (scenario_3): same rows as acc
        tr["acc_ran"] = tr["acc"].copy()
    return tr


def gen_M(name, pool):
    scn, base, n_envs, steps = M_CASES[name]
    trs = list(pool.map(_gen_M_env, [(scn, base, e, steps) for e in range(n_envs)]))
    stacked = {k: np.stack([t[k] for t in trs]) for k in trs[0]}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), scenario=scn, base_seed=base, l1_level=False,
                        numpy_version=np.__version__, **stacked)
    return name, int(stacked["violations"].sum()), float(stacked["reward"].mean())


def gen_B(name, pool):
    scn, base, n_envs, steps = B_CASES[name]
    trs = list(pool.map(_gen_B_env, [(scn, base, e, steps) for e in range(n_envs)]))
    stacked = {k: np.stack([t[k] for t in trs]) for k in trs[0]}   # [E, T, ...]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), scenario=scn, base_seed=base,
                        numpy_version=np.__version__, **stacked)
    return name, int(stacked["violations"].sum()), float(stacked["reward"].mean())


def gen_known():
    import refharness as rh
    from ranslice_b200 import philox as px
    ref = rh.load_reference()
    cm = ref.channel_models
    mcs = cm.MCSCodeset()
    e = np.arange(-40, 41)
    lut_mcs = np.zeros(len(e), np.int32)
    lut_bps = np.zeros(len(e), np.float64)
    for i, s in enumerate(e):
        lut_mcs[i], lut_bps[i] = mcs.mcs_rate_vs_error(int(s), 0.1)
    lut_rate = np.array([int(158 * b) for b in lut_bps], np.int32)     # schedulers.py:45 truncation
    rng = np.random.default_rng(99)
    resp_in, resp_mcs, resp_out = [], [], []
    for _ in range(400):
        n = int(rng.integers(1, 41))
        m = int(rng.integers(0, 26))
        snr = rng.normal(mcs.snr[m], 6.0, size=n)
        p = mcs.response(m, snr)
        pad = np.full(40, np.nan)
        pad[:n] = snr
        resp_in.append(pad)
        resp_mcs.append(m)
        resp_out.append(float(np.asarray(p).reshape(-1)[0]))
    # macro_cell through a Philox stream (the transform definitions are ours; the formula is theirs)
    nom = {}
    for pname in ("macro_cell_urban_2GHz", "macro_cell_urban_900MHz", "macro_cell_rural"):
        st = px.PhiloxStream(4242, 3, px.STREAM_CHAN)
        gen = cm.NominalSINR(st, pname)
        vals, ctr = [], []
        for _ in range(200):
            vals.append(gen.generate())
            ctr.append(st.n)
        nom[pname] = (np.array(vals), np.array(ctr))
    np.savez_compressed(
        os.path.join(OUT, "known_answers.npz"), numpy_version=np.__version__,
        A=mcs.A, B=mcs.B, e_snr=e, lut_mcs=lut_mcs, lut_bps=lut_bps, lut_rate=lut_rate,
        snr_ref=mcs.snr, rate=mcs.rate, order=mcs.order,
        modulation=np.array([{"qpsk": 0, "16qam": 1, "64qam": 2}[m] for m in mcs.modulation]),
        resp_in=np.array(resp_in), resp_mcs=np.array(resp_mcs), resp_out=np.array(resp_out),
        nominal_2GHz=nom["macro_cell_urban_2GHz"][0], nominal_2GHz_ctr=nom["macro_cell_urban_2GHz"][1],
        nominal_900MHz=nom["macro_cell_urban_900MHz"][0],
        nominal_rural=nom["macro_cell_rural"][0])
    return "known_answers", 0, 0.0


def main():
    names = sys.argv[1:] or (["known_answers"] + list(A_CASES) + list(B_CASES) + list(M_CASES))
    os.makedirs(OUT, exist_ok=True)
    with ProcessPoolExecutor(8) as pool:
        futs = []
        for n in names:
            if n in A_CASES:
                futs.append(pool.submit(gen_A, n))
            elif n == "known_answers":
                futs.append(pool.submit(gen_known))
        for n in names:
            if n in B_CASES:
                print(gen_B(n, pool), flush=True)
            elif n in M_CASES:
                print(gen_M(n, pool), flush=True)
        for f in futs:
            print(f.result(), flush=True)


if __name__ == "__main__":
    main()
