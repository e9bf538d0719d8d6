"""Pins the KBRL oracle (oracle/kbrl_oracle.c) against controller trajectories recorded from the
UNMODIFIED reference (tools/make_golden_kbrl.py): every decision (selected action, adjusted flag, hits,
dictionary sizes, security factors, margins) must match exactly; dictionaries within 1e-9 relative
(numpy routes its dot products through BLAS, whose summation order is unspecified)."""
import numpy as np
import pytest

import oracle_lib as ol


def check_final_dictionaries(g, learner, rtol, atol_cf, atol_ki):
    """final landmarks / coefficients / K^-1 of every learner against the fixture (K_long keeps K^-1 in full only for
    the small dictionaries; diagonal and row sums for all)."""
    full = set(g["kinv_full_learners"].tolist()) if "kinv_full_learners" in g else None
    for s in range(len(g["dims"])):
        lm, cf, ki = learner(s)
        D, d = len(cf), int(g["dims"][s])
        assert D == int(g["sizes"][-1][s])
        assert np.array_equal(lm, g["final_landmarks"][s, :D, :d])
        assert np.allclose(cf, g["final_coeff"][s, :D], rtol=rtol, atol=atol_cf)
        if full is None:
            assert np.allclose(ki, g["final_kinv"][s, :D, :D], rtol=rtol, atol=atol_ki)
        else:
            if s in full:
                assert np.allclose(ki, g["final_kinv"][sorted(full).index(s), :D, :D], rtol=rtol, atol=atol_ki)
            scale = np.abs(g["kinv_diag"][s, :D]).max()
            assert np.allclose(np.diag(ki), g["kinv_diag"][s, :D], rtol=rtol, atol=atol_ki * scale)
            assert np.allclose(ki.sum(axis=1), g["kinv_rowsum"][s, :D], rtol=rtol, atol=atol_ki * scale * D)


@pytest.mark.parametrize("name", ["K_scn0", "K_scn1", "K_tie", "K_plus", "K_long0"])
def test_kbrl_oracle_replays_reference_controller(golden, name):
    g = golden(name)
    kb = ol.OracleKBRL(g["dims"], int(g["n_prbs"]), g["init_action"], g["init_sec"], tuple(g["accuracy_range"]), float(g["alfa"]),
                       tie_seed=int(g["seed"]), plus=name == "K_plus")
    for t in range(len(g["state"])):
        hits = kb.update_control(g["state"][t], g["action"][t], g["labels"][t])
        assert np.array_equal(hits, g["hits"][t]), t
        assert np.array_equal(kb.control()["sizes"], g["sizes"][t]), t
        a, adj = kb.select_action(g["new_state"][t])
        c = kb.control()
        assert np.array_equal(a, g["next_action"][t]) and adj == g["adjusted"][t], t
        assert np.array_equal(c["security_factors"], g["security_factors"][t]) and np.array_equal(c["margins"], g["margins"][t]), t
    assert c["tie_breaks"] == int(g["tie_calls"])          # every np.random.choice of the reference, and no other
    if name == "K_tie":
        assert c["tie_breaks"] > 20
    assert np.allclose(c["accuracies"], g["accuracies"], rtol=0, atol=1e-15)
    check_final_dictionaries(g, kb.learner, 1e-9 if name != "K_long0" else 1e-7, 1e-12 if name != "K_long0" else 1e-9, 1e-9 if name != "K_long0" else 1e-7)
