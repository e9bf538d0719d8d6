"""Pins the KBRL oracle (oracle/kbrl_oracle.c) against controller trajectories recorded from the
UNMODIFIED reference (tools/make_golden_kbrl.py): every decision (selected action, adjusted flag, hits,
dictionary sizes, security factors, margins) must match exactly; dictionaries within 1e-9 relative
(numpy routes its dot products through BLAS, whose summation order is unspecified)."""
import numpy as np
import pytest

import oracle_lib as ol


@pytest.mark.parametrize("name", ["K_scn0", "K_scn1"])
def test_kbrl_oracle_replays_reference_controller(golden, name):
    g = golden(name)
    kb = ol.OracleKBRL(g["dims"], int(g["n_prbs"]), g["init_action"], g["init_sec"], tuple(g["accuracy_range"]), float(g["alfa"]))
    for t in range(len(g["state"])):
        hits = kb.update_control(g["state"][t], g["action"][t], g["labels"][t])
        assert np.array_equal(hits, g["hits"][t]), t
        assert np.array_equal(kb.control()["sizes"], g["sizes"][t]), t
        a, adj = kb.select_action(g["new_state"][t])
        c = kb.control()
        assert np.array_equal(a, g["next_action"][t]) and adj == g["adjusted"][t], t
        assert np.array_equal(c["security_factors"], g["security_factors"][t]) and np.array_equal(c["margins"], g["margins"][t]), t
    assert c["tie_breaks"] == 0 and int(g["tie_calls"]) == 0
    assert np.allclose(c["accuracies"], g["accuracies"], rtol=0, atol=1e-15)
    for s in range(len(g["dims"])):
        lm, cf, ki = kb.learner(s)
        D, d = len(cf), int(g["dims"][s])
        assert np.array_equal(lm, g["final_landmarks"][s, :D, :d])
        assert np.allclose(cf, g["final_coeff"][s, :D], rtol=1e-9, atol=1e-12)
        assert np.allclose(ki, g["final_kinv"][s, :D, :D], rtol=1e-9, atol=1e-9)
