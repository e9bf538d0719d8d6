"""tools/pf_stats_oracle.py (CPU, oracle built with -DORC_PF_STATS): the what-if of the batched ProportionalFair step that
sized the warp kernel's RB loop keeps building and keeps its claim -- fewer warp-wide steps than the
chunk-by-chunk loop -- and the instrumentation does not change what the oracle computes."""
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pf_what_if_builds_and_predicts_fewer_steps(tmp_path):
    so = str(tmp_path / "liboracle_pfstats.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-DORC_PF_STATS", "-DORC_PF_SPEC=8",
                           "-shared", "-pthread", "-o", so, os.path.join(ROOT, "oracle", "ranslice_oracle.c"),
                           os.path.join(ROOT, "oracle", "kbrl_oracle.c"), "-lm"])
    env = dict(os.environ, RANSLICE_ORACLE_LIB=so)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pf_stats_oracle.py"), "--mux", "--envs", "2", "--burn", "150", "--steps", "10"],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"batched what-if: (\d+) warp-wide steps .* current loop takes (\d+)\)", out.stdout)
    assert m, out.stdout
    batched, current = int(m.group(1)), int(m.group(2))
    assert 0 < batched < current / 3, (batched, current)      # ~10x fewer in the multiplexed L1 at steady state; 3x is a floor here
    # same trajectory with and without the instrumentation
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import oracle_lib as ol; "
            "from ranslice_b200.tables import load_tables; b = ol.OracleBatch(load_tables(), 0, 3, 77, l1_mux=True); b.reset(); "
            "rng = np.random.default_rng(1); acc = 0.0\n"
            "for t in range(40):\n o, r, l, v, f = b.step(rng.integers(60, 200, (3, 1))); acc += float(o.astype(np.float64).sum()) + float(r.sum())\n"
            "print(repr(acc))") % (ROOT, os.path.join(ROOT, "tests"))
    a = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    b = subprocess.run([sys.executable, "-c", code], env=os.environ.copy(), capture_output=True, text=True, timeout=600)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr[-1000:], b.stderr[-1000:])
    assert a.stdout == b.stdout and np.isfinite(float(a.stdout))
