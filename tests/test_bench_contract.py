"""bench.py contract on a CPU-only box: the reference arm prints exactly ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--burn-in", "3", "--no-python-reference"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("batched env-steps/sec") and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_python_reference_leg_times_the_unmodified_reference_when_staged():
    """cpu_baseline.python_reference: the reference's own create_env / step(), staged under baseline/_ref by build(); None when
    it has not been staged (the leg is reported, never required)."""
    sys.path.insert(0, ROOT)
    import bench
    r = bench.python_reference_throughput(0, burn_in=2, steps=3, timeout=300)
    if not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "node_b.py")):
        assert r is None
        return
    assert r["kind"] == "reference" and r["cores"] == 1 and r["value"] > 0, r


def test_rank_cpu_binding_is_a_no_op_for_a_single_rank():
    sys.path.insert(0, ROOT)
    import bench
    before = os.sched_getaffinity(0)
    info = bench.bind_rank_to_cpus(0, 1)
    assert os.sched_getaffinity(0) == before and info["mode"] == "none"
