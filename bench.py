#!/usr/bin/env python3
"""bench.py -- batched env-steps/s of the RAN-slicing step path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--scenario 0]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # CPU arm (oracle port on the host cores)

One "step" = one env.step() of every env of the batch (50 TTIs x all slices).  Workload at any N:
scenario_0 (5 eMBB slices, 200 PRBs), 65536 envs per GPU (the configuration the north_star target
is quoted on), random-simplex action policy (wrapper.py:77-82), population burned in for 600 steps
(30 s mean UE holding time) before anything is timed.  Envs shard across ranks by global env id;
there is no collective on the step path (weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SCN = {0: (5, 200, 50), 1: (5, 150, 36), 2: (5, 100, 22), 3: (2, 70, 13)}     # S, n_prbs, V
NOMINAL_BALG = {0: 154e3, 3: 27e3}       # SURVEY 8(d) nominal algorithmic bytes per env-step
BASE_SEED = 20260000
METRIC = "batched env-steps/sec (scenario_0)"


# stdout carries exactly ONE line (the JSON): everything else a library prints at C level (NCCL's version banner ...)
# is diverted to stderr by pointing fd 1 at fd 2 for the duration of the run.
_REAL_STDOUT = None


def _divert_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def simplex_actions(rng, N, S, n_prbs):
    w = rng.random((N, S + 1), dtype=np.float32)
    return np.floor(n_prbs * w[:, :S] / w.sum(axis=1, keepdims=True)).astype(np.int32)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self._out = os.fdopen(fd, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self._out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self._out.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path).read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_port_throughput(scn, n_envs, burn_in, steps, warmup, threads, seed0=BASE_SEED):
    """Oracle port (oracle/ranslice_oracle.c) on the host cores: the CPU restatement of the same path."""
    import oracle_lib as ol
    from ranslice_b200.tables import load_tables
    S, n_prbs, _ = SCN[scn]
    b = ol.OracleBatch(load_tables(), scn, n_envs, seed0, n_threads=threads)
    b.reset()
    rng = np.random.default_rng(1)
    for _ in range(burn_in + warmup):
        b.step(simplex_actions(rng, n_envs, S, n_prbs))
    acts = [simplex_actions(rng, n_envs, S, n_prbs) for _ in range(steps)]
    t0 = time.perf_counter()
    for a in acts:
        b.step(a)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


_PYREF_SNIPPET = r"""
import json, os, sys, time
sys.path.insert(0, %(tests)r); sys.path.insert(0, %(root)r)
import numpy as np
import refharness as rh
env, _ = rh.make_env_native(0, %(scn)d)                  # default_rng(0) + np.random.seed(0), like golden trace A
S, n_prbs = env.n_slices, env.n_prbs
act = rh.simplex_actions(0, S, n_prbs, %(burn)d + %(steps)d)
env.reset()
for t in range(%(burn)d):
    env.step(act[t])
t0 = time.perf_counter()
for t in range(%(burn)d, %(burn)d + %(steps)d):
    env.step(act[t])
dt = time.perf_counter() - t0
print(json.dumps({"value": %(steps)d / dt, "seconds": dt}))
"""


def python_reference_throughput(scn, burn_in=300, steps=20, timeout=400):
    """The UNMODIFIED Python reference (staged under baseline/_ref by __graft_entry__.build(); git-ignored, travels to the
    GPU box), single process, through its own create_env / step(): env-steps/s after `burn_in` steps.  None if absent."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref, "node_b.py")):
        return None
    code = _PYREF_SNIPPET % {"tests": os.path.join(ROOT, "tests"), "root": ROOT, "scn": scn, "burn": burn_in, "steps": steps}
    try:
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RANSLICE_REFERENCE=ref), capture_output=True,
                           text=True, timeout=timeout)
        out = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:                                    # noqa: BLE001 -- a baseline that cannot run is reported, not fatal
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    return {"value": out["value"], "unit": "env-steps/s", "cores": 1, "kind": "reference",
            "sample": "unmodified reference create_env(rng, %d).step(), single process, %d steps after %d burn-in steps "
                      "(native numpy seeding)" % (scn, steps, burn_in)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = 16 * threads
    burn = args.burn_in                                        # the same population burn-in as the GPU arm
    val, dt = cpu_port_throughput(args.scenario, n_envs, burn, args.steps, args.warmup, threads)
    sample = "%d envs x %d steps after %d burn-in steps, %d pthreads" % (n_envs, args.steps, burn, threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "scenario_%d, bounded CPU sample: %s" % (args.scenario, sample),
                       "note": "C port of the reference's algorithm (oracle/), pinned bit-exact to the reference, on all host "
                               "threads; the unmodified Python reference itself is timed single-process in "
                               "cpu_baseline.python_reference (it is ~1e4 x slower per core)"},
            "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_python_reference:
        line["cpu_baseline"]["python_reference"] = python_reference_throughput(args.scenario, steps=20)
    _emit(line)


# --------------------------------------------------------------------------------------------- GPU arm
def bind_rank_to_cpus(local, world):
    """One process per GPU: keep each rank (and the pinned buffers it first-touches) on CPUs of its own -- the cores local
    to its GPU when the cgroup allows them, else a private slice of the allowed set.  Returns a description."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        if world == 1 or os.environ.get("RS_BENCH_AFFINITY", "auto") == "none":
            return {"mode": "none", "cpus": len(allowed)}
        import torch
        prop = torch.cuda.get_device_properties(local)
        node, local_cpus = None, []
        bus = getattr(prop, "pci_bus_id", None)
        if bus is not None:
            path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), bus, getattr(prop, "pci_device_id", 0))
            try:
                node = int(open(path + "/numa_node").read())
                for part in open(path + "/local_cpulist").read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    local_cpus += list(range(int(lo), int(hi or lo) + 1))
            except Exception:
                pass
        near = [c for c in allowed if c in set(local_cpus)]
        pool = near if len(near) >= 2 else allowed
        ranks_sharing = world                                   # conservative: every rank may share this pool
        chunk = max(1, len(pool) // ranks_sharing)
        mine = pool[(local % ranks_sharing) * chunk:(local % ranks_sharing + 1) * chunk] or pool
        os.sched_setaffinity(0, mine)
        return {"mode": "near" if pool is near else "slice", "gpu_numa_node": node, "cpus": len(mine), "first_cpu": mine[0]}
    except Exception as e:                                     # noqa: BLE001
        return {"mode": "failed: %s" % type(e).__name__}


def measure(env, scn, K, warmup, burn_in, dev, rank, world, local, barrier, sample_clocks=True):
    """Device-resident, end-to-end (blocking and pipelined) and per-kernel numbers of one env batch; every timed region
    is bracketed by barrier() (synchronize + dist.barrier + synchronize)."""
    import torch
    S, n_prbs, V = SCN[scn]
    E = env.n_envs
    rng = np.random.default_rng(1000 + rank)
    n_act = warmup + K
    host_act = torch.empty((n_act, E, S), dtype=torch.int32, pin_memory=True)
    for i in range(n_act):
        host_act[i] = torch.from_numpy(simplex_actions(rng, E, S, n_prbs))
    dev_act = host_act.to(dev)
    out = env.step_device(dev_act[0])                       # allocates the output tensors (part of burn-in)
    for i in range(burn_in):                                # population burn-in (untimed set-up)
        env.step_device(dev_act[i % n_act], out)
    torch.cuda.synchronize()

    # ---- device-resident arm (inputs already in HBM)
    for i in range(warmup):
        env.step_device(dev_act[i], out)
    k0, _ = env.counters()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0 and sample_clocks:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        env.step_device(dev_act[warmup + i], out)
    ev1.record()
    barrier()
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    k1, trace_elems = env.counters()

    # ---- end-to-end arm: public host API, pinned host buffers, H2D + kernels + D2H of EVERY step inside the timed region.
    # (a) blocking call (step_host_inplace = rs_step): what an agent that needs step i's result before choosing action i+1
    # gets; (b) the pipelined call (step_host_async / wait = rs_step_async / rs_wait, two steps in flight: the D2H of step i
    # overlaps the kernels of step i+1): what a feedback-free policy (or an agent alternating two half-batches) gets.
    # Both are printed; e2e.value is the better of the two and says which.
    for i in range(warmup):
        env.step_host_inplace(host_act[i].numpy())
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        env.step_host_inplace(host_act[warmup + i].numpy())
    torch.cuda.synchronize()
    t_sync = time.perf_counter() - t0
    barrier()
    hbs = [env.alloc_host_buffers(), env.alloc_host_buffers()]
    pending = None
    for i in range(warmup):
        tk = env.step_host_async(host_act[i].numpy(), hbs[i & 1])
        if pending is not None:
            env.wait(pending)
        pending = tk
    env.wait(pending)
    barrier()
    t0 = time.perf_counter()
    pending = None
    for i in range(K):
        tk = env.step_host_async(host_act[warmup + i].numpy(), hbs[i & 1])
        if pending is not None:
            env.wait(pending)                                 # results of the previous step are now in host memory
        pending = tk
    env.wait(pending)
    t_pipe = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 and sample_clocks else None

    # ---- per-kernel durations (CUDA events on the launching stream, inside the library; kernels serialised on one stream)
    prof = env.profile_steps([dev_act[warmup + (i % K)] for i in range(min(K, 10))], out)
    n_live = int(env.n_ues().sum())
    return {"t_dev": t_dev, "t_sync": t_sync, "t_pipe": t_pipe, "launches": k1 - k0, "trace_elems": trace_elems, "prof": prof,
            "n_live": n_live, "clocks": clocks}


def roofline_blocks(env, scn, m, peaks, sm_mhz):
    """HBM roofline of the dominant kernel (algorithmic bytes of SURVEY 8d over its own duration) and the issue-slot roofline."""
    S, n_prbs, V = SCN[scn]
    E = env.n_envs
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    prof = m["prof"]
    # algorithmic bytes of one launch of the eMBB slice kernel: 2*B_state + B_io + B_trace, B_trace counted by the kernel
    Se = env.n_l1_embb
    b_state = Se * 16 * E + 88 * m["n_live"]
    b_io = (4 * Se + 40 * env.n_embb + 8 * Se) * E            # its action entries, obs columns, labels / violations
    b_trace = 4 * m["trace_elems"]
    b_alg = 2 * b_state + b_io + b_trace
    k_ms = prof["dominant_ms"] if prof["dominant_ms"] > 0 else prof["embb_ms"]
    achieved = b_alg / (k_ms * 1e-3) / 1e9
    tr = {}
    try:
        allp = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        for cand in (allp if isinstance(allp, list) else [allp]):
            if cand.get("envs_per_gpu") == E and cand.get("scenario") == scn:
                tr = cand
    except Exception:
        pass
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": tr.get("dram_bytes_per_launch"), "peak_source": peak_src, "kernel": prof["kernel"],
            "kernel_ms": k_ms, "kernel_ms_is": "embb_step_smem alone (event pair around that kernel; the sort pre-pass, the "
                                               "general kernel and the mMTC / reward kernels are in side_kernels_ms)",
            "side_kernels_ms": {k: prof[k] for k in ("sort_ms", "embb_rest_ms", "mmtc_scan_ms", "mmtc_step_ms", "reward_ms")},
            "bytes_per_launch": b_alg, "bytes_per_env_step": b_alg / E, "nominal_bytes_per_env_step": NOMINAL_BALG.get(scn),
            "note": "issue/latency-bound path (serial PF loop, fp64 decisions); tables are L2-resident, "
                    "so DRAM traffic is far below the algorithmic bytes (SURVEY 8d caveat); see issue_roofline"}
    if env.n_mmtc and prof["mmtc_scan_ms"] > 0:               # the one genuinely HBM-bound kernel of the path
        scan_bytes = 4.0 * 1000 * env.n_mmtc * E
        roof["mmtc_scan"] = {"achieved": scan_bytes / (prof["mmtc_scan_ms"] * 1e-3) / 1e9, "unit": "GB/s",
                             "frac": scan_bytes / (prof["mmtc_scan_ms"] * 1e-3) / 1e9 / peak, "bytes_per_launch": scan_bytes}
    issue = None
    if tr.get("warp_insts_per_launch") and sm_mhz:
        sms = tr.get("sm_count", 148)
        peak_issue = sms * 4 * sm_mhz * 1e6                   # one warp instruction per scheduler per cycle
        ach = tr["warp_insts_per_launch"] / (k_ms * 1e-3)
        issue = {"achieved": ach, "peak": peak_issue, "unit": "warp-inst/s", "frac": ach / peak_issue,
                 "lanes_per_inst": tr.get("lanes_per_inst"), "thread_frac": ach / peak_issue * (tr.get("lanes_per_inst") or 32) / 32,
                 "source": "warp instructions per launch from the committed ncu capture (%s), duration measured live"
                           % tr.get("source", "profiles/traffic.json")}
    return roof, issue


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--scenario", type=int, default=0)
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--burn-in", type=int, default=600)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs[1..3] side measurements")
    args = ap.parse_args()
    _divert_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = bind_rank_to_cpus(local, world)                # before any pinned allocation (first touch)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    warmup = max(args.warmup, 3)
    K = args.steps

    from ranslice_b200 import create_batched_env
    from ranslice_b200.sharding import max_over_ranks, weak_shard
    scn = args.scenario
    S, n_prbs, V = SCN[scn]
    first_env, E = weak_shard(args.envs_per_gpu, rank)      # weak scaling: fixed envs per GPU, global env ids
    env = create_batched_env(BASE_SEED, scn, E, device=local, first_env_id=first_env, kernel_variant=args.variant)
    env.reset()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m = measure(env, scn, K, warmup, args.burn_in, dev, rank, world, local, barrier)
    if world > 1:                                             # per-rank diagnostics (stderr): which shard is the slow one, and where
        sys.stderr.write("[bench rank %d] dev %.3f ms/step  blocking %.3f  pipelined %.3f  affinity %s\n"
                         % (rank, 1e3 * m["t_dev"] / K, 1e3 * m["t_sync"] / K, 1e3 * m["t_pipe"] / K, json.dumps(affinity)))
    t_dev, t_pipe, t_sync = max_over_ranks(m["t_dev"]), max_over_ranks(m["t_pipe"]), max_over_ranks(m["t_sync"])   # slowest shard decides
    total_envs = E * world

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        clocks = m["clocks"]
        roof, issue = roofline_blocks(env, scn, m, peaks, (clocks or {}).get("sm_mhz"))
        e2e_pipe, e2e_sync = total_envs * K / t_pipe, total_envs * K / t_sync
        best_pipe = e2e_pipe >= e2e_sync
        line = {
            "metric": METRIC, "value": total_envs * K / t_dev, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": warmup, "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "scenario_%d, %d envs per GPU (%d total), %d TTIs/step, random-simplex actions, "
                                   "%d-step population burn-in" % (scn, E, total_envs, env.slots_per_step, args.burn_in),
                       "l2": "persistent env state of one launch (%.0f MB) exceeds the 126 MB L2; fading tables "
                             "(36 MB) are L2-resident by design" % (env.state_bytes() / 1e6),
                       "kernel_variant": env.kernel_variant_name(), "live_ues_per_slice": m["n_live"] / (E * max(env.n_embb, 1)),
                       "cpu_affinity": affinity},
            "clocks": clocks,
            "e2e": {"value": max(e2e_pipe, e2e_sync), "unit": "env-steps/s", "h2d_bytes_per_step": 4 * S * E,
                    "d2h_bytes_per_step": (4 * V + 4 + 8 * S + 4) * E, "ms_per_step": 1e3 * min(t_pipe, t_sync) / K,
                    "mode": "pipelined host API (rs_step_async / rs_wait, 2 steps in flight)" if best_pipe
                            else "blocking host API (rs_step)",
                    "pipelined_value": e2e_pipe, "pipelined_ms_per_step": 1e3 * t_pipe / K,
                    "blocking_value": e2e_sync, "blocking_ms_per_step": 1e3 * t_sync / K,
                    "note": "blocking = every result on the host before the next action is chosen (any learning agent); "
                            "pipelined = results of step i land while step i+1 runs (feedback-free policy, or two half-batches)"},
            "gpu_launches": m["launches"],
            "roofline": roof,
        }
        if issue:
            line["issue_roofline"] = issue
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_cpu = 16 * threads
            val, dt = cpu_port_throughput(scn, n_cpu, args.burn_in, 20, 2, threads)
            line["cpu_baseline"] = {"value": val, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                    "sample": "%d envs x 20 steps after %d burn-in steps (same burn-in as the GPU arm), %d "
                                              "pthreads, oracle C port" % (n_cpu, args.burn_in, threads)}
            if not args.no_python_reference:
                line["cpu_baseline"]["python_reference"] = python_reference_throughput(scn)
    env.close()
    del env
    if rank == 0 and world == 1 and not args.no_configs and scn == 0 and args.envs_per_gpu == 65536:
        line["configs"] = side_configs(dev, local, K, warmup, args.burn_in, barrier, peaks)
    if rank == 0:
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def side_configs(dev, local, K, warmup, burn_in, barrier, peaks):
    """The other single-GPU configurations of BASELINE.json, each measured here with its own clock record:
    configs[1] scenario_0 x 4096 envs, configs[3] scenario_3 x 65536 envs, configs[2] scenario_0 x 16384 envs with the KBRL
    update / select kernels in the loop (tools/kbrl_loop.py, 300 steps here; the 2000-step protocol of SURVEY 8d is in profiles/)."""
    from ranslice_b200 import create_batched_env
    out = []
    for name, scn, E in (("configs[1]: scenario_0, 4096 envs", 0, 4096), ("configs[3]: scenario_3, 65536 envs", 3, 65536)):
        env = create_batched_env(BASE_SEED, scn, E, device=local)
        env.reset()
        m = measure(env, scn, K, warmup, burn_in, dev, 0, 1, local, barrier)
        roof, issue = roofline_blocks(env, scn, m, peaks, (m["clocks"] or {}).get("sm_mhz"))
        out.append({"config": name, "value": E * K / m["t_dev"], "unit": "env-steps/s", "ms_per_step": 1e3 * m["t_dev"] / K,
                    "e2e_blocking": E * K / m["t_sync"], "e2e_pipelined": E * K / m["t_pipe"], "gpu_launches": m["launches"],
                    "roofline": {k: roof[k] for k in ("achieved", "peak", "frac", "kernel_ms", "side_kernels_ms", "bytes_per_env_step") if k in roof}
                                | ({"mmtc_scan": roof["mmtc_scan"]} if "mmtc_scan" in roof else {}),
                    "clocks": m["clocks"]})
        env.close()
        del env
    sampler = ClockSampler(local)
    sampler.start()
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kbrl_loop.py"), "--envs", "16384", "--steps", "280",
                            "--warm", "20", "--report", "100,200,300", "--resident"], capture_output=True, text=True, timeout=600)
        kb = json.loads(r.stdout.strip().splitlines()[-1])
        keep = ("env_steps_per_s", "ms_per_step_wall", "ms_env", "ms_update_control", "ms_select_action", "dict_mean", "dict_p99",
                "dict_max", "cap_hits", "pool_hits", "pool_used_gb", "after_steps")
        entry = {"config": "configs[2]: scenario_0, 16384 envs, KBRL update/select kernels in the loop (device-resident controller)",
                 "value": kb["env_steps_per_s"], "unit": "env-steps/s"}
        entry.update({k: kb[k] for k in keep if k in kb})
    except Exception as e:                                    # noqa: BLE001
        entry = {"config": "configs[2]", "error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    entry["clocks"] = sampler.stop()
    out.append(entry)
    return out


if __name__ == "__main__":
    main()
