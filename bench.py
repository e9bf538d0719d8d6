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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = 16 * threads
    burn = min(args.burn_in, 200)
    val, dt = cpu_port_throughput(args.scenario, n_envs, burn, args.steps, args.warmup, threads)
    S, n_prbs, V = SCN[args.scenario]
    sample = "%d envs x %d steps after %d burn-in steps, %d pthreads" % (n_envs, args.steps, burn, threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "scenario_%d, bounded CPU sample: %s" % (args.scenario, sample),
                       "note": "the reference is pure Python and cannot travel to the GPU box; this arm is the C "
                               "port of its algorithm (oracle/), pinned bit-exact to the reference; the Python "
                               "reference itself measured 8.5 env-steps/s/core in the build container (BASELINE.md)"},
            "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# --------------------------------------------------------------------------------------------- GPU arm
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
    args = ap.parse_args()
    _divert_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    # ---- actions: random-simplex policy, generated on the host (pinned) once, copied to HBM for the device arm
    rng = np.random.default_rng(1000 + rank)
    n_act = warmup + K
    host_act = torch.empty((n_act, E, S), dtype=torch.int32, pin_memory=True)
    for i in range(n_act):
        host_act[i] = torch.from_numpy(simplex_actions(rng, E, S, n_prbs))
    dev_act = host_act.to(dev)
    out = env.step_device(dev_act[0])                       # allocates the output tensors (part of burn-in)
    for i in range(args.burn_in):                           # population burn-in (untimed set-up)
        env.step_device(dev_act[i % n_act], out)
    torch.cuda.synchronize()

    # ---- device-resident arm (inputs already in HBM)
    for i in range(warmup):
        env.step_device(dev_act[i], out)
    k0, _ = env.counters()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        env.step_device(dev_act[warmup + i], out)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    k1, trace_elems = env.counters()
    launches = k1 - k0

    # ---- end-to-end arm: public host API, pinned host buffers, H2D + kernels + D2H of EVERY step inside the timed region.
    # (a) blocking call (step_host_inplace = rs_step); (b) the pipelined call (step_host_async / wait = rs_step_async /
    # rs_wait, two steps in flight: the D2H of step i overlaps the kernels of step i+1).  The headline e2e is (b): a
    # random policy has no feedback from step i to step i+1, and a learning agent gets the same overlap by alternating two
    # half-batches.
    for i in range(warmup):
        env.step_host_inplace(host_act[i].numpy())
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        env.step_host_inplace(host_act[warmup + i].numpy())
    torch.cuda.synchronize()
    t_e2e_sync = time.perf_counter() - t0
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
    t_e2e = time.perf_counter() - t0
    barrier()

    # ---- dominant-kernel duration (CUDA events on the launching stream, inside the library)
    prof = env.profile_steps([dev_act[warmup + (i % K)] for i in range(min(K, 10))], out)
    n_live = int(env.n_ues().sum())

    t_dev, t_e2e, t_e2e_sync = max_over_ranks(t_dev), max_over_ranks(t_e2e), max_over_ranks(t_e2e_sync)   # slowest shard decides
    total_envs = E * world

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        # algorithmic bytes of one launch: 2*B_state + B_io + B_trace, B_trace counted by the kernel
        b_state = S * 16 * E + 88 * n_live
        b_io = (4 * S + 4 * V + 4 + 8 * S) * E
        b_trace = 4 * trace_elems
        b_mtc = 4200 * env.n_mmtc * E                         # one scan of the 1000 next-arrival words + backlog
        b_alg = 2 * b_state + b_io + b_trace + b_mtc
        traffic = None                                        # dram bytes per launch from the committed ncu capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tr.get("envs_per_gpu") == E and tr.get("scenario") == scn:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        k_ms = prof["embb_ms"] if scn != 3 else prof["embb_ms"] + prof["mmtc_ms"]
        achieved = b_alg / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": total_envs * K / t_dev, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": warmup, "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "scenario_%d, %d envs per GPU (%d total), %d TTIs/step, random-simplex actions, "
                                   "%d-step population burn-in" % (scn, E, total_envs, env.slots_per_step, args.burn_in),
                       "l2": "persistent env state of one launch (%.0f MB) exceeds the 126 MB L2; fading tables "
                             "(36 MB) are L2-resident by design" % (env.state_bytes() / 1e6),
                       "kernel_variant": env.kernel_variant_name(), "live_ues_per_slice": n_live / (E * max(env.n_embb, 1))},
            "clocks": clocks,
            "e2e": {"value": total_envs * K / t_e2e, "unit": "env-steps/s", "h2d_bytes_per_step": 4 * S * E,
                    "d2h_bytes_per_step": (4 * V + 4 + 8 * S + 4) * E, "ms_per_step": 1e3 * t_e2e / K,
                    "mode": "pipelined host API (rs_step_async / rs_wait, 2 steps in flight)",
                    "blocking_value": total_envs * K / t_e2e_sync, "blocking_ms_per_step": 1e3 * t_e2e_sync / K},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": prof["kernel"],
                         "kernel_ms": k_ms, "bytes_per_launch": b_alg,
                         "bytes_per_env_step": b_alg / E, "nominal_bytes_per_env_step": NOMINAL_BALG.get(scn),
                         "note": "issue/latency-bound path (serial PF loop, fp64 decisions); tables are L2-resident, "
                                 "so DRAM traffic is far below the algorithmic bytes (SURVEY 8d caveat)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_cpu = 16 * threads
            burn = 200
            val, dt = cpu_port_throughput(scn, n_cpu, burn, 20, 2, threads)
            line["cpu_baseline"] = {"value": val, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                    "sample": "%d envs x 20 steps after %d burn-in steps, %d pthreads (oracle C port; "
                                              "Python reference: 8.5 env-steps/s/core, BASELINE.md)" % (n_cpu, burn, threads)}
        _emit(line)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
