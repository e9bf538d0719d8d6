"""N>1 host-side path on CPU: two gloo ranks each step their own shard of envs (the oracle stands in for the
GPU here) keyed by global env id; the host gather must reproduce the single-process batch bit for bit, and
the timing reduction must be the max over ranks."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, steps, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import oracle_lib as ol
    from ranslice_b200.sharding import gather_on_host, max_over_ranks, shard_range
    from ranslice_b200.tables import load_tables
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    b = ol.OracleBatch(load_tables(), 3, hi - lo, 4242, first_env=lo)
    b.reset()
    rng = np.random.default_rng(7)
    obs_all = None
    for t in range(steps):
        w = rng.random((n_total, 3))
        a = np.floor(70 * w[:, :2] / w.sum(axis=1, keepdims=True)).astype(np.int64)
        obs, rew, lab, vio, _ = b.step(a[lo:hi])
        obs_all = gather_on_host(obs, n_total)
        rew_all = gather_on_host(rew, n_total)
    tmax = max_over_ranks(1.0 + rank)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), obs=obs_all, rew=rew_all, tmax=tmax)
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_batch(tmp_path, tables):
    import oracle_lib as ol
    from ranslice_b200.sharding import shard_range
    n_total, steps = 7, 12                     # odd size: ragged shards (4 + 3)
    assert [shard_range(n_total, r, 2) for r in range(2)] == [(0, 4), (4, 7)]
    assert shard_range(0, 0, 2) == (0, 0)
    mp.spawn(_worker, args=(2, _free_port(), n_total, steps, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npz")
    whole = ol.OracleBatch(tables, 3, n_total, 4242)
    whole.reset()
    rng = np.random.default_rng(7)
    for t in range(steps):
        w = rng.random((n_total, 3))
        a = np.floor(70 * w[:, :2] / w.sum(axis=1, keepdims=True)).astype(np.int64)
        obs, rew, _, _, _ = whole.step(a)
    assert np.array_equal(got["obs"], obs) and np.array_equal(got["rew"], rew)
    assert float(got["tmax"]) == 2.0
