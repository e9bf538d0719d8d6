"""ranslice-b200: batched RAN-slicing env step path for B200.

Public surface (mirrors the reference's):
    create_env(rng, n, ...)            scenario_creator.create_env drop-in (single env, gym API)
    create_batched_env(rng, n, N, ...) N envs in lockstep on one GPU
    BatchedRanSlice, RanSlice
"""
from .scenario_creator import create_batched_env, create_env, scenarios  # noqa: F401


def __getattr__(name):
    if name == "BatchedRanSlice":
        from .batched import BatchedRanSlice
        return BatchedRanSlice
    if name == "RanSlice":
        from .ran_slice import RanSlice
        return RanSlice
    raise AttributeError(name)
