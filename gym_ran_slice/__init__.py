"""Drop-in for the reference's ``gym_ran_slice`` package (gym-ran_slice/gym_ran_slice/__init__.py:1-8): registers the id
``RanSlice-v1`` and accepts the reference's constructor kwargs ``node_b=`` / ``penalty=`` (scenario_creator.py:181:
``gym.make('gym_ran_slice:RanSlice-v1', node_b=node, penalty=penalty)``).

The reference's ``NodeB`` is a Python object graph; the native env is configured by a scenario description instead.  So
``node_b`` may be
  * a :class:`NodeBSpec` (or a dict with its fields): scenario index / dict, seed, slots_per_step, propagation_type, L1_level;
  * an object with the attributes the reference's NodeB exposes (``n_prbs``, ``slots_per_step``, ``slices_l1`` with ``.type``
    and ``.slices_ran``): the scenario is read off it (slice counts, PRBs, slots per step); its Python state is not used.
Registration happens on import when ``gym`` or ``gymnasium`` is importable; ``make()`` below works without either.
"""
from dataclasses import dataclass

ENV_ID = "RanSlice-v1"


@dataclass
class NodeBSpec:
    scenario: object = 0                      # index into scenario_creator.scenarios or a {'n_prbs', 'n_embb', 'n_mmtc'} dict
    seed: int = 0
    slots_per_step: int = 50
    propagation_type: str = "macro_cell_urban_2GHz"
    L1_level: bool = True
    device: int = 0


def _spec_from_node_b(node_b):
    if isinstance(node_b, NodeBSpec):
        return node_b
    if isinstance(node_b, dict):
        return NodeBSpec(**node_b)
    slices = getattr(node_b, "slices_l1", None)
    if slices is None:
        raise TypeError("node_b must be a NodeBSpec, a dict of its fields, or an object with NodeB's attributes")
    n_embb = sum(len(l1.slices_ran) for l1 in slices if getattr(l1, "type", "") == "eMBB")
    n_mmtc = sum(len(l1.slices_ran) for l1 in slices if getattr(l1, "type", "") != "eMBB")
    l1_level = all(len(l1.slices_ran) == 1 for l1 in slices)
    return NodeBSpec(scenario={"n_prbs": int(node_b.n_prbs), "n_embb": n_embb, "n_mmtc": n_mmtc},
                     slots_per_step=int(getattr(node_b, "slots_per_step", 50)), L1_level=l1_level,
                     seed=int(getattr(node_b, "seed", 0)))


def RanSlice(node_b=None, penalty=100, **kw):
    """Entry point of the registered id: ``RanSlice(node_b=..., penalty=...)`` (ran_slice.py:19) -> native single-env facade."""
    from ranslice_b200.batched import BatchedRanSlice
    from ranslice_b200.ran_slice import RanSlice as _Facade
    spec = _spec_from_node_b(node_b if node_b is not None else NodeBSpec(**kw))
    batched = BatchedRanSlice(scenario=spec.scenario, n_envs=1, base_seed=spec.seed, slots_per_step=spec.slots_per_step,
                              propagation_type=spec.propagation_type, penalty=penalty, device=spec.device,
                              l1_level=spec.L1_level)
    return _Facade(batched)


def make(id=ENV_ID, **kw):
    """``gym.make`` for boxes without gym: accepts 'RanSlice-v1' and 'gym_ran_slice:RanSlice-v1'."""
    if id.split(":")[-1] != ENV_ID:
        raise ValueError("unknown id %r" % id)
    return RanSlice(**kw)


def _register():
    done = []
    for modname in ("gym", "gymnasium"):
        try:
            mod = __import__(modname + ".envs.registration", fromlist=["register"])
            mod.register(id=ENV_ID, entry_point="gym_ran_slice:RanSlice")
            done.append(modname)
        except Exception:                      # not installed, or already registered
            pass
    return done


REGISTERED_WITH = _register()
