"""Host-side mirror of the reference's ``scenario_creator.py`` config surface.

Same names and meaning as the reference (scenario list, traffic / SLA / mMTC descriptions,
``create_env`` signature, scenario_creator.py:26-183); the environment it returns is the native
single-env facade :class:`ranslice_b200.ran_slice.RanSlice` (or the batched env through
``create_batched_env``).  The numeric traffic / SLA constants are fixed inside the native
library exactly as below; they are exported here for callers that read them.
"""
import numpy as np

scenario_1 = {'n_prbs': 200, 'n_embb': 5, 'n_mmtc': 0}     # scenario_creator.py:26-30
scenario_2 = {'n_prbs': 150, 'n_embb': 3, 'n_mmtc': 2}     # :32-36
scenario_3 = {'n_prbs': 100, 'n_embb': 1, 'n_mmtc': 4}     # :38-42
scenario_4 = {'n_prbs': 70, 'n_embb': 1, 'n_mmtc': 1}      # :44-48
scenarios = [scenario_1, scenario_2, scenario_3, scenario_4]

CBR_description = {'lambda': 2.0 / 60.0, 't_mean': 30.0, 'bit_rate': 500000}                     # :55-60
VBR_description = {'lambda': 5.0 / 60.0, 't_mean': 30.0, 'p_size': 1000, 'b_size': 500, 'b_rate': 1}  # :62-69
SLA_embb = {'cbr_th': 10e6, 'cbr_prb': 20, 'cbr_queue': 10e4, 'vbr_th': 15e6, 'vbr_prb': 30, 'vbr_queue': 15e4}
state_variables_embb = ['cbr_traffic', 'cbr_th', 'cbr_prb', 'cbr_queue', 'cbr_snr',
                        'vbr_traffic', 'vbr_th', 'vbr_prb', 'vbr_queue', 'vbr_snr']               # :80-82
MTC_description = {'n_devices': 1000, 'repetition_set': [2, 4, 8, 16, 32, 64, 128],
                   'period_set': [1000, 50000, 10000, 15000, 20000, 25000, 50000, 100000]}        # :86-90
state_variables_mmtc = ['devices', 'avg_rep', 'delay']                                             # :92
SLA_mmtc = {'delay': 300}

PROPAGATION = {'macro_cell_urban_2GHz': (128.1, 37.6), 'macro_cell_urban_900MHz': (120.9, 37.6),
               'macro_cell_rural': (95.5, 34.1)}                                                   # channel_models.py:117-124


def seed_from_rng(rng):
    """The native env is keyed by a 64-bit Philox seed.  ``rng`` may be an int (used as is) or a
    ``numpy.random.Generator`` (one 63-bit draw), matching ``create_env(rng, ...)``'s first argument."""
    if isinstance(rng, (int, np.integer)):
        return int(rng)
    return int(rng.integers(0, 2 ** 63 - 1))


def create_batched_env(rng, n, n_envs, slots_per_step=50, propagation_type='macro_cell_urban_2GHz',
                       L1_level=True, penalty=100, device=0, first_env_id=0, **kw):
    from .batched import BatchedRanSlice
    # L1_level=False: the eMBB RAN slices share one L1 scheduler (scenario_creator.py:168-177); correctness-first kernel
    return BatchedRanSlice(scenario=n, n_envs=n_envs, base_seed=seed_from_rng(rng), slots_per_step=slots_per_step,
                           propagation_type=propagation_type, penalty=penalty, device=device,
                           first_env_id=first_env_id, l1_level=bool(L1_level), **kw)


def create_env(rng, n, slots_per_step=50, propagation_type='macro_cell_urban_2GHz', L1_level=True, penalty=100,
               device=0):
    """Drop-in for ``scenario_creator.create_env`` (scenario_creator.py:100-183): returns a gym-style
    single env with ``reset()/step(action)`` and attrs ``n_prbs, n_slices, n_variables,
    action_space, observation_space``."""
    from .ran_slice import RanSlice
    batched = create_batched_env(rng, n, 1, slots_per_step, propagation_type, L1_level, penalty, device)
    return RanSlice(batched)
