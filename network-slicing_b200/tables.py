"""Lookup tables of the step path, repacked for the device.

``data/tables.npz`` is produced by ``tools/pack_tables.py`` from the reference's
``datasets/fading_trace_*.csv`` and ``datasets/mcs_codeset.csv`` (values identical to the
reference's own ``pd.read_csv`` parse, checked by that tool).  Layout handed to the native
library: fading traces **time-major** ``[3][10001][100]`` fp64, so that the PRB window a UE
reads in one TTI is contiguous (the reference is PRB-major, ``channel_models.py:188``);
time column 10000 is NaN exactly like the reference's trailing empty CSV field.
"""
import os

import numpy as np

N_SAMPLES = 10001
TRACE_ROWS = 100
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tables.npz")
_cache = {}


class Tables:
    def __init__(self, trace, mcs_rate, mcs_snr, mcs_order, mcs_mod):
        self.trace = trace            # float64 [3, 10001, 100]
        self.mcs_rate = mcs_rate      # float64 [26]
        self.mcs_snr = mcs_snr        # float64 [26]
        self.mcs_order = mcs_order    # int32 [26]
        self.mcs_mod = mcs_mod        # int32 [26]  0 qpsk / 1 16qam / 2 64qam


def load_tables(path=None):
    path = path or _DATA
    if path in _cache:
        return _cache[path]
    z = np.load(path)
    vals = z["trace_mant"].astype(np.float64) / np.power(10.0, z["trace_dec"].astype(np.float64))
    trace = np.full((3, N_SAMPLES, TRACE_ROWS), np.nan, np.float64)
    trace[:, :N_SAMPLES - 1, :] = vals.transpose(0, 2, 1)
    t = Tables(np.ascontiguousarray(trace), np.ascontiguousarray(z["mcs_rate"], np.float64),
               np.ascontiguousarray(z["mcs_snr"], np.float64),
               np.ascontiguousarray(z["mcs_order"], np.int32),
               np.ascontiguousarray(z["mcs_modulation"], np.int32))
    _cache[path] = t
    return t
