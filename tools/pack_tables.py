#!/usr/bin/env python3
"""Pack the reference's lookup tables (datasets/*.csv) into network-slicing_b200/data/tables.npz.

Build-container tool (needs /root/reference/datasets).  The fading traces are printed with
<= 6 significant digits, so every value is stored losslessly as (int32 mantissa, int8 number
of decimals): value = mantissa / 10**decimals, which is a correctly rounded fp64 division
(both operands exact).  The script checks the reconstruction against the reference's own
parse (``pd.read_csv(filename, header=None)``, channel_models.py:143) and reports mismatches.
The trailing empty field of every row (NaN column 10000) is not stored; the
loader re-creates it.
"""
import os
import re
import sys
from decimal import Decimal

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DS = os.path.join(os.environ.get("RANSLICE_REFERENCE", "/root/reference"), "datasets")
NAMES = ["fading_trace_EPA_3kmph.csv", "fading_trace_ETU_3kmph.csv", "fading_trace_EVA_60kmph.csv"]


def pack_one(fn):
    rows = open(fn).read().strip("\n").split("\n")
    mant = np.zeros((len(rows), 10000), np.int32)
    dec = np.zeros((len(rows), 10000), np.int8)
    for r, line in enumerate(rows):
        toks = line.split(",")
        assert len(toks) == 10001 and toks[-1].strip() == "", (len(toks), toks[-1])
        for c, t in enumerate(toks[:-1]):
            d = Decimal(t.strip())
            sign, digits, exp = d.as_tuple()
            m = int("".join(map(str, digits))) * (-1 if sign else 1)
            if exp > 0:
                m *= 10 ** exp
                exp = 0
            assert abs(m) < 2 ** 31 and -exp < 23
            mant[r, c] = m
            dec[r, c] = -exp
    return mant, dec


def main():
    mants, decs = [], []
    for n in NAMES:
        fn = os.path.join(DS, n)
        mant, dec = pack_one(fn)
        ref = pd.read_csv(fn, header=None).to_numpy()
        assert ref.shape == (100, 10001) and np.isnan(ref[:, 10000]).all()
        rec = mant.astype(np.float64) / np.power(10.0, dec.astype(np.float64))
        diff = rec != ref[:, :10000]
        print(n, "mismatches vs pandas parse:", int(diff.sum()),
              "max rel", float(np.max(np.abs(rec - ref[:, :10000]) / np.maximum(np.abs(rec), 1e-300))))
        mants.append(mant)
        decs.append(dec)
    mcs = pd.read_csv(os.path.join(DS, "mcs_codeset.csv"))
    out = os.path.join(ROOT, "network-slicing_b200", "data", "tables.npz")
    np.savez_compressed(
        out, trace_mant=np.stack(mants), trace_dec=np.stack(decs), trace_names=np.array(NAMES),
        mcs_rate=mcs["rate"].to_numpy(), mcs_snr=mcs["snr"].to_numpy(),
        mcs_order=mcs["order"].to_numpy().astype(np.int32),
        mcs_modulation=np.array([{"qpsk": 0, "16qam": 1, "64qam": 2}[m] for m in mcs["modulation"]], np.int32))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
