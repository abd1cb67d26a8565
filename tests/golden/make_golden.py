"""Generates tests/golden/chain_golden.npz: regression pins of the CPU oracle on the seeded
synthetic HRIT/LRIT bursts.  The reference ships no golden vectors, so these pin
the oracle against itself across edits; the independent pins are in tests/test_oracle.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_ffi as o  # noqa: E402
from xritdemod_b200 import siggen  # noqa: E402

out = {}
for mode in ("hrit", "lrit"):
    n = 1 << 18
    p = siggen.params(mode, 0, n=n, ramp_len=n)
    x = siggen.generate(p, n)
    sym = o.Chain(o.config(mode == "hrit")).process(x)
    out[mode + "_n"] = n
    out[mode + "_iq_head"] = x[:64].copy()
    out[mode + "_nsym"] = len(sym)
    out[mode + "_sym_head"] = sym[:256].copy()
    out[mode + "_sym_tail"] = sym[-256:].copy()
    out[mode + "_resum"] = np.float64(sym.real.astype(np.float64).sum())
np.savez(os.path.join(HERE, "chain_golden.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items()})
