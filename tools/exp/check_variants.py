"""Chain bit-exactness for every AGC/Costas kernel variant (GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_ffi as o
from xritdemod_b200 import demod as xd, siggen
N = 1 << 22
for mode in ("hrit", "lrit"):
    p = siggen.params(mode, 0, n=N, ramp_len=1 << 20); x = siggen.generate(p, N)
    ref = o.Chain(o.config(mode == "hrit")).process(x)
    for k in (2, 3, 4, 5, 6):
        d = xd.Demodulator(mode=mode); d.set_tuning(loop_kernel=k)
        s = d.demod(x)
        ok = len(s) == len(ref) and np.array_equal(s.view(np.float32), ref.view(np.float32))
        d2 = xd.Demodulator(mode=mode); d2.set_tuning(loop_kernel=k, costas_seg=4096, costas_warm=512, agc_seg=2048, agc_warm=64)
        parts = np.concatenate([d2.demod(x[i:i + 300001]) for i in range(0, N, 300001)])
        ok2 = len(parts) == len(ref) and np.array_equal(parts.view(np.float32), ref.view(np.float32))
        print(mode, "variant", k, "one-shot", "BIT-EXACT" if ok else "MISMATCH", "| tiny segments, chunked", "BIT-EXACT" if ok2 else "MISMATCH", flush=True)
