import os, sys, time
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_ffi as o
from xritdemod_b200 import demod
rng = np.random.default_rng(7)
for name, n in (("noise only", 16 << 20), ("zeros then noise", 16 << 20)):
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) * np.float32(0.1)
    if name.startswith("zeros"):
        x[: n // 2] = 0
    ref = o.Chain(o.config(True)).process(x)
    d = demod.Demodulator(mode="hrit")
    t = time.time(); y = d.demod(x); dt = time.time() - t
    st = d.stats()
    same = len(y) == len(ref) and np.array_equal(y.view(np.uint32), ref.view(np.uint32))
    print(name, "n", n, "symbols", len(y), len(ref), "bit-exact", same, "%.1f ms" % (dt * 1e3),
          {k: st[k] for k in ("agc_rounds", "costas_rounds", "mm_rounds", "mm_redo", "mm_bail", "costas_redo")}, flush=True)
