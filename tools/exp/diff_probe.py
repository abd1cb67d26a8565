import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from xritdemod_b200 import demod as xd, siggen
import oracle_ffi
N = 300000
x = siggen.generate(siggen.params("hrit", 7, n=N, ramp_len=N), N)
orc = oracle_ffi.Chain(oracle_ffi.config(True)).process(x)
small = dict(costas_seg=16384, costas_warm=2048, agc_seg=8192, agc_warm=1024, mm_seg=60000, mm_warm=30000)
def first_diff(a, b):
    n = min(len(a), len(b))
    ne = np.nonzero(a[:n].view(np.uint64) != b[:n].view(np.uint64))[0]
    return (len(a), len(b), int(ne[0]) if len(ne) else -1, len(ne))
for kw in (dict(), dict(guided=2), dict(rerun_kernel=3), dict(rerun_kernel=7), dict(rerun_kernel=4, chase=2), dict(loop_kernel=4)):
    for rep in range(2):
        d = xd.Demodulator(mode="hrit")
        t = dict(small); t.update(kw)
        d.set_tuning(**t)
        a = d.demod(x[:170001])
        blob = d.checkpoint()
        b = d.demod(x[170001:])
        d.restore(blob)
        b2 = d.demod(x[170001:])
        sym = np.concatenate([a, b])
        print(kw, rep, "vs oracle", first_diff(sym, orc), "b vs b2", first_diff(b, b2), flush=True)
