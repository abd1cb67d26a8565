import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from xritdemod_b200 import demod as xd, siggen
import oracle_ffi
N = 300000
x = siggen.generate(siggen.params("hrit", 7, n=N, ramp_len=N), N)
_, taps = oracle_ffi.Chain(oracle_ffi.config(True)).process(x, taps=True)
rrc, cos = taps["rrc"], taps["costas"]
def first_diff(a, b):
    ne = np.nonzero(a.view(np.uint64) != b.view(np.uint64))[0]
    return (int(ne[0]) if len(ne) else -1, len(ne))
bad = {}
for kernel in (4, 6, 7, 3, 5):
    out = []
    for cut in list(range(169990, 170012)) + [163841, 163843, 165001, 168001, 172001, 180001, 150001, 140001]:
        c = xd.CostasLoop()
        c.set_loop_kernel(kernel)
        c.set_tuning(16384, 2048)
        y = np.concatenate([c.Work(rrc[:cut]), c.Work(rrc[cut:])])
        fd = first_diff(y, cos)
        if fd[0] >= 0:
            out.append((cut, cut % 16384, fd))
    print("kernel", kernel, "failing cuts (cut, last len, first diff/count):", out, flush=True)
