import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from xritdemod_b200 import demod as xd, siggen
N = 1 << 20
p = siggen.params("hrit", 0, n=N, ramp_len=N); x = siggen.generate(p, N)
c = xd.CostasLoop()
cuts = [0, 1, 8, 15, 1000, 65535 + 1000, 300000, 300007, 1 << 20]
for a, b in zip(cuts[:-1], cuts[1:]):
    print("call", a, b, flush=True)
    c.Work(x[a:b])
print("ok")
