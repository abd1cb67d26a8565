"""Experiment: how fast do two Costas trajectories merge bitwise (oracle, CPU)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_ffi as o
from xritdemod_b200 import siggen
N = 1 << 22
p = siggen.params("hrit", 0, n=N, ramp_len=1 << 20); x = siggen.generate(p, N)
ch = o.Chain(o.config(True)); sym, taps = ch.process(x, taps=True)
r = taps["rrc"]
# true trajectory states at block boundaries
B = 256
c = o.Costas()
states = []
for i in range(0, N, B):
    states.append(c.state); c.work(r[i:i+B])
states = np.array(states, np.float32)
def run_from(s0, st, nblk):
    c = o.Costas(); c.state = st
    out = []
    for b in range(nblk):
        c.work(r[s0 + b*B: s0 + (b+1)*B]); out.append(c.state)
    return np.array(out, np.float32)
rng = np.random.default_rng(1)
for kind in ("cold", "ulp_freq", "ulp_phase", "ulp10_freq"):
    res = []
    for trial in range(40):
        b0 = int(rng.integers(4096, N // B - 400))
        s0 = b0 * B
        ph, fr = states[b0]
        if kind == "cold": st = (0.0, 0.0)
        elif kind == "ulp_freq": st = (ph, np.nextafter(np.float32(fr), np.float32(1)))
        elif kind == "ulp10_freq": st = (ph, np.float32(fr) + 10*np.spacing(np.float32(fr)))
        else: st = (np.nextafter(np.float32(ph), np.float32(10)), fr)
        nb = 384
        tr = run_from(s0, (float(st[0]), float(st[1])), nb)
        true = states[b0+1:b0+1+nb]
        # allow pi-rotated merge: compare freq exact and phase exact
        eq = (tr[:,0] == true[:,0]) & (tr[:,1] == true[:,1])
        eqf = (tr[:,1] == true[:,1])
        first = int(np.argmax(eq)) if eq.any() else -1
        firstf = int(np.argmax(eqf)) if eqf.any() else -1
        res.append((first, firstf))
    a = np.array(res)
    m = a[:,0]
    print(kind, "merged(exact):", (m>=0).sum(), "/", len(m), "median blocks(256):", np.median(m[m>=0]) if (m>=0).any() else None,
          "max", m.max(), "| freq-equal first:", np.median(a[:,1][a[:,1]>=0]) if (a[:,1]>=0).any() else None, (a[:,1]>=0).sum())
