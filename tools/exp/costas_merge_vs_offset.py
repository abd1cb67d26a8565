import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_ffi as o
from xritdemod_b200 import siggen
N = 1 << 22
for chn in (7, 0):
    p = siggen.params("hrit", chn, n=125000000, ramp_len=1 << 20); x = siggen.generate(p, N)
    ch = o.Chain(o.config(True)); sym, taps = ch.process(x, taps=True)
    r = taps["rrc"]
    B = 512
    # true states at block boundaries
    c = o.Costas(0.0037, 2)
    st = [c.state]
    for b in range(N // B):
        c.work(r[b * B:(b + 1) * B]); st.append(c.state)
    res = []
    for start in range(1 << 20, N - (1 << 19), 150000):
        s0 = start // B
        tp, tf = st[s0]
        for kind, ph0, f0 in (("same rep, freq 0", tp + 0.05, 0.0), ("other rep", tp + 0.05 + (2*np.pi if tp < 0 else -2*np.pi), 0.0), ("same rep, true freq", tp + 0.05, tf)):
            cb = o.Costas(0.0037, 2); cb.state = (float(np.float32(ph0)), float(np.float32(f0)))
            merged = None
            for b in range(s0, min(s0 + 400000 // B, N // B)):
                cb.work(r[b * B:(b + 1) * B])
                if cb.state == st[b + 1]:
                    merged = (b + 1 - s0) * B; break
            res.append((kind, merged))
    for kind in ("same rep, freq 0", "other rep", "same rep, true freq"):
        v = [m for k, m in res if k == kind]
        print("channel", chn, kind, "merge samples:", v)
