"""How fast do two M&M trajectories merge bitwise (oracle, CPU)?"""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_ffi as o
from xritdemod_b200 import siggen
N = 1 << 22
p = siggen.params("hrit", 0, n=N, ramp_len=1 << 20); x = siggen.generate(p, N)
ch = o.Chain(o.config(True)); sym, taps = ch.process(x, taps=True)
c = taps["costas"]; sps = ch.sps
gm = np.float32(0.0037); go = np.float32(gm * gm / np.float32(4))
B = 4096
def run(m, s0, nblk, rec):
    for b in range(nblk):
        m.work(c[s0 + b * B: s0 + (b + 1) * B]); st = m.state
        rec.append((st.mu, st.omega, st.next_index, st.p0[0], st.p1[0]))
def key(st): return (st.mu, st.omega, st.next_index, st.p0[0], st.p1[0])
m = o.Mm(sps, go, 0.5, gm, 0.005)
start = 1 << 20
m.work(c[:start]); st0 = m.state
for kind in ("omega+1ulp", "mu+1e-3", "omega+40ulp", "cold"):
    ma = o.Mm(sps, go, 0.5, gm, 0.005); mb = o.Mm(sps, go, 0.5, gm, 0.005)
    ma.work(c[:start]); mb.work(c[:start])
    sb = mb.state
    if kind == "omega+1ulp": sb.omega = float(np.nextafter(np.float32(sb.omega), np.float32(10)))
    elif kind == "omega+40ulp": sb.omega = float(np.float32(sb.omega) + 40 * np.spacing(np.float32(sb.omega)))
    elif kind == "mu+1e-3": sb.mu = float(np.float32(sb.mu + 1e-3))
    else:
        sb.mu = 0.5; sb.omega = float(sps)
    mb.state = sb
    ra, rb = [], []
    nblk = 700
    run(ma, start, nblk, ra); run(mb, start, nblk, rb)
    eq = [a == b for a, b in zip(ra, rb)]
    first = eq.index(True) if True in eq else -1
    print(kind, "first merged block (x%d samples):" % B, first, "stays merged:", all(eq[first:]) if first >= 0 else None,
          "| d_omega ulps at blocks 10,100,300,699:", [round((rb[i][1] - ra[i][1]) / 2.384e-7) for i in (10, 100, 300, 699)],
          "d_mu:", ["%.1e" % (rb[i][0] - ra[i][0]) for i in (10, 100, 300, 699)])
print("sps", sps, "st0", st0.mu, st0.omega, st0.next_index)
for i in (0, 10, 100, 300, 699): print(i, ra[i], rb[i])
