"""Emulate the windowed Newton iteration for Costas: lanes hold believed (phase,freq); literal step;
exact prefix sum of exact differences (float64 here as a stand-in for int64 fixed point);
accept the leading run with s[i+1] == out[i]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_ffi as o
from xritdemod_b200 import siggen
f32 = np.float32
N = 1 << 19
mode = sys.argv[2] if len(sys.argv) > 2 else "hrit"
p = siggen.params(mode, 0, n=N, ramp_len=1 << 20); x = siggen.generate(p, N)
ch = o.Chain(o.config(mode == "hrit")); sym, taps = ch.process(x, taps=True)
r = taps["rrc"]
alpha, beta = o.costas_gains(0.0037); alpha = f32(alpha); beta = f32(beta)
TWO_PI_F = f32(6.28318500518798828125)
def step(ph, fr, xr, xi):
    cs = np.cos(-ph).astype(f32); sn = np.sin(-ph).astype(f32)
    yr = (xr * cs - xi * sn).astype(f32); yi = (xr * sn + xi * cs).astype(f32)
    e = (yr * yi).astype(f32)
    e = (f32(0.5) * (np.abs(e + f32(1)) - np.abs(e - f32(1)))).astype(f32)
    fr2 = (fr + beta * e).astype(f32)
    ph2 = ((ph + fr2).astype(f32) + (alpha * e).astype(f32)).astype(f32)
    hi = ph2 > TWO_PI_F; lo = ph2 < -TWO_PI_F
    ph2 = np.where(hi, (ph2.astype(np.float64) - 2*np.pi).astype(f32), ph2)
    ph2 = np.where(lo, (ph2.astype(np.float64) + 2*np.pi).astype(f32), ph2)
    fr2 = np.clip(fr2, f32(-1), f32(1))
    return ph2, fr2
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
xr = r.real.copy(); xi = r.imag.copy()
start = 200000
# exact base state: run the same literal step sequentially from a cold state for a while
ph = f32(0); fr = f32(0)
for i in range(start - 60000, start):
    a, b = step(np.array([ph]), np.array([fr]), xr[i:i+1], xi[i:i+1]); ph, fr = a[0], b[0]
base = start; bph, bfr = ph, fr
# believed states: linear extrapolation
idx = np.arange(NT)
sph = (np.float64(bph) + idx * np.float64(bfr)); sph = np.mod(sph + 2*np.pi, 4*np.pi) - 2*np.pi
sph = sph.astype(f32); sfr = np.full(NT, bfr, f32)
pos = base + idx            # sample index of each slot (window = [base, base+NT))
adv = []
iters = 0
total = 0
while total < 60000:
    iters += 1
    oph, ofr = step(sph, sfr, xr[base:base+NT], xi[base:base+NT])
    # acceptance: lanes 1..A-1 satisfy s[i] == out[i-1]
    ok = (sph[1:] == oph[:-1]) & (sfr[1:] == ofr[:-1])
    A = NT if ok.all() else int(np.argmin(ok)) + 1    # lanes 0..A-1 exact; new base = lane A with state out[A-1]
    # proposals from exact prefix sums (float64 stand-in)
    dph = oph.astype(np.float64) - sph.astype(np.float64)
    dfr = ofr.astype(np.float64) - sfr.astype(np.float64)
    P = np.float64(sph[0]) + np.cumsum(dph); Fq = np.float64(sfr[0]) + np.cumsum(dfr)   # proposal for lanes 1..NT
    nph = np.empty(NT + 1, f32); nfr = np.empty(NT + 1, f32)
    nph[0] = sph[0]; nfr[0] = sfr[0]; nph[1:] = P.astype(f32); nfr[1:] = Fq.astype(f32)
    # slide by A: new window = lanes A..A+NT-1; lanes beyond NT are extrapolated from the end state
    eph, efr = nph[NT], nfr[NT]
    ext = np.float64(eph) + np.arange(1, A + 1) * np.float64(efr); ext = np.mod(ext + 2*np.pi, 4*np.pi) - 2*np.pi
    sph = np.concatenate([nph[A:NT+1], ext.astype(f32)])[:NT]
    sfr = np.concatenate([nfr[A:NT+1], np.full(A, efr, f32)])[:NT]
    # the base lane must be literally exact: out[A-1]
    sph[0] = oph[A-1]; sfr[0] = ofr[A-1]
    base += A; total += A; adv.append(A)
adv = np.array(adv)
print("NT", NT, "iters", iters, "samples", total, "mean advance/iter %.1f" % adv.mean(), "min", adv.min(), "median", np.median(adv), "p10", np.percentile(adv, 10))
