"""Windowed Newton iteration for Costas, v2: int64 fixed point (2^-60), literal wrap-rule
extrapolation and proposal normalisation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_ffi as o
from xritdemod_b200 import siggen
f32 = np.float32
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = sys.argv[2] if len(sys.argv) > 2 else "hrit"
TOTAL = int(sys.argv[3]) if len(sys.argv) > 3 else 60000
N = 1 << 19
p = siggen.params(mode, 0, n=N, ramp_len=1 << 20); x = siggen.generate(p, N)
ch = o.Chain(o.config(mode == "hrit")); sym, taps = ch.process(x, taps=True)
r = taps["rrc"]
alpha, beta = o.costas_gains(0.0037); alpha = f32(alpha); beta = f32(beta)
TWO_PI_F = f32(6.28318500518798828125); TWO_PI = 2 * np.pi
S = 2.0 ** 60
def step(ph, fr, xr, xi):
    cs = np.cos(-ph).astype(f32); sn = np.sin(-ph).astype(f32)
    yr = (xr * cs - xi * sn).astype(f32); yi = (xr * sn + xi * cs).astype(f32)
    e = (yr * yi).astype(f32)
    e = (f32(0.5) * (np.abs(e + f32(1)) - np.abs(e - f32(1)))).astype(f32)
    fr2 = (fr + beta * e).astype(f32)
    ph2 = ((ph + fr2).astype(f32) + (alpha * e).astype(f32)).astype(f32)
    hi = ph2 > TWO_PI_F; lo = ph2 < -TWO_PI_F
    ph2 = np.where(hi, (ph2.astype(np.float64) - TWO_PI).astype(f32), ph2)
    ph2 = np.where(lo, (ph2.astype(np.float64) + TWO_PI).astype(f32), ph2)
    fr2 = np.clip(fr2, f32(-1), f32(1))
    return ph2, fr2
def fix(a): return np.trunc(np.asarray(a).astype(np.float64) * S).astype(np.int64)
def unfix(i): return (np.asarray(i).astype(np.float64) / S).astype(f32)
TWO_PI_FIX = np.int64(int(TWO_PI * S))
def norm(P):
    P = np.where(P > TWO_PI_FIX, P - TWO_PI_FIX, P)
    P = np.where(P < -TWO_PI_FIX, P + TWO_PI_FIX, P)
    return P
xr = r.real.copy(); xi = r.imag.copy()
start = 100000
ph = f32(0); fr = f32(0)
for i in range(start - 60000, start):
    a, b = step(np.array([ph]), np.array([fr]), xr[i:i+1], xi[i:i+1]); ph, fr = a[0], b[0]
base = start
def extrap(eph, efr, k0, k1):
    k = np.arange(k0, k1, dtype=np.int64)
    P = fix([eph])[0] + k * fix([efr])[0]
    q = np.where(P > TWO_PI_FIX, (P // TWO_PI_FIX), 0) + np.where(P < -TWO_PI_FIX, -((-P) // TWO_PI_FIX), 0)
    P = P - q * TWO_PI_FIX
    return unfix(P), np.full(len(k), efr, f32)
sph, sfr = extrap(ph, fr, 0, NT); sph[0] = ph; sfr[0] = fr
adv = []; total = 0; iters = 0
while total < TOTAL:
    iters += 1
    oph, ofr = step(sph, sfr, xr[base:base+NT], xi[base:base+NT])
    ok = (sph[1:] == oph[:-1]) & (sfr[1:] == ofr[:-1])
    A = NT if ok.all() else int(np.argmin(ok)) + 1
    Dp = fix(oph) - fix(sph); Df = fix(ofr) - fix(sfr)
    P = norm(fix(sph[:1])[0] + np.cumsum(Dp)); Fq = fix(sfr[:1])[0] + np.cumsum(Df)
    nph = np.concatenate([sph[:1], unfix(P)]); nfr = np.concatenate([sfr[:1], unfix(Fq)])
    eph, efr = nph[NT], nfr[NT]
    xph, xfr = extrap(eph, efr, 1, A + 1)
    sph = np.concatenate([nph[A:NT+1], xph])[:NT]; sfr = np.concatenate([nfr[A:NT+1], xfr])[:NT]
    sph[0] = oph[A-1]; sfr[0] = ofr[A-1]
    base += A; total += A; adv.append(A)
adv = np.array(adv)
print("NT", NT, mode, "iters", iters, "samples", total, "mean advance/iter %.1f" % adv.mean(), "min", adv.min(), "median", np.median(adv),
      "p10", np.percentile(adv, 10), "frac iters with A<=4: %.3f" % (adv <= 4).mean())
