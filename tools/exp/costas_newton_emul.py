"""Picard (prefix sum of literal differences) vs true Newton (per-slot Jacobian, affine recurrence) window iteration for Costas."""
import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_ffi as o
from xritdemod_b200 import siggen
f32 = np.float32
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 128
method = sys.argv[2] if len(sys.argv) > 2 else "newton"
TOTAL = int(sys.argv[3]) if len(sys.argv) > 3 else 40000
N = 1 << 19
p = siggen.params("hrit", 0, n=N, ramp_len=1 << 20); x = siggen.generate(p, N)
ch = o.Chain(o.config(True)); sym, taps = ch.process(x, taps=True)
r = taps["rrc"]
alpha, beta = o.costas_gains(0.0037); alpha = f32(alpha); beta = f32(beta)
TWO_PI_F = f32(6.28318500518798828125); TWO_PI = 2 * np.pi
def step(ph, fr, xr, xi, jac=False):
    cs = np.cos(-ph).astype(f32); sn = np.sin(-ph).astype(f32)
    yr = (xr * cs - xi * sn).astype(f32); yi = (xr * sn + xi * cs).astype(f32)
    e0 = (yr * yi).astype(f32)
    e = (f32(0.5) * (np.abs(e0 + f32(1)) - np.abs(e0 - f32(1)))).astype(f32)
    fr2 = (fr + beta * e).astype(f32)
    ph2 = ((ph + fr2).astype(f32) + (alpha * e).astype(f32)).astype(f32)
    hi = ph2 > TWO_PI_F; lo = ph2 < -TWO_PI_F
    ph2 = np.where(hi, (ph2.astype(np.float64) - TWO_PI).astype(f32), ph2)
    ph2 = np.where(lo, (ph2.astype(np.float64) + TWO_PI).astype(f32), ph2)
    fr2 = np.clip(fr2, f32(-1), f32(1))
    if jac:
        ed = np.where(np.abs(e0) < 1, yi.astype(np.float64) ** 2 - yr.astype(np.float64) ** 2, 0.0)
        return ph2, fr2, ed
    return ph2, fr2
def wrapd(d):   # phase difference modulo 2 pi into (-pi, pi]
    return d - TWO_PI * np.round(d / TWO_PI)
def norm(P):
    P = np.where(P > TWO_PI, P - TWO_PI, P); P = np.where(P < -TWO_PI, P + TWO_PI, P); return P
xr = r.real.copy(); xi = r.imag.copy()
start = 100000
ph = f32(0); fr = f32(0)
for i in range(start - 60000, start):
    a, b = step(np.array([ph]), np.array([fr]), xr[i:i+1], xi[i:i+1]); ph, fr = a[0], b[0]
base = start
def extrap(eph, efr, k0, k1):
    k = np.arange(k0, k1, dtype=np.float64)
    P = float(eph) + k * float(efr)
    P = P - TWO_PI * np.trunc(P / TWO_PI)
    return P.astype(f32), np.full(len(k), efr, f32)
sph, sfr = extrap(ph, fr, 0, NT); sph[0] = ph; sfr[0] = fr
adv = []; total = 0; iters = 0
while total < TOTAL:
    iters += 1
    oph, ofr, ed = step(sph, sfr, xr[base:base+NT], xi[base:base+NT], jac=True)
    ok = (sph[1:] == oph[:-1]) & (sfr[1:] == ofr[:-1])
    A = NT if ok.all() else int(np.argmin(ok)) + 1
    if method == "picard":
        Dp = wrapd(oph.astype(np.float64) - sph.astype(np.float64)); Df = ofr.astype(np.float64) - sfr.astype(np.float64)
        P = norm(float(sph[0]) + np.cumsum(Dp)); Fq = float(sfr[0]) + np.cumsum(Df)
        nph = np.concatenate([sph[:1], P.astype(f32)]); nfr = np.concatenate([sfr[:1], Fq.astype(f32)])
    else:
        # residuals res_r = o_r - s_{r+1}; delta_{r+1} = J_r delta_r + res_r; last slot predicts one beyond the window
        dph = np.zeros(NT + 1); dfr = np.zeros(NT + 1)
        for q in range(NT):
            if q + 1 < NT:
                rp = wrapd(float(oph[q]) - float(sph[q + 1])); rf = float(ofr[q]) - float(sfr[q + 1])
            else:
                rp = 0.0; rf = 0.0
            g = ed[q]
            dph[q + 1] = (1 + (float(alpha) + float(beta)) * g) * dph[q] + dfr[q] + rp
            dfr[q + 1] = float(beta) * g * dph[q] + dfr[q] + rf
        nph = norm(sph.astype(np.float64) + dph[:NT]).astype(f32); nfr = (sfr.astype(np.float64) + dfr[:NT]).astype(f32)
        # state after the window = literal output of the last slot shifted by its correction
        endp = norm(np.array([float(oph[NT - 1]) + dph[NT]]))[0]; endf = float(ofr[NT - 1]) + dfr[NT]
        nph = np.concatenate([nph, [f32(endp)]]); nfr = np.concatenate([nfr, [f32(endf)]])
    eph, efr = nph[NT], nfr[NT]
    xph, xfr = extrap(eph, efr, 1, A + 1)
    sph = np.concatenate([nph[A:NT+1], xph])[:NT].astype(f32); sfr = np.concatenate([nfr[A:NT+1], xfr])[:NT].astype(f32)
    sph[0] = oph[A-1]; sfr[0] = ofr[A-1]
    base += A; total += A; adv.append(A)
adv = np.array(adv)
print("NT", NT, method, "iters", iters, "samples", total, "mean advance/iter %.1f" % adv.mean(), "median", np.median(adv),
      "evaluations per sample %.2f" % (NT * iters / total))
