import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
exec(open(os.path.join(ROOT, "tools/exp/costas_window_emul.py")).read().split("NT = int")[0])
NT = 1024
xr = r.real.copy(); xi = r.imag.copy()
start = 200000
ph = f32(0); fr = f32(0)
for i in range(start - 60000, start):
    a, b = step(np.array([ph]), np.array([fr]), xr[i:i+1], xi[i:i+1]); ph, fr = a[0], b[0]
# true trajectory
tph = np.empty(NT + 1, f32); tfr = np.empty(NT + 1, f32); tph[0] = ph; tfr[0] = fr
for i in range(NT):
    a, b = step(tph[i:i+1], tfr[i:i+1], xr[start+i:start+i+1], xi[start+i:start+i+1]); tph[i+1] = a[0]; tfr[i+1] = b[0]
idx = np.arange(NT)
sph = (np.float64(ph) + idx * np.float64(fr)).astype(f32); sfr = np.full(NT, fr, f32)
for it in range(12):
    oph, ofr = step(sph, sfr, xr[start:start+NT], xi[start:start+NT])
    dph = oph.astype(np.float64) - sph.astype(np.float64); dfr = ofr.astype(np.float64) - sfr.astype(np.float64)
    P = np.float64(sph[0]) + np.cumsum(dph); Fq = np.float64(sfr[0]) + np.cumsum(dfr)
    sph = np.concatenate([[sph[0]], P[:-1].astype(f32)]); sfr = np.concatenate([[sfr[0]], Fq[:-1].astype(f32)])
    eq = (sph == tph[:NT]) & (sfr == tfr[:NT])
    first_bad = int(np.argmin(eq)) if not eq.all() else NT
    eph = np.abs(sph.astype(np.float64) - tph[:NT]); efr = np.abs(sfr.astype(np.float64) - tfr[:NT])
    print(it, "exact prefix", first_bad, "n exact", eq.sum(), "phase err @64,256,1023: %.2e %.2e %.2e" % (eph[64], eph[256], eph[1023]),
          "freq err: %.2e %.2e %.2e" % (efr[64], efr[256], efr[1023]), "ph,fr true@prefix", tph[first_bad] if first_bad<NT else None)
