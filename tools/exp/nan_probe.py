"""Which stage stalls on non-finite input?  Each case runs in its own process under a timeout."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASE = r'''
import sys, time
sys.path.insert(0, %r)
import numpy as np
from xritdemod_b200 import demod as xd, siggen
n = 600000
x = siggen.generate(siggen.params("hrit", 0, n=n, ramp_len=n), n)
what, bad = sys.argv[1], float(sys.argv[2])
y = x.copy(); y[300000] = bad; y[300001] = 0
t = time.time()
if what == "agc":
    out = xd.AGC().Work(y)
elif what == "rrc":
    out = xd.FirFilter(1, xd.rrc_taps(1, 2.5e6, 927000.0, 0.3, 63)).Work(y)
elif what == "costas":
    out = xd.CostasLoop().Work(y)
elif what == "mm":
    gm = np.float32(0.0037)
    out = xd.ClockRecovery(2.696872, gm * gm / np.float32(4), 0.5, gm, 0.005).Work(y)
else:
    d = xd.Demodulator(mode="hrit")
    try:
        out = d.demod(y)
    except xd.XrdError as e:
        out = np.zeros(1); print("error", e)
print(what, bad, "done in %%.2f s" %% (time.time() - t), len(out), flush=True)
''' % ROOT
for what in ("agc", "rrc", "costas", "mm", "chain"):
    for bad in ("nan", "inf"):
        try:
            r = subprocess.run([sys.executable, "-c", CASE, what, bad], capture_output=True, text=True, timeout=60)
            print((r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
        except subprocess.TimeoutExpired:
            print(what, bad, "TIMEOUT (60 s)", flush=True)
