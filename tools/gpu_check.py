"""Quick GPU bring-up check: every stage and the chain against the oracle, verbose."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_ffi as o
from xritdemod_b200 import demod as xd, siggen

def cmp(name, a, b):
    a = np.ascontiguousarray(a).view(np.float32).reshape(-1); b = np.ascontiguousarray(b).view(np.float32).reshape(-1)
    if len(a) != len(b):
        print("  %-28s LENGTH %d vs %d" % (name, len(a), len(b))); n = min(len(a), len(b)); a, b = a[:n], b[:n]
    ne = np.nonzero(a != b)[0]
    if len(ne) == 0:
        print("  %-28s BIT-EXACT (%d floats)" % (name, len(a)))
    else:
        print("  %-28s %d/%d differ, first @%d (%r vs %r) max|d| %g rms %g" % (name, len(ne), len(a), ne[0], a[ne[0]], b[ne[0]], np.abs(a-b).max(), np.sqrt(((a-b)**2).mean())))
    return len(ne) == 0

print(xd.device_check(0))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
mode = sys.argv[2] if len(sys.argv) > 2 else "hrit"
p = siggen.params(mode, 0, n=N, ramp_len=min(N, 1 << 20)); x = siggen.generate(p, N)
cfg = o.config(mode == "hrit")
ch = o.Chain(cfg); t = time.time(); sym, taps = ch.process(x, taps=True); print("oracle %.2f s, %d symbols" % (time.time() - t, len(sym)))
sps = ch.sps
# stages
rt = xd.rrc_taps(1, cfg.sample_rate, cfg.symbol_rate, cfg.rrc_alpha, 63)
cmp("rrc taps", rt, o.rrc_taps(cfg.sample_rate, cfg.symbol_rate, cfg.rrc_alpha, 63))
cmp("mmse table", xd.mmse_table(), o.mmse_table())
t = time.time(); a = xd.AGC(0.01, 0.5, 1.0, 4000.0).Work(x); print("agc %.3f s" % (time.time() - t)); cmp("agc", a, taps["agc"])
t = time.time(); r = xd.FirFilter(1, rt).Work(taps["agc"]); print("rrc %.3f s" % (time.time() - t)); cmp("rrc fir", r, taps["rrc"])
t = time.time(); c = xd.CostasLoop(0.0037, 2).Work(taps["rrc"]); print("costas %.3f s" % (time.time() - t)); cmp("costas", c, taps["costas"])
gm = np.float32(0.0037); go = np.float32(gm * gm / np.float32(4))
mm = xd.ClockRecovery(sps, go, 0.5, gm, 0.005)
t = time.time(); s = mm.Work(taps["costas"]); print("mm %.3f s" % (time.time() - t)); cmp("mm", s, sym)
# small segments to exercise fix-up
mm2 = xd.ClockRecovery(sps, go, 0.5, gm, 0.005); mm2.set_tuning(20000, 30000)
t = time.time(); s = mm2.Work(taps["costas"]); print("mm small-seg %.3f s" % (time.time() - t)); cmp("mm small seg", s, sym)
d = xd.Demodulator(mode=mode)
t = time.time(); s = d.demod(x); print("chain %.3f s" % (time.time() - t)); cmp("chain", s, sym); print(d.stats())
d2 = xd.Demodulator(mode=mode)
parts = [d2.demod(x[i:i + 300000]) for i in range(0, N, 300000)]
cmp("chain chunked", np.concatenate(parts), sym); print(d2.stats())
