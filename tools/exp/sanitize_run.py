"""Small end-to-end run touching every kernel family (for compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from xritdemod_b200 import demod as xd, siggen
N = 400000
x = siggen.generate(siggen.params("hrit", 0, n=N, ramp_len=N), N)
CASES = (dict(), dict(loop_kernel=4), dict(loop_kernel=1), dict(mm_lanes=256),
           dict(mm_lanes=256, mm_walk_lanes=128), dict(mm_walk_lanes=256), dict(mm_lanes=512, mm_rerun=2), dict(mm_kernel=2, mm_lanes=256),
           dict(mm_warm=2000))
for kw in CASES[int(os.environ.get("FIRST_CASE", "0")):]:
    d = xd.Demodulator(mode="hrit")
    t = dict(costas_seg=16384, costas_warm=4096, agc_seg=8192, agc_warm=1024, mm_seg=60000, mm_warm=30000)
    t.update(kw)
    d.set_tuning(**t)
    a = d.demod(x[:250001]); b = d.demod(x[250001:])
    st = d.stats()
    print(kw, len(a) + len(b), st["kernel_launches"], "mm_redo", st["mm_redo"], "mm_bail", st["mm_bail"], flush=True)
x10 = siggen.generate(siggen.params("hrit10", 0, n=N, ramp_len=N), N)
d = xd.Demodulator(mode="hrit", sample_rate=10000000, decimation=4)
print("decimated", len(d.demod(x10)))
d = xd.Demodulator(mode="lrit", n_channels=3)
xs = np.stack([siggen.generate(siggen.params("lrit", c, n=100000, ramp_len=100000), 100000) for c in range(3)])
print("channels", [len(s) for s in d.demod(xs)])
print("i8", d.soft_i8(np.ones(1000, np.complex64))[:3])
d = xd.Demodulator(mode="hrit")
print("fused i8", d.demod_i8(x[:200000])[:3])
