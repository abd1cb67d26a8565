"""Small end-to-end run touching every kernel family (for compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from xritdemod_b200 import demod as xd, siggen
N = 400000
x = siggen.generate(siggen.params("hrit", 0, n=N, ramp_len=N), N)
for kw in (dict(), dict(loop_kernel=4), dict(loop_kernel=1), dict(mm_lanes=0x20000 + (2 << 8) + 16), dict(mm_lanes=256)):
    d = xd.Demodulator(mode="hrit")
    d.set_tuning(costas_seg=16384, costas_warm=4096, agc_seg=8192, agc_warm=1024, mm_seg=60000, mm_warm=30000, **kw)
    a = d.demod(x[:250001]); b = d.demod(x[250001:])
    print(kw, len(a) + len(b), d.stats()["kernel_launches"], flush=True)
x10 = siggen.generate(siggen.params("hrit10", 0, n=N, ramp_len=N), N)
d = xd.Demodulator(mode="hrit", sample_rate=10000000, decimation=4)
print("decimated", len(d.demod(x10)))
d = xd.Demodulator(mode="lrit", n_channels=3)
xs = np.stack([siggen.generate(siggen.params("lrit", c, n=100000, ramp_len=100000), 100000) for c in range(3)])
print("channels", [len(s) for s in d.demod(xs)])
print("i8", d.soft_i8(np.ones(1000, np.complex64))[:3])
