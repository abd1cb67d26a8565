"""Round-2 kernels under compute-sanitizer: small end-to-end runs through the TMA FIR, the record-guided Costas re-runs
(three CTA shapes, guided on and off, chased segments), the fused S16 / raw-format ingest, the diag tap, checkpoint /
restore and the decoder front half.  usage: compute-sanitizer --tool memcheck|racecheck python tools/exp/sanitize_run_r2.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from xritdemod_b200 import demod as xd, siggen

N = 300000
x = siggen.generate(siggen.params("hrit", 7, n=N, ramp_len=N), N)
small = dict(costas_seg=16384, costas_warm=2048, agc_seg=8192, agc_warm=1024, mm_seg=60000, mm_warm=30000)
ref = None
for kw in (dict(), dict(guided=2), dict(rerun_kernel=3), dict(rerun_kernel=7), dict(rerun_kernel=4, chase=2), dict(loop_kernel=4)):
    d = xd.Demodulator(mode="hrit")
    t = dict(small)
    t.update(kw)
    d.set_tuning(**t)
    a = d.demod(x[:170001])
    blob = d.checkpoint()
    b = d.demod(x[170001:])
    d.restore(blob)
    b2 = d.demod(x[170001:])
    st = d.stats()
    sym = np.concatenate([a, b])
    if ref is None:
        ref = sym
    print(kw, len(sym), "same" if np.array_equal(sym.view(np.uint32), ref.view(np.uint32)) and np.array_equal(b, b2) else "DIFFERENT",
          "launches", st["kernel_launches"], "costas_redo", st["costas_redo"], "diag", round(d.diag().snr_db, 1), flush=True)
for taps in (15, 63, 255):
    d = xd.Demodulator(mode="hrit", rrc_taps=taps)
    print("rrc taps", taps, len(d.demod(x[:100000])), flush=True)
d = xd.Demodulator(mode="hrit")
d.set_tuning(**small)
print("s16 fused", len(d.demod(siggen.to_s16(x[:150000]), xd.XRD_S16IQ)), "u8", len(d.demod(siggen.to_u8(x[:150000]), xd.XRD_U8IQ)), flush=True)
x10 = siggen.generate(siggen.params("hrit10", 0, n=N, ramp_len=N), N)
d = xd.Demodulator(mode="hrit", sample_rate=10000000, decimation=4)
print("decimated s16", len(d.demod(siggen.to_s16(x10), xd.XRD_S16IQ)), "s8", len(d.demod(siggen.to_s8(x10), xd.XRD_S8IQ)), flush=True)
d = xd.Demodulator(mode="hrit")
soft = d.demod_i8(x)
f = xd.DecoderFront(lrit=False, soft_mode=1)
frames, meta, cons = f.run(soft)
print("decoder front", len(soft), "soft bytes ->", len(frames), "frames, consumed", cons, "correlate", f.correlate(soft[:20000].view(np.uint8)), flush=True)
