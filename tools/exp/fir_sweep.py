"""RRC FIR time on the 125 M-sample stream for a tap sweep (stage time from xrd_get_stats, CUDA events)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from xritdemod_b200 import demod, siggen
n = 125_000_000
h = torch.empty(2 * n, dtype=torch.float32).pin_memory()
siggen.generate(siggen.params("hrit", 0, n=n, ramp_len=1 << 20), n, out=h.numpy().view(np.complex64))
x = h.cuda()
for taps in [int(t) for t in os.environ.get("TAPS", "7,15,31,63,127,255").split(",")]:
    d = demod.Demodulator(mode="hrit", rrc_taps=taps)
    cap = d.symbol_capacity(n)
    sym = torch.empty(2 * cap, dtype=torch.float32, device="cuda")
    best = 1e9
    for rep in range(3):
        d.reset()
        d.demod_device(x.data_ptr(), n, sym.data_ptr(), cap)
        best = min(best, d.stats()["ms_fir_rrc"])
    print("taps %3d  rrc %.3f ms  stage traffic %.0f GB/s (%.1f %% of 6547.8)  %.1f TFLOP/s" % (
        taps, best, 16.0 * n / best / 1e6, 16.0 * n / best / 1e6 / 65.478, 4.0 * taps * n / best / 1e9), flush=True)
    d.close()
