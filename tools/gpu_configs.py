"""Device-resident throughput of the other BASELINE.json configs (parity for them is in tests/test_gpu_parity.py):
  C4: 10 Msps HRIT input, decimation 4 (241-tap Hamming LPF), RRC tap sweep {15,31,63,127,255}
  C5: 256 concurrent LRIT channels, 4 Mi samples each, one call
usage: python tools/gpu_configs.py [c4_samples] [c5_channels] [c5_samples]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from xritdemod_b200 import demod, siggen

def timed(d, x, n, cap, sym, reps=3):
    best = None
    for _ in range(reps):
        d.reset()
        torch.cuda.synchronize(); t = time.perf_counter()
        cnt = d.demod_device(x.data_ptr(), n, sym.data_ptr(), cap)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
        best = dt if best is None else min(best, dt)
    return best, cnt, d.stats()

n4 = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n5 = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 22
out = {"c4": [], "c5": None}
p = siggen.params("hrit10", 0, n=n4, ramp_len=1 << 20)
h = torch.empty(2 * n4, dtype=torch.float32)
siggen.generate(p, n4, out=h.numpy().view(np.complex64))
x = h.cuda(); del h
for taps in (15, 31, 63, 127, 255):
    d = demod.Demodulator(mode="hrit", sample_rate=10000000, decimation=4, rrc_taps=taps)
    cap = d.symbol_capacity(n4)
    sym = torch.empty(2 * cap, dtype=torch.float32, device="cuda")
    ms, cnt, st = timed(d, x, n4, cap, sym)
    rec = dict(rrc_taps=taps, ms=ms, msps=n4 / ms / 1e3, nsym=int(cnt[0]),
               stage_ms={k: st[k] for k in ("ms_fir_dec", "ms_agc", "ms_fir_rrc", "ms_costas", "ms_mm")})
    out["c4"].append(rec); print("C4", json.dumps(rec), flush=True)
    d.close(); del sym
del x; torch.cuda.empty_cache()
# C5: distinct seed / carrier / timing per channel
h = torch.empty((nch, 2 * n5), dtype=torch.float32)
for c in range(nch):
    siggen.generate(siggen.params("lrit", c, n=n5, ramp_len=1 << 20), n5, out=h[c].numpy().view(np.complex64))
x = h.cuda(); del h
d = demod.Demodulator(mode="lrit", n_channels=nch)
cap = d.symbol_capacity(n5)
sym = torch.empty((nch, 2 * cap), dtype=torch.float32, device="cuda")
ms, cnt, st = timed(d, x, n5, cap, sym)
rec = dict(channels=nch, samples_per_channel=n5, ms=ms, msps=nch * n5 / ms / 1e3, nsym_total=int(cnt.sum()),
           stage_ms={k: st[k] for k in ("ms_agc", "ms_fir_rrc", "ms_costas", "ms_mm")},
           fixups={k: st[k] for k in ("agc_redo", "costas_redo", "mm_redo", "mm_rounds", "costas_rounds")})
out["c5"] = rec; print("C5", json.dumps(rec), flush=True)
