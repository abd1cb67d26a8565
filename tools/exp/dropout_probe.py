"""stage times / rounds on streams that lose the signal for a while (usage: dropout_probe.py [tuning ...])"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from xritdemod_b200 import demod, siggen
n = 32_000_000
x = siggen.generate(siggen.params("hrit", 0, n=n, ramp_len=1 << 20), n)
cases = {"clean": x}
for burst in (100_000, 1_000_000):
    y = x.copy(); y[n // 2: n // 2 + burst] = siggen.noise_only(7, burst); cases["noise burst %d" % burst] = y
    z = x.copy(); z[n // 2: n // 2 + burst] = 0; cases["zero burst %d" % burst] = z
cases["noise only 4M"] = siggen.noise_only(11, 4_000_000)
for spec in sys.argv[1:] or [""]:
    kw = {k: int(v, 0) for k, v in (kv.split("=") for kv in spec.split(",") if kv)}
    for name, buf in cases.items():
        d = demod.Demodulator(mode="hrit")
        if kw: d.set_tuning(**kw)
        cap = d.symbol_capacity(len(buf))
        xd = torch.from_numpy(buf.view(np.float32)).cuda(); sym = torch.empty(2 * cap, dtype=torch.float32, device="cuda")
        d.demod_device(xd.data_ptr(), len(buf), sym.data_ptr(), cap); d.reset()
        s0 = d.stats(); torch.cuda.synchronize(); t = time.perf_counter()
        d.demod_device(xd.data_ptr(), len(buf), sym.data_ptr(), cap); torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
        s = d.stats()
        print("%-12s %-22s %9.2f ms | agc %.2f rrc %.2f costas %.2f mm %.2f | rounds a/c/m %d/%d/%d redo %d/%d/%d bail %d launches %d" % (
            spec, name, dt, s["ms_agc"], s["ms_fir_rrc"], s["ms_costas"], s["ms_mm"], s["agc_rounds"] - s0["agc_rounds"],
            s["costas_rounds"] - s0["costas_rounds"], s["mm_rounds"] - s0["mm_rounds"], s["agc_redo"] - s0["agc_redo"],
            s["costas_redo"] - s0["costas_redo"], s["mm_redo"] - s0["mm_redo"], s["mm_bail"] - s0["mm_bail"],
            s["kernel_launches"] - s0["kernel_launches"]), flush=True)
