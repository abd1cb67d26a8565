"""End-to-end (host buffers) timing of xrd_demod_batch for a list of tunings; checks the digest stays the same.
usage: python tools/gpu_e2e.py N_SAMPLES 'h2d_pieces=1' 'h2d_pieces=4' ..."""
import ctypes as C, hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from xritdemod_b200 import demod, siggen

n = int(sys.argv[1])
p = siggen.params("hrit", 0, n=n, ramp_len=1 << 20)
h = torch.empty(2 * n, dtype=torch.float32).pin_memory()
siggen.generate(p, n, out=h.numpy().view(np.complex64))
ref = None
for spec in sys.argv[2:] or [""]:
    kw = {k: int(v) for k, v in (kv.split("=") for kv in spec.split(",") if kv)}
    d = demod.Demodulator(mode="hrit")
    if kw:
        d.set_tuning(**kw)
    cap = d.symbol_capacity(n)
    hs = torch.empty(2 * cap, dtype=torch.float32).pin_memory()
    cnt = np.zeros(1, np.int64)
    best = None
    for rep in range(4):
        d.reset()
        torch.cuda.synchronize(); t = time.perf_counter()
        rc = demod.lib().xrd_demod_batch(d._h, C.c_void_p(h.data_ptr()), n, 0, C.c_void_p(hs.data_ptr()), cap,
                                         cnt.ctypes.data_as(C.POINTER(C.c_int64)))
        dt = (time.perf_counter() - t) * 1e3
        assert rc == 0, rc
        st = d.stats()
        if best is None or dt < best[0]:
            best = (dt, st)
    ns = int(cnt[0])
    digest = hashlib.sha1(hs[: 2 * ns].numpy().tobytes()).hexdigest()[:12]
    ref = ref or digest
    st = best[1]
    print("%-28s e2e %8.2f ms (%.0f Msps) | agc %.2f rrc %.2f costas %.2f mm %.2f | nsym %d sha %s %s" % (
        spec, best[0], n / best[0] / 1e3, st["ms_agc"], st["ms_fir_rrc"], st["ms_costas"], st["ms_mm"], ns, digest,
        "OK" if digest == ref else "MISMATCH"), flush=True)
