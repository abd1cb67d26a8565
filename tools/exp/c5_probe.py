import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from xritdemod_b200 import demod, siggen
nch, n5 = int(sys.argv[1]), int(sys.argv[2])
h = torch.empty((nch, 2 * n5), dtype=torch.float32)
for c in range(nch):
    siggen.generate(siggen.params("lrit", c, n=n5, ramp_len=1 << 20), n5, out=h[c].numpy().view(np.complex64))
x = h.cuda()
d = demod.Demodulator(mode="lrit", n_channels=nch)
cap = d.symbol_capacity(n5)
sym = torch.empty((nch, 2 * cap), dtype=torch.float32, device="cuda")
for rep in range(2):
    d.reset(); torch.cuda.synchronize(); t = time.perf_counter()
    d.demod_device(x.data_ptr(), n5, sym.data_ptr(), cap); torch.cuda.synchronize()
    print((time.perf_counter() - t) * 1e3, d.stats())
