"""GPU parity at the sizes BASELINE.json quotes the metric on (`-m gpu`), against the CPU oracle run on the
same seeded inputs on the box's host cores (one oracle chain per thread; ctypes releases the GIL):

  configs[2]  C3: the eight 125 000 000-sample HRIT streams of the 8-GPU run (seeds 0x5EED0000 + 0..7), stream 7
              (-48 Hz carrier offset: the slow-merging Costas case) included -- stream 0 is covered in
              test_gpu_parity.py::test_full_size_stream_properties
  configs[3]  C4: 125 000 000 input samples at 10 Msps, decimation 4 (241-tap LPF), 63-tap RRC
  configs[4]  C5: 256 LRIT channels x 4 194 304 samples in ONE call (n_channels = 256)

Every comparison is bit-exact (which implies the north_star's 1e-4 RMS).
"""
import threading

import numpy as np
import pytest

from conftest import assert_bitexact

pytestmark = pytest.mark.gpu
N_STREAM = 125_000_000


def _oracle_async(oracle, cfg, x):
    """run one oracle chain over x on its own thread; returns (thread, result holder)"""
    out = {}

    def work():
        out["sym"] = oracle.Chain(cfg).process(x)

    t = threading.Thread(target=work)
    t.start()
    return t, out


def test_c3_streams_1_to_7_at_full_size(gpu, xrd, oracle, siggen):
    d = xrd.Demodulator(mode="hrit")
    pending = []
    times = {}
    for stream in range(1, 8):
        p = siggen.params("hrit", stream, n=N_STREAM, ramp_len=1 << 20)
        x = siggen.generate(p, N_STREAM)
        d.reset()
        got = d.demod(x)
        st = d.stats()
        times[stream] = (p.carrier_hz, st["ms_agc"] + st["ms_fir_rrc"] + st["ms_costas"] + st["ms_mm"], st["ms_costas"])
        pending.append((stream, got) + _oracle_async(oracle, oracle.config(True), x))
        while len(pending) > 3:                     # bound the host memory held by streams in flight
            s, g, t, out = pending.pop(0)
            t.join()
            assert_bitexact(g, out["sym"], "C3 stream %d at 125 M samples" % s)
    for s, g, t, out in pending:
        t.join()
        assert_bitexact(g, out["sym"], "C3 stream %d at 125 M samples" % s)
    print("C3 per stream (carrier Hz, chain ms, Costas ms):", {k: tuple(round(v, 2) for v in t) for k, t in times.items()})


def test_c4_decimated_chain_at_full_size(gpu, xrd, oracle, siggen):
    kw = dict(sample_rate=10000000, decimation=4, rrc_taps=63)
    p = siggen.params("hrit10", 0, n=N_STREAM, ramp_len=1 << 20)
    x = siggen.generate(p, N_STREAM)
    t, out = _oracle_async(oracle, oracle.config(True, **kw), x)
    d = xrd.Demodulator(mode="hrit", **kw)
    got = d.demod(x)
    assert d.state().n_in == N_STREAM
    t.join()
    assert_bitexact(got, out["sym"], "C4: 125 M samples at 10 Msps, decimation 4, 63-tap RRC")


def test_c5_256_channels_x_4mi_in_one_call(gpu, xrd, oracle, siggen):
    nch, n = 256, 1 << 22
    xs = np.empty((nch, n), np.complex64)
    for c in range(nch):
        siggen.generate(siggen.params("lrit", c, n=n, ramp_len=1 << 20), n, out=xs[c])
    d = xrd.Demodulator(mode="lrit", n_channels=nch)
    got = d.demod(xs)
    assert len(got) == nch
    cfg = oracle.config(False)
    refs = [None] * nch
    import os

    nthreads = max(1, min(32, os.cpu_count() or 1))

    def work(k):
        for c in range(k, nch, nthreads):
            refs[c] = oracle.Chain(cfg).process(xs[c])

    th = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
    [t.start() for t in th]
    [t.join() for t in th]
    for c in range(nch):
        assert_bitexact(got[c], refs[c], "C5 channel %d of 256 x 4 Mi" % c)
