"""Seeded fuzz of the streaming seam: random call boundaries (biased to sit a few samples after the positions the
kernels treat specially: checkpoints every 2048 samples, segment ends, the 16-sample tiles), random segmentations,
kernel shapes and re-run strategies, on the eight C3 streams, against the oracle demodulating the same stream in one
piece.  The demodulated symbols must not depend on any of it, bit for bit.  (A sibling of the script that found the
stale-checkpoint bug of round 2, profiles/r02_compute_sanitizer.md.)"""
import os

import numpy as np
import pytest

from conftest import assert_bitexact, make_signal

pytestmark = pytest.mark.gpu
N = 600000
SEEDS = int(os.environ.get("XRD_FUZZ_SEEDS", "12"))   # a longer campaign: XRD_FUZZ_SEEDS=300 pytest tests/test_gpu_fuzz.py
BASE = int(os.environ.get("XRD_FUZZ_BASE", "0"))      # ... on other seeds: XRD_FUZZ_BASE=100000
SPECIAL = (2048, 4096, 8192, 16384, 32768, 65536, 60000, 100000)


def _cuts(rng, n, k):
    out = set()
    while len(out) < k:
        if rng.random() < 0.7:
            m = int(rng.choice(SPECIAL))
            c = m * int(rng.integers(1, max(2, n // m))) + int(rng.integers(-3, 41))
        else:
            c = int(rng.integers(1, n))
        if 0 < c < n:
            out.add(c)
    return sorted(out)


def _tuning(rng):
    t = {}
    if rng.random() < 0.8:
        t["costas_seg"] = int(rng.choice([4096, 8192, 16384, 32768, 65536]))
        t["costas_warm"] = int(rng.choice([512, 2048, 4096, 16384]))
    if rng.random() < 0.6:
        t["agc_seg"] = int(rng.choice([2048, 4096, 8192, 32768]))
        t["agc_warm"] = int(rng.choice([256, 1024, 4096]))
    if rng.random() < 0.7:
        t["mm_seg"] = int(rng.choice([40000, 60000, 100000, 150000]))
        t["mm_warm"] = int(rng.choice([3000, 20000, 30000, 80000]))
    r = rng.random()
    if r < 0.25:
        t["loop_kernel"] = int(rng.choice([2, 3, 4, 5, 6, 7]))
    elif r < 0.5:
        t["rerun_kernel"] = int(rng.choice([2, 3, 4, 5, 6, 7]))
    if rng.random() < 0.3:
        t["guided"] = 2
    if rng.random() < 0.3:
        t["chase"] = 2
    if rng.random() < 0.3:
        t["mm_walk_lanes"] = int(rng.choice([128, 256, 512]))
    if rng.random() < 0.15:
        t["mm_rerun"] = 2
    if rng.random() < 0.15:
        t["mm_lanes"] = int(rng.choice([256, 512]))
    return t


@pytest.mark.parametrize("seed", range(SEEDS))
def test_random_call_boundaries_and_tunings(gpu, xrd, oracle, siggen, seed):
    rng = np.random.default_rng(BASE + 1000 + seed)
    for case in range(6):
        mode = "hrit" if rng.random() < 0.7 else "lrit"
        channel = int(rng.integers(0, 8))
        _, x = make_signal(mode, N, channel=channel)
        s16 = rng.random() < 0.25   # the receiver's int16 IQ, converted by the first kernel of the chain
        if s16:
            raw = siggen.to_s16(x)
            ref = oracle.Chain(oracle.config(mode == "hrit")).process(oracle.convert_s16(raw))
        else:
            ref = oracle.Chain(oracle.config(mode == "hrit")).process(x)
        tune = _tuning(rng)
        cuts = [0] + _cuts(rng, N, int(rng.integers(1, 5))) + [N]
        d = xrd.Demodulator(mode=mode)
        if tune:
            d.set_tuning(**tune)
        if s16:
            sym = np.concatenate([d.demod(raw[2 * a:2 * b], type=1) for a, b in zip(cuts[:-1], cuts[1:])])
        else:
            sym = np.concatenate([d.demod(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
        assert_bitexact(sym, ref, "seed %d case %d: %s ch %d s16 %d, cuts %s, tuning %s" % (
            seed, case, mode, channel, s16, cuts[1:-1], tune))


@pytest.mark.parametrize("seed", range(max(3, SEEDS // 4)))
def test_random_multi_channel_batches(gpu, xrd, oracle, seed):
    """several channels per call (configs[4] shape), ragged calls: every channel equals its own one-piece oracle run"""
    rng = np.random.default_rng(BASE + 3000 + seed)
    nch = int(rng.integers(2, 6))
    n = 300000
    xs = np.stack([make_signal("lrit", n, channel=c)[1] for c in range(nch)])
    refs = [oracle.Chain(oracle.config(False)).process(xs[c]) for c in range(nch)]
    for case in range(3):
        tune = _tuning(rng)
        tune.pop("mm_lanes", None)
        cuts = [0] + _cuts(rng, n, int(rng.integers(1, 4))) + [n]
        d = xrd.Demodulator(mode="lrit", n_channels=nch)
        if tune:
            d.set_tuning(**tune)
        parts = [d.demod(np.ascontiguousarray(xs[:, a:b])) for a, b in zip(cuts[:-1], cuts[1:])]
        for c in range(nch):
            assert_bitexact(np.concatenate([p[c] for p in parts]), refs[c], "seed %d case %d channel %d/%d cuts %s tuning %s" % (
                seed, case, c, nch, cuts[1:-1], tune))


@pytest.mark.parametrize("seed", range(max(3, SEEDS // 4)))
def test_random_ragged_stage_calls(gpu, xrd, oracle, seed):
    """the same for the three loop operators on their own (SatHelper seam: Work(in, out, n) with any n)"""
    rng = np.random.default_rng(BASE + 2000 + seed)
    channel = int(rng.integers(0, 8))
    _, x = make_signal("hrit", N, channel=channel)
    ch = oracle.Chain(oracle.config(True))
    sym, taps = ch.process(x, taps=True)
    for case in range(4):
        cuts = [0] + _cuts(rng, N, int(rng.integers(1, 5))) + [N]
        kernel = int(rng.choice([2, 3, 4, 5, 6, 7]))
        a = xrd.AGC()
        a.set_loop_kernel(kernel)
        a.set_tuning(int(rng.choice([2048, 4096, 16384])), int(rng.choice([256, 1024])))
        y = np.concatenate([a.Work(x[p:q]) for p, q in zip(cuts[:-1], cuts[1:])])
        assert_bitexact(y, taps["agc"], "AGC kernel %d cuts %s" % (kernel, cuts[1:-1]))
        c = xrd.CostasLoop()
        c.set_loop_kernel(kernel)
        c.set_tuning(int(rng.choice([4096, 16384, 32768])), int(rng.choice([512, 2048, 8192])))
        y = np.concatenate([c.Work(taps["rrc"][p:q]) for p, q in zip(cuts[:-1], cuts[1:])])
        assert_bitexact(y, taps["costas"], "Costas kernel %d cuts %s" % (kernel, cuts[1:-1]))
        gm = np.float32(0.0037)
        m = xrd.ClockRecovery(ch.sps, gm * gm / np.float32(4), 0.5, gm, 0.005)
        m.set_tuning(int(rng.choice([40000, 100000])), int(rng.choice([20000, 30000, 80000])))
        y = np.concatenate([m.Work(taps["costas"][p:q]) for p, q in zip(cuts[:-1], cuts[1:])])
        assert_bitexact(y, sym, "M&M cuts %s" % (cuts[1:-1],))
