"""An independent float64 model of the three feedback loops and the MMSE interpolator, written from
SURVEY.md Appendix A.3-A.6 (the GNU Radio 3.7 block semantics the reference names as the definition of its
stages, demodulator/demod_tcp_qt.py:265-276) -- NOT from oracle/xrit_oracle.c -- and compared with the oracle
at tolerance.  It pins what a restatement can get wrong without noticing: operation order inside a step
(gain applied before the AGC update, Costas output before the loop advance), the 0/1 slicer, the +-2*pi wrap,
the time-reversed interpolator rows and rint(mu * 128), the omega clipping.  Plain Python loops in double
precision: a second, structurally different statement of the same algorithms.
"""
import math

import numpy as np
import pytest

from conftest import make_signal

N = 1 << 20


# ------------------------------------------------------------------ the model
def model_agc(x, rate=0.01, ref=0.5, gain=1.0, max_gain=4000.0):
    """A.3: y = x * g;  g += rate * (ref - |y|);  clamp to max_gain (output uses the gain BEFORE the update)"""
    re, im = x.real.astype(np.float64).tolist(), x.imag.astype(np.float64).tolist()
    out_r, out_i = [0.0] * len(re), [0.0] * len(re)
    g = float(gain)
    hyp = math.hypot
    for n in range(len(re)):
        yr, yi = re[n] * g, im[n] * g
        out_r[n], out_i[n] = yr, yi
        g += rate * (ref - hyp(yr, yi))
        if max_gain > 0 and g > max_gain:
            g = max_gain
    return np.array(out_r) + 1j * np.array(out_i), g


def model_costas(x, bw=0.0037):
    """A.4: second-order loop, damping sqrt(2)/2; detector Re*Im clipped to +-1; phase wrapped at +-2*pi"""
    damping = math.sqrt(2.0) / 2.0
    denom = 1.0 + 2.0 * damping * bw + bw * bw
    alpha, beta = 4.0 * damping * bw / denom, 4.0 * bw * bw / denom
    re, im = x.real.astype(np.float64).tolist(), x.imag.astype(np.float64).tolist()
    out_r, out_i = [0.0] * len(re), [0.0] * len(re)
    phase = freq = 0.0
    two_pi = 2.0 * math.pi
    cos, sin = math.cos, math.sin
    for n in range(len(re)):
        c, s = cos(-phase), sin(-phase)
        yr = re[n] * c - im[n] * s
        yi = re[n] * s + im[n] * c
        out_r[n], out_i[n] = yr, yi
        e = yr * yi
        e = 1.0 if e > 1.0 else (-1.0 if e < -1.0 else e)
        freq += beta * e
        phase += freq + alpha * e
        while phase > two_pi:
            phase -= two_pi
        while phase < -two_pi:
            phase += two_pi
        freq = 1.0 if freq > 1.0 else (-1.0 if freq < -1.0 else freq)
    return np.array(out_r) + 1j * np.array(out_i), (phase, freq)


def model_mmse_table():
    """A.6: least-squares 8-tap fractional delay over |f| <= B = 0.25, 129 rows; R h = p"""
    B = 0.25
    t = np.arange(8) - 4.0
    R = np.sinc(2 * B * (t[:, None] - t[None, :]))
    rows = [np.linalg.solve(R, np.sinc(2 * B * (t + k / 128.0))) for k in range(129)]
    return np.array(rows)


def model_mm(x, omega, gain_omega, mu, gain_mu, omega_rel_limit):
    """A.5 with the A.6 interpolator: out = sum_j T[k][7 - j] * in[j], k = rint(mu * 128); 0/1 slicer"""
    tab = model_mmse_table()[:, ::-1].tolist()   # time-reversed rows
    re, im = x.real.astype(np.float64).tolist(), x.imag.astype(np.float64).tolist()
    n_in = len(re)
    omega_mid, omega_lim = omega, omega_rel_limit * omega
    p0r = p0i = p1r = p1i = p2r = p2i = 0.0
    c0r = c0i = c1r = c1i = c2r = c2i = 0.0
    ii = 0
    sym_r, sym_i = [], []
    floor = math.floor
    while ii + 8 <= n_in:
        row = tab[int(round(mu * 128.0))]          # Python round() is round-half-even, like rint
        ar = ai = 0.0
        for j in range(8):
            ar += row[j] * re[ii + j]
            ai += row[j] * im[ii + j]
        p2r, p2i, p1r, p1i, p0r, p0i = p1r, p1i, p0r, p0i, ar, ai
        c2r, c2i, c1r, c1i = c1r, c1i, c0r, c0i
        c0r, c0i = (1.0 if p0r > 0 else 0.0), (1.0 if p0i > 0 else 0.0)
        # x = (c0 - c2) * conj(p1);  y = (p0 - p2) * conj(c1);  mm = Re(y - x)
        xr = (c0r - c2r) * p1r + (c0i - c2i) * p1i
        yr = (p0r - p2r) * c1r + (p0i - p2i) * c1i
        mm = yr - xr
        sym_r.append(p0r)
        sym_i.append(p0i)
        mm = 1.0 if mm > 1.0 else (-1.0 if mm < -1.0 else mm)
        omega += gain_omega * mm
        d = omega - omega_mid
        d = omega_lim if d > omega_lim else (-omega_lim if d < -omega_lim else d)
        omega = omega_mid + d
        mu += omega + gain_mu * mm
        f = floor(mu)
        ii += int(f)
        mu -= f
        if ii < 0:
            ii = 0
    return np.array(sym_r) + 1j * np.array(sym_i)


def _rrc_f64(fs, rs, alpha, ntaps):
    """textbook root-raised-cosine impulse response, unit DC gain"""
    t = (np.arange(ntaps) - ntaps // 2) / fs * rs
    h = np.zeros(ntaps)
    for i, ti in enumerate(t):
        if abs(ti) < 1e-12:
            h[i] = 1 - alpha + 4 * alpha / np.pi
        elif abs(abs(ti) - 1 / (4 * alpha)) < 1e-9:
            h[i] = alpha / np.sqrt(2) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) +
                                         (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
        else:
            h[i] = (np.sin(np.pi * ti * (1 - alpha)) + 4 * alpha * ti * np.cos(np.pi * ti * (1 + alpha))) / (
                np.pi * ti * (1 - (4 * alpha * ti) ** 2))
    return h / h.sum()


def _stats(a, b):
    d = a.real - b.real.astype(np.float64)
    return float(np.sqrt(np.mean(d * d))), float(np.abs(d).max()), float(np.mean(np.abs(d) > 5e-4))


# ------------------------------------------------------------------ the comparison
@pytest.fixture(scope="module")
def oracle_stages(oracle):
    _, x = make_signal("hrit", N)
    ch = oracle.Chain(oracle.config(True))
    sym, taps = ch.process(x, taps=True)
    return dict(x=x, sym=sym, sps=ch.sps, **taps)


def test_agc_against_the_float64_model(oracle_stages):
    s = oracle_stages
    y, g = model_agc(s["x"])
    err = np.abs(y - s["agc"]).max() / np.abs(s["agc"]).max()
    print("AGC: max relative deviation oracle (f32) vs float64 model: %.3g" % err)
    assert err < 2e-5


def test_costas_against_the_float64_model(oracle_stages):
    s = oracle_stages
    y, (phase, freq) = model_costas(s["rrc"])
    d = np.abs(y - s["costas"])
    print("Costas: max |dy| %.3g, rms %.3g (oracle f32 vs float64 model, same input)" % (d.max(), np.sqrt(np.mean(d * d))))
    # both lock on the same branch from the same start; what is left is the f32 phase resolution (~5e-7 rad) filtered
    # by the loop
    assert d.max() < 2e-4 and np.sqrt(np.mean(d * d)) < 3e-5


def test_mm_against_the_float64_model(oracle_stages):
    s = oracle_stages
    gm = 0.0037
    sym = model_mm(s["costas"], float(s["sps"]), float(np.float32(gm) * np.float32(gm) / np.float32(4)), 0.5,
                   float(np.float32(gm)), float(np.float32(0.005)))
    assert abs(len(sym) - len(s["sym"])) <= 1
    n = min(len(sym), len(s["sym"]))
    rms, mx, frac = _stats(sym[:n], s["sym"][:n])
    print("M&M: rms %.3g max %.3g fraction(|d| > 5e-4) %.3g over %d symbols (oracle f32 vs float64 model)" % (rms, mx, frac, n))
    # The float32 loop is not the float64 loop plus small noise: omega moves by gain_omega * mm ~ 3e-7 per symbol, about
    # one ulp of a float32 omega (2^-22), so the float32 recurrence rounds most of every increment away and its timing
    # wanders ~1e-3 sample around the exact-arithmetic one; interpolator rows (rint(mu * 128)) then differ on ~10 % of
    # the symbols, ~1e-3 each.  That is the measured distance between two CORRECT implementations in different
    # arithmetic -- and why the GPU path reproduces the float32 recurrence bit for bit instead of approximating it.
    assert rms <= 1.5e-3 and frac < 0.25 and mx < 2e-2
    strong = np.abs(s["sym"][:n].real) > 0.1
    assert np.array_equal(np.sign(sym[:n].real[strong]), np.sign(s["sym"][:n].real[strong]))


def test_whole_chain_against_the_float64_model(oracle_stages):
    """AGC -> RRC -> Costas -> M&M entirely in the float64 model against the oracle's soft symbols"""
    s = oracle_stages
    y, _ = model_agc(s["x"])
    taps = _rrc_f64(2.5e6, 927000.0, float(np.float32(0.3)), 63)
    y = np.convolve(y, taps)[: len(y)]
    y, _ = model_costas(y)
    gm = 0.0037
    sps = 2.5e6 / 927000.0
    sym = model_mm(y, float(np.float32(sps)), float(np.float32(gm) * np.float32(gm) / np.float32(4)), 0.5,
                   float(np.float32(gm)), float(np.float32(0.005)))
    assert abs(len(sym) - len(s["sym"])) <= 1
    n = min(len(sym), len(s["sym"]))
    rms, mx, frac = _stats(sym[:n], s["sym"][:n])
    print("chain: rms %.3g max %.3g fraction(|d| > 5e-4) %.3g over %d symbols (oracle f32 vs float64 model)" % (rms, mx, frac, n))
    # hard decisions agree everywhere the eye is open
    strong = np.abs(s["sym"][:n].real) > 0.1
    assert np.array_equal(np.sign(sym[:n].real[strong]), np.sign(s["sym"][:n].real[strong]))
    assert rms <= 1.5e-3 and frac < 0.25 and mx < 2e-2
