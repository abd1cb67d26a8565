"""Synthetic LRIT/HRIT-shaped BPSK IQ streams (host side, deterministic).

Stands in for the reference's offline source, CFileFrontend
(reference demodulator/src/CFileFrontend.cpp:33-62): raw interleaved complex<float>.
Every sample is a pure function of (seed, absolute sample index) -- see csrc/siggen.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libxrdsig.so")
_SRC = os.path.join(_HERE, "csrc", "siggen.c")
_LIB = None


class SigParams(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_double),
        ("symbol_rate", C.c_double),
        ("rrc_alpha", C.c_double),
        ("timing_offset", C.c_double),
        ("carrier_hz", C.c_double),
        ("phase0", C.c_double),
        ("amp_start", C.c_double),
        ("amp_end", C.c_double),
        ("ramp_len", C.c_uint64),
        ("esn0_db", C.c_double),
        ("noise", C.c_int32),
        ("reserved", C.c_int32),
        ("seed", C.c_uint64),
    ]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SRC) > os.path.getmtime(_SO):
        subprocess.check_call(
            ["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"]
        )
    return _SO


def noise_only(seed, n, sigma=0.2):
    """complex white noise, no signal: what the chain sees when the downlink drops out"""
    rng = np.random.default_rng(seed)
    out = np.empty(n, np.complex64)
    for a in range(0, n, 1 << 22):
        b = min(n, a + (1 << 22))
        out[a:b] = (sigma * (rng.standard_normal(b - a) + 1j * rng.standard_normal(b - a))).astype(np.complex64)
    return out


def _lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.xrd_siggen_cf32.argtypes = [C.POINTER(SigParams), C.c_uint64, C.c_uint64, C.c_void_p]
        L.xrd_siggen_bits.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_void_p]
        L.xrd_cf32_to_s16.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.xrd_cf32_to_s8.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.xrd_cf32_to_u8.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.xrd_siggen_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


MODES = {
    # mode: (sample_rate, symbol_rate, rrc_alpha) -- reference Parameters.h:16-24
    "lrit": (1.25e6, 293883.0, 0.5),
    "hrit": (2.5e6, 927000.0, 0.3),
    "hrit10": (10.0e6, 927000.0, 0.3),
}


def _u01(seed, k):
    """deterministic uniform(0,1) from (seed, k) -- splitmix64, matches nothing else"""
    x = (seed * 0x9E3779B97F4A7C15 + k * 0xD1B54A32D192ED03 + 0x632BE59BD9B4E019) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    return (x >> 11) / float(1 << 53)


def params(mode="hrit", channel=0, noise=True, esn0_db=12.0, ramp_len=None, n=None,
           carrier_hz=None, amp=(0.05, 0.3)):
    """SURVEY.md section 8d signal: seed 0x5EED0000+channel, timing U(0,1) symbol, carrier
    within +-1 kHz, random phase, amplitude ramp 0.05 -> 0.3, AWGN at Es/N0."""
    fs, rs, alpha = MODES[mode]
    seed = 0x5EED0000 + channel
    p = SigParams()
    p.sample_rate, p.symbol_rate, p.rrc_alpha = fs, rs, alpha
    p.timing_offset = _u01(seed, 1)
    p.carrier_hz = (2.0 * _u01(seed, 2) - 1.0) * 1000.0 if carrier_hz is None else carrier_hz
    p.phase0 = 2.0 * np.pi * _u01(seed, 3)
    p.amp_start, p.amp_end = amp
    if ramp_len is None:
        ramp_len = n if n is not None else 1 << 20
    p.ramp_len = int(ramp_len)
    p.esn0_db = esn0_db
    p.noise = 1 if noise else 0
    p.seed = seed
    return p


def generate(p, n, start=0, out=None):
    """n complex samples for absolute indices [start, start+n) as complex64"""
    if out is None:
        out = np.empty(n, np.complex64)
    assert out.dtype == np.complex64 and out.flags.c_contiguous and len(out) >= n
    _lib().xrd_siggen_cf32(C.byref(p), start, n, out.ctypes.data_as(C.c_void_p))
    return out[:n]


def bits(seed, k_start, n):
    out = np.empty(n, np.int8)
    _lib().xrd_siggen_bits(seed, k_start, n, out.ctypes.data_as(C.c_void_p))
    return out


def to_s16(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty(2 * len(x), np.int16)
    _lib().xrd_cf32_to_s16(x.ctypes.data_as(C.c_void_p), len(x), out.ctypes.data_as(C.c_void_p))
    return out


def set_threads(n):
    """worker threads of generate(); launchers like torchrun export OMP_NUM_THREADS=1"""
    _lib().xrd_siggen_set_threads(int(n))


def to_u8(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty(2 * len(x), np.uint8)
    _lib().xrd_cf32_to_u8(x.ctypes.data_as(C.c_void_p), len(x), out.ctypes.data_as(C.c_void_p))
    return out


def to_s8(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty(2 * len(x), np.int8)
    _lib().xrd_cf32_to_s8(x.ctypes.data_as(C.c_void_p), len(x), out.ctypes.data_as(C.c_void_p))
    return out
