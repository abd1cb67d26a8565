"""ctypes binding of the CPU oracle (oracle/libxrit_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under xritdemod_b200/ may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None


class XoConfig(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_uint32),
        ("symbol_rate", C.c_uint32),
        ("decimation", C.c_uint32),
        ("rrc_taps", C.c_uint32),
        ("loop_order", C.c_int32),
        ("rrc_alpha", C.c_float),
        ("pll_alpha", C.c_float),
        ("clock_alpha", C.c_float),
        ("clock_mu", C.c_float),
        ("clock_omega_limit", C.c_float),
        ("agc_rate", C.c_float),
        ("agc_ref", C.c_float),
        ("agc_gain", C.c_float),
        ("agc_max_gain", C.c_float),
    ]


class XoMmState(C.Structure):
    _fields_ = [
        ("mu", C.c_float),
        ("omega", C.c_float),
        ("p0", C.c_float * 2),
        ("p1", C.c_float * 2),
        ("p2", C.c_float * 2),
        ("c0", C.c_float * 2),
        ("c1", C.c_float * 2),
        ("c2", C.c_float * 2),
        ("next_index", C.c_int64),
    ]


def build(force=False):
    so = os.path.join(ORACLE_DIR, "libxrit_oracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("xrit_oracle.c", "xrit_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    fp = C.POINTER(C.c_float)
    vp = C.c_void_p
    L.xo_rrc_taps.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, vp]
    L.xo_rrc_taps.restype = C.c_int
    L.xo_lowpass_ntaps.argtypes = [C.c_double, C.c_double]
    L.xo_lowpass_ntaps.restype = C.c_int
    L.xo_lowpass_taps.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, vp]
    L.xo_lowpass_taps.restype = C.c_int
    L.xo_mmse_table.argtypes = [vp]
    L.xo_costas_gains.argtypes = [C.c_float, fp, fp]
    L.xo_sincosf.argtypes = [C.c_float, fp, fp]
    L.xo_sincosf_array.argtypes = [vp, C.c_int64, vp, vp]
    L.xo_set_libm_sincos.argtypes = [C.c_int]

    L.xo_fir_new.argtypes = [C.c_uint, vp, C.c_int]
    L.xo_fir_new.restype = vp
    L.xo_fir_free.argtypes = [vp]
    L.xo_fir_work.argtypes = [vp, vp, vp, C.c_int]

    L.xo_agc_new.argtypes = [C.c_float] * 4
    L.xo_agc_new.restype = vp
    L.xo_agc_free.argtypes = [vp]
    L.xo_agc_work.argtypes = [vp, vp, vp, C.c_int]
    L.xo_agc_gain.argtypes = [vp]
    L.xo_agc_gain.restype = C.c_float
    L.xo_agc_set_gain.argtypes = [vp, C.c_float]

    L.xo_costas_new.argtypes = [C.c_float, C.c_int]
    L.xo_costas_new.restype = vp
    L.xo_costas_free.argtypes = [vp]
    L.xo_costas_work.argtypes = [vp, vp, vp, C.c_int]
    L.xo_costas_get.argtypes = [vp, fp, fp]
    L.xo_costas_set.argtypes = [vp, C.c_float, C.c_float]

    L.xo_mm_new.argtypes = [C.c_float] * 5
    L.xo_mm_new.restype = vp
    L.xo_mm_free.argtypes = [vp]
    L.xo_mm_work.argtypes = [vp, vp, vp, C.c_int]
    L.xo_mm_work.restype = C.c_int
    L.xo_mm_trace.argtypes = [vp, vp, vp, vp, vp, C.c_int64]
    L.xo_mm_get.argtypes = [vp, C.POINTER(XoMmState)]
    L.xo_mm_set.argtypes = [vp, C.POINTER(XoMmState)]

    L.xo_config_defaults.argtypes = [C.POINTER(XoConfig), C.c_int]
    L.xo_chain_new.argtypes = [C.POINTER(XoConfig)]
    L.xo_chain_new.restype = vp
    L.xo_chain_free.argtypes = [vp]
    L.xo_chain_process.argtypes = [vp, vp, C.c_int64, vp, C.c_int64]
    L.xo_chain_process.restype = C.c_int64
    L.xo_chain_process_tap.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, vp, vp, vp, vp]
    L.xo_chain_process_tap.restype = C.c_int64
    L.xo_chain_sps.argtypes = [vp]
    L.xo_chain_sps.restype = C.c_float
    L.xo_soft_i8.argtypes = [vp, C.c_int64, vp]
    L.xo_convert_s16.argtypes = [vp, C.c_int64, vp]
    L.xo_convert_s8.argtypes = [vp, C.c_int64, vp]
    L.xo_diag_i8.argtypes = [vp, C.c_int64, vp]
    up = C.POINTER(C.c_uint)
    L.xo_conv_encode.argtypes = [vp, C.c_int64, up, vp]
    L.xo_nrzm_encode.argtypes = [vp, C.c_int64, C.POINTER(C.c_uint8), vp]
    L.xo_nrzm_decode_bytes.argtypes = [vp, C.c_int64]
    u32p = C.POINTER(C.c_uint32)
    L.xo_correlate.argtypes = [vp, C.c_uint32, vp, C.c_int, u32p, u32p, u32p]
    L.xo_fix_packet_180.argtypes = [vp, C.c_int64]
    L.xo_viterbi27_decode.argtypes = [vp, C.c_int, C.c_int, vp]
    L.xo_viterbi27_decode.restype = C.c_int
    L.xo_decoder_front.argtypes = [vp, C.c_int64, C.c_int, C.c_int, vp, vp, vp, C.c_int64, C.POINTER(C.c_int64)]
    L.xo_decoder_front.restype = C.c_int64
    L.xo_convert_u8.argtypes = [vp, C.c_int64, vp]
    L.xo_rtl_alpha.argtypes = [C.c_uint32]
    L.xo_rtl_alpha.restype = C.c_float
    L.xo_convert_rtl_u8.argtypes = [vp, C.c_int64, C.c_float, fp, vp]
    L.xo_set_fir_simd.argtypes = [C.c_int]
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def config(hrit=True, **kw):
    cfg = XoConfig()
    lib().xo_config_defaults(C.byref(cfg), 1 if hrit else 0)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def rrc_taps(fs, sym_rate, alpha, ntaps, gain=1.0):
    out = np.zeros(ntaps | 1, np.float32)
    n = lib().xo_rrc_taps(gain, fs, sym_rate, alpha, ntaps, _p(out))
    return out[:n]


def lowpass_taps(fs, cutoff, tw, gain=1.0):
    n = lib().xo_lowpass_ntaps(fs, tw)
    out = np.zeros(n, np.float32)
    lib().xo_lowpass_taps(gain, fs, cutoff, tw, _p(out))
    return out


def mmse_table():
    t = np.zeros((129, 8), np.float32)
    lib().xo_mmse_table(_p(t))
    return t


def costas_gains(bw):
    a, b = C.c_float(), C.c_float()
    lib().xo_costas_gains(bw, C.byref(a), C.byref(b))
    return a.value, b.value


def _cf(a):
    """view any complex64 / float32 array as a contiguous float32 IQ array"""
    a = np.ascontiguousarray(a)
    if a.dtype == np.complex64:
        a = a.view(np.float32)
    assert a.dtype == np.float32
    return a.reshape(-1)


class Fir:
    def __init__(self, decim, taps):
        taps = np.ascontiguousarray(taps, np.float32)
        self.decim = decim
        self.h = lib().xo_fir_new(decim, _p(taps), len(taps))

    def work(self, x, n_out=None):
        x = _cf(x)
        if n_out is None:
            n_out = (len(x) // 2) // self.decim
        out = np.empty(2 * n_out, np.float32)
        lib().xo_fir_work(self.h, _p(x), _p(out), n_out)
        return out.view(np.complex64)

    def __del__(self):
        lib().xo_fir_free(self.h)


class Agc:
    def __init__(self, rate=0.01, ref=0.5, gain=1.0, max_gain=4000.0):
        self.h = lib().xo_agc_new(rate, ref, gain, max_gain)

    def work(self, x):
        x = _cf(x)
        out = np.empty_like(x)
        lib().xo_agc_work(self.h, _p(x), _p(out), len(x) // 2)
        return out.view(np.complex64)

    @property
    def gain(self):
        return lib().xo_agc_gain(self.h)

    @gain.setter
    def gain(self, g):
        lib().xo_agc_set_gain(self.h, g)

    def __del__(self):
        lib().xo_agc_free(self.h)


class Costas:
    def __init__(self, bw=0.0037, order=2):
        self.h = lib().xo_costas_new(bw, order)

    def work(self, x):
        x = _cf(x)
        out = np.empty_like(x)
        lib().xo_costas_work(self.h, _p(x), _p(out), len(x) // 2)
        return out.view(np.complex64)

    @property
    def state(self):
        a, b = C.c_float(), C.c_float()
        lib().xo_costas_get(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    @state.setter
    def state(self, pf):
        lib().xo_costas_set(self.h, pf[0], pf[1])

    def __del__(self):
        lib().xo_costas_free(self.h)


class Mm:
    def __init__(self, omega, gain_omega, mu, gain_mu, omega_rel_limit):
        self.h = lib().xo_mm_new(omega, gain_omega, mu, gain_mu, omega_rel_limit)

    def work(self, x):
        x = _cf(x)
        n = len(x) // 2
        out = np.empty(2 * (n + 16), np.float32)
        ns = lib().xo_mm_work(self.h, _p(x), _p(out), n)
        return out[: 2 * ns].view(np.complex64).copy()

    def trace(self, cap):
        self.tr = dict(ii=np.zeros(cap, np.int64), mu=np.zeros(cap, np.float32),
                       omega=np.zeros(cap, np.float32), mm=np.zeros(cap, np.float32))
        t = self.tr
        lib().xo_mm_trace(self.h, _p(t["ii"]), _p(t["mu"]), _p(t["omega"]), _p(t["mm"]), cap)
        return t

    @property
    def state(self):
        st = XoMmState()
        lib().xo_mm_get(self.h, C.byref(st))
        return st

    @state.setter
    def state(self, st):
        lib().xo_mm_set(self.h, C.byref(st))

    def __del__(self):
        lib().xo_mm_free(self.h)


class Chain:
    """processSamples() restated: decimator -> AGC -> RRC -> Costas -> M&M."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.h = lib().xo_chain_new(C.byref(cfg))

    @property
    def sps(self):
        return lib().xo_chain_sps(self.h)

    def process(self, iq, taps=False):
        iq = _cf(iq)
        n = len(iq) // 2
        D = max(1, self.cfg.decimation)
        cap = int(n / D / max(1.5, self.sps * 0.9)) + 64
        sym = np.empty(2 * cap, np.float32)
        if not taps:
            ns = lib().xo_chain_process(self.h, _p(iq), n, _p(sym), cap)
            return sym[: 2 * ns].view(np.complex64).copy()
        m = n // D
        dec = np.empty(2 * m, np.float32) if D > 1 else None
        agc = np.empty(2 * m, np.float32)
        rrc = np.empty(2 * m, np.float32)
        cos = np.empty(2 * m, np.float32)
        ns = lib().xo_chain_process_tap(self.h, _p(iq), n, _p(sym), cap, _p(dec), _p(agc), _p(rrc), _p(cos))
        v = lambda a: None if a is None else a.view(np.complex64)
        return sym[: 2 * ns].view(np.complex64).copy(), dict(dec=v(dec), agc=v(agc), rrc=v(rrc), costas=v(cos))

    def __del__(self):
        lib().xo_chain_free(self.h)


def soft_i8(sym):
    sym = _cf(sym)
    out = np.empty(len(sym) // 2, np.int8)
    lib().xo_soft_i8(_p(sym), len(out), _p(out))
    return out


def convert_s16(x):
    x = np.ascontiguousarray(x, np.int16).reshape(-1)
    out = np.empty(len(x), np.float32)
    lib().xo_convert_s16(_p(x), len(x) // 2, _p(out))
    return out.view(np.complex64)


def convert_s8(x):
    x = np.ascontiguousarray(x, np.int8).reshape(-1)
    out = np.empty(len(x), np.float32)
    lib().xo_convert_s8(_p(x), len(x) // 2, _p(out))
    return out.view(np.complex64)


def diag_i8(floats):
    v = np.ascontiguousarray(floats, np.float32).reshape(-1)
    out = np.empty(len(v), np.int8)
    lib().xo_diag_i8(_p(v), len(v), _p(out))
    return out


def convert_u8(x):
    x = np.ascontiguousarray(x, np.uint8).reshape(-1)
    out = np.empty(len(x), np.float32)
    lib().xo_convert_u8(_p(x), len(x) // 2, _p(out))
    return out.view(np.complex64)


class RtlU8:
    """RtlFrontend's u8 conversion with its DC blocker; state carried across calls"""

    def __init__(self, sample_rate):
        self.alpha = lib().xo_rtl_alpha(sample_rate)
        self.avg = C.c_float(0.0)

    def convert(self, x):
        x = np.ascontiguousarray(x, np.uint8).reshape(-1)
        out = np.empty(len(x), np.float32)
        lib().xo_convert_rtl_u8(_p(x), len(x) // 2, self.alpha, C.byref(self.avg), _p(out))
        return out.view(np.complex64)


# ---- decoder front half ----
UW = {"lrit": (0xfca2b63db00d9794, 0x035d49c24ff2686b), "hrit": (0xfc4ef4fd0cc2df89, 0x25010b02f33d2076)}   # newdecoder.cpp:21-24


def conv_encode(bits, state=0):
    bits = np.ascontiguousarray(bits, np.uint8)
    out = np.empty(2 * len(bits), np.uint8)
    st = C.c_uint(state)
    lib().xo_conv_encode(_p(bits), len(bits), C.byref(st), _p(out))
    return out, st.value


def nrzm_encode(bits, last=0):
    bits = np.ascontiguousarray(bits, np.uint8)
    out = np.empty(len(bits), np.uint8)
    l = C.c_uint8(last)
    lib().xo_nrzm_encode(_p(bits), len(bits), C.byref(l), _p(out))
    return out, l.value


def correlate(data, words):
    d = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    w = np.array(words, np.uint64)
    a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib().xo_correlate(_p(d), len(d), _p(w), len(w), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def viterbi27(soft, n_bits, soft_mode=0):
    s = np.ascontiguousarray(soft).view(np.uint8).reshape(-1)
    out = np.zeros((n_bits + 7) // 8, np.uint8)
    ber = lib().xo_viterbi27_decode(_p(s), n_bits, soft_mode, _p(out))
    return out, ber


class DecoderFront:
    """xo_decoder_front with the 64 carried soft bytes"""

    def __init__(self, lrit=True, soft_mode=0):
        self.lrit, self.soft_mode = lrit, soft_mode
        self.last_end = np.full(64, 128, np.uint8)

    def run(self, soft):
        s = np.ascontiguousarray(soft).view(np.uint8).reshape(-1)
        cap = len(s) // 16384 + 1
        frames = np.zeros((cap, 1024), np.uint8)
        meta = np.zeros((cap, 4), np.int32)
        cons = C.c_int64()
        nf = lib().xo_decoder_front(_p(s), len(s), 1 if self.lrit else 0, self.soft_mode, _p(self.last_end), _p(frames),
                                    _p(meta), cap, C.byref(cons))
        return frames[:nf].copy(), meta[:nf].copy(), cons.value
