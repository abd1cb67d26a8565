"""Python host-side binding of the C ABI in include/xrd.h (libxrd.so).

Class names and argument order mirror the reference's operator seam
(reference demodulator/src/demodulator.cpp:443-450, 135-157):

    FirFilter(decimation, taps).Work(x)            -> complex64 array
    AGC(rate, reference, gain, max_gain).Work(x)
    CostasLoop(loop_bw, order).Work(x)
    ClockRecovery(omega, gain_omega, mu, gain_mu, omega_rel_limit).Work(x) -> symbols
    Demodulator(config)                            -> processSamples() as one object

There is no CPU fallback: loading fails loudly when libxrd.so is missing, and every compute
call fails with XrdError when no sm_100 device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBXRD_PATH = os.environ.get("XRD_LIBXRD") or os.path.join(_HERE, "libxrd.so")   # (override: kernel experiments)

XRD_FLOATIQ, XRD_S16IQ, XRD_S8IQ, XRD_U8IQ, XRD_RTLU8IQ = 0, 1, 2, 3, 4
_NP_OF_TYPE = {XRD_FLOATIQ: np.float32, XRD_S16IQ: np.int16, XRD_S8IQ: np.int8, XRD_U8IQ: np.uint8, XRD_RTLU8IQ: np.uint8}

# every symbol include/xrd.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "xrd_config_defaults", "xrd_create", "xrd_destroy", "xrd_last_error", "xrd_add_samples", "xrd_process",
    "xrd_demod_batch", "xrd_demod_batch_i8", "xrd_demod_device", "xrd_soft_i8", "xrd_get_state", "xrd_set_state",
    "xrd_checkpoint_size", "xrd_checkpoint_save", "xrd_checkpoint_load", "xrd_symbol_capacity", "xrd_get_diag", "xrd_reset",
    "xrd_stream",
    "xrd_set_tuning", "xrd_get_stats",
    "xrd_design_rrc", "xrd_design_lowpass", "xrd_mmse_table", "xrd_costas_gains",
    "xrd_fir_create", "xrd_agc_create", "xrd_costas_create", "xrd_clock_recovery_create", "xrd_stage_work",
    "xrd_stage_set_tuning", "xrd_stage_set_loop_kernel", "xrd_stage_destroy", "xrd_stage_last_error", "xrd_device_check", "xrd_version",
    "xrd_decoder_front_create", "xrd_decoder_front_destroy", "xrd_decoder_front_reset", "xrd_correlate", "xrd_decoder_front_run",
]


class XrdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("xrd error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_uint32), ("symbol_rate", C.c_uint32), ("decimation", C.c_uint32),
        ("rrc_taps", C.c_uint32), ("loop_order", C.c_int32), ("rrc_alpha", C.c_float),
        ("pll_alpha", C.c_float), ("clock_alpha", C.c_float), ("clock_mu", C.c_float),
        ("clock_omega_limit", C.c_float), ("agc_rate", C.c_float), ("agc_ref", C.c_float),
        ("agc_gain", C.c_float), ("agc_max_gain", C.c_float), ("device_ordinal", C.c_int32),
        ("n_channels", C.c_int32),
    ]


class LoopState(C.Structure):
    _fields_ = [
        ("agc_gain", C.c_float), ("costas_phase", C.c_float), ("costas_freq", C.c_float),
        ("mm_mu", C.c_float), ("mm_omega", C.c_float), ("mm_p0", C.c_float * 2), ("mm_p1", C.c_float * 2),
        ("mm_next", C.c_int64), ("n_in", C.c_uint64), ("n_sym", C.c_uint64),
    ]


class Diag(C.Structure):
    _fields_ = [
        ("n_frame", C.c_int32), ("frame", C.c_int8 * 1024), ("n_symbols", C.c_uint64), ("mean_abs_i", C.c_double),
        ("mean_sq_i", C.c_double), ("mean_sq_q", C.c_double), ("snr_db", C.c_float), ("lock", C.c_float),
    ]


class FrameMeta(C.Structure):
    _fields_ = [("offset", C.c_int64), ("correlation", C.c_int32), ("word", C.c_int32), ("bit_errors", C.c_int32),
                ("reserved", C.c_int32)]


class Tuning(C.Structure):
    _fields_ = [
        ("agc_seg", C.c_int32), ("agc_warm", C.c_int32), ("costas_seg", C.c_int32), ("costas_warm", C.c_int32),
        ("mm_seg", C.c_int64), ("mm_warm", C.c_int64), ("mm_lanes", C.c_int32), ("mm_kernel", C.c_int32),
        ("mm_rerun", C.c_int32), ("mm_walk_lanes", C.c_int32), ("loop_kernel", C.c_int32), ("rerun_kernel", C.c_int32),
        ("h2d_pieces", C.c_int32), ("h2d_piece_min_ki", C.c_int32), ("costas_chains_per_sm", C.c_int32),
        ("agc_chains_per_sm", C.c_int32), ("agc_kernel", C.c_int32), ("chase", C.c_int32),
        ("guided", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64), ("agc_rounds", C.c_uint64), ("costas_rounds", C.c_uint64),
        ("mm_rounds", C.c_uint64), ("agc_redo", C.c_uint64), ("costas_redo", C.c_uint64), ("mm_redo", C.c_uint64),
        ("mm_windows", C.c_uint64), ("mm_iters", C.c_uint64), ("agc_iters", C.c_uint64),
        ("costas_iters", C.c_uint64), ("ms_fir_dec", C.c_float), ("ms_agc", C.c_float),
        ("ms_fir_rrc", C.c_float), ("ms_costas", C.c_float), ("ms_mm", C.c_float), ("mm_bail", C.c_uint64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


SYMBOLS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int)

_LIB = None


def lib():
    """dlopen libxrd.so (built in-tree by xritdemod_b200.build); raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIBXRD_PATH):
        raise ImportError(
            "xritdemod_b200: %s not found -- run `python -m xritdemod_b200.build` (needs nvcc). "
            "There is no CPU fallback." % LIBXRD_PATH)
    L = C.CDLL(LIBXRD_PATH)
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    L.xrd_version.restype = C.c_char_p
    L.xrd_device_check.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int)]
    L.xrd_config_defaults.argtypes = [C.POINTER(Config), C.c_int]
    L.xrd_config_defaults.restype = None
    L.xrd_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.xrd_destroy.argtypes = [vp]
    L.xrd_destroy.restype = None
    L.xrd_last_error.argtypes = [vp]
    L.xrd_last_error.restype = C.c_char_p
    L.xrd_add_samples.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
    L.xrd_process.argtypes = [vp, C.c_int64, SYMBOLS_CB, vp]
    L.xrd_process.restype = C.c_int64
    L.xrd_demod_batch.argtypes = [vp, vp, C.c_size_t, C.c_int, vp, C.c_size_t, i64p]
    L.xrd_demod_batch_i8.argtypes = [vp, vp, C.c_size_t, C.c_int, vp, C.c_size_t, i64p]
    L.xrd_demod_device.argtypes = [vp, vp, C.c_size_t, C.c_int, vp, C.c_size_t, i64p]
    L.xrd_soft_i8.argtypes = [vp, vp, C.c_size_t, vp]
    L.xrd_get_state.argtypes = [vp, C.c_int, C.POINTER(LoopState)]
    L.xrd_set_state.argtypes = [vp, C.c_int, C.POINTER(LoopState)]
    L.xrd_checkpoint_size.argtypes = [vp]
    L.xrd_checkpoint_size.restype = C.c_size_t
    L.xrd_checkpoint_save.argtypes = [vp, vp, C.c_size_t]
    L.xrd_checkpoint_load.argtypes = [vp, vp, C.c_size_t]
    L.xrd_symbol_capacity.argtypes = [vp, C.c_size_t]
    L.xrd_symbol_capacity.restype = C.c_int64
    L.xrd_get_diag.argtypes = [vp, C.c_int, C.POINTER(Diag)]
    L.xrd_reset.argtypes = [vp]
    L.xrd_stream.argtypes = [vp]
    L.xrd_stream.restype = vp
    L.xrd_set_tuning.argtypes = [vp, C.POINTER(Tuning)]
    L.xrd_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.xrd_design_rrc.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, vp, C.c_int]
    L.xrd_design_lowpass.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, vp, C.c_int]
    L.xrd_mmse_table.argtypes = [vp]
    L.xrd_mmse_table.restype = None
    L.xrd_costas_gains.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.xrd_costas_gains.restype = None
    L.xrd_fir_create.argtypes = [C.c_int, C.c_uint, vp, C.c_int, C.POINTER(vp)]
    L.xrd_agc_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(vp)]
    L.xrd_costas_create.argtypes = [C.c_int, C.c_float, C.c_int, C.POINTER(vp)]
    L.xrd_clock_recovery_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                            C.POINTER(vp)]
    L.xrd_stage_work.argtypes = [vp, vp, vp, C.c_int]
    L.xrd_stage_set_tuning.argtypes = [vp, C.c_int64, C.c_int64]
    L.xrd_stage_set_loop_kernel.argtypes = [vp, C.c_int]
    L.xrd_stage_destroy.argtypes = [vp]
    L.xrd_stage_destroy.restype = None
    L.xrd_stage_last_error.argtypes = [vp]
    L.xrd_stage_last_error.restype = C.c_char_p
    L.xrd_decoder_front_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.xrd_decoder_front_destroy.argtypes = [vp]
    L.xrd_decoder_front_destroy.restype = None
    L.xrd_decoder_front_reset.argtypes = [vp]
    u32p = C.POINTER(C.c_uint32)
    L.xrd_correlate.argtypes = [vp, vp, C.c_uint32, u32p, u32p, u32p]
    L.xrd_decoder_front_run.argtypes = [vp, vp, C.c_size_t, vp, C.POINTER(FrameMeta), C.c_size_t, C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t)]
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _iq(x):
    """any complex64 / float32 array -> contiguous float32 interleaved view"""
    x = np.ascontiguousarray(x)
    if x.dtype == np.complex64:
        x = x.view(np.float32)
    if x.dtype != np.float32:
        raise TypeError("expected complex64 or interleaved float32 samples")
    return x.reshape(-1)


def device_check(device=0):
    name = C.create_string_buffer(256)
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    rc = lib().xrd_device_check(device, name, 256, C.byref(sm), C.byref(ma), C.byref(mi))
    return rc, name.value.decode(), sm.value, (ma.value, mi.value)


def default_config(mode="hrit", **kw):
    cfg = Config()
    lib().xrd_config_defaults(C.byref(cfg), 1 if mode == "hrit" else 0)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


# ---- tap designers: SatHelper::Filters ----
def rrc_taps(gain, sample_rate, symbol_rate, alpha, ntaps):
    out = np.zeros(ntaps | 1, np.float32)
    n = lib().xrd_design_rrc(gain, sample_rate, symbol_rate, alpha, ntaps, _p(out), len(out))
    if n < 0:
        raise XrdError(n, "xrd_design_rrc")
    return out[:n]


def lowpass_taps(gain, sample_rate, cutoff, transition_width):
    out = np.zeros(1, np.float32)
    n = lib().xrd_design_lowpass(gain, sample_rate, cutoff, transition_width, _p(out), 1)
    if n < -1:
        out = np.zeros(-n, np.float32)
        n = lib().xrd_design_lowpass(gain, sample_rate, cutoff, transition_width, _p(out), len(out))
    if n < 0:
        raise XrdError(n, "xrd_design_lowpass")
    return out[:n]


def mmse_table():
    t = np.zeros((129, 8), np.float32)
    lib().xrd_mmse_table(_p(t))
    return t


def costas_gains(bw):
    a, b = C.c_float(), C.c_float()
    lib().xrd_costas_gains(bw, C.byref(a), C.byref(b))
    return a.value, b.value


# ---- stage operators ----
class _Stage:
    def __init__(self):
        self._h = C.c_void_p()

    def _check(self, rc):
        if rc < 0:
            raise XrdError(rc, lib().xrd_stage_last_error(self._h).decode())
        return rc

    def set_tuning(self, seg=0, warm=0):
        self._check(lib().xrd_stage_set_tuning(self._h, seg, warm))

    def set_loop_kernel(self, kernel):
        self._check(lib().xrd_stage_set_loop_kernel(self._h, kernel))

    def close(self):
        if self._h:
            lib().xrd_stage_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _work(self, x, n_in_per_out=1, out_len=None):
        x = _iq(x)
        n_in = len(x) // 2
        length = n_in // n_in_per_out
        out = np.empty(2 * (out_len if out_len is not None else length), np.float32)
        rc = self._check(lib().xrd_stage_work(self._h, _p(x), _p(out), length))
        return out, rc


class FirFilter(_Stage):
    """SatHelper::FirFilter(decimation, taps) -- demodulator.cpp:446,450"""

    def __init__(self, decimation, taps, device=0):
        super().__init__()
        taps = np.ascontiguousarray(taps, np.float32)
        self.decimation = max(1, int(decimation))
        rc = lib().xrd_fir_create(device, self.decimation, _p(taps), len(taps), C.byref(self._h))
        if rc:
            raise XrdError(rc, lib().xrd_stage_last_error(None).decode())

    def Work(self, x):
        out, _ = self._work(x, self.decimation)
        return out.view(np.complex64)


class AGC(_Stage):
    """SatHelper::AGC(rate, reference, gain, maxGain) -- demodulator.cpp:447"""

    def __init__(self, rate=0.01, reference=0.5, gain=1.0, max_gain=4000.0, device=0):
        super().__init__()
        rc = lib().xrd_agc_create(device, rate, reference, gain, max_gain, C.byref(self._h))
        if rc:
            raise XrdError(rc, lib().xrd_stage_last_error(None).decode())

    def Work(self, x):
        out, _ = self._work(x)
        return out.view(np.complex64)


class CostasLoop(_Stage):
    """SatHelper::CostasLoop(loopBandwidth, order) -- demodulator.cpp:448"""

    def __init__(self, loop_bw=0.0037, order=2, device=0):
        super().__init__()
        rc = lib().xrd_costas_create(device, loop_bw, order, C.byref(self._h))
        if rc:
            raise XrdError(rc, lib().xrd_stage_last_error(None).decode())

    def Work(self, x):
        out, _ = self._work(x)
        return out.view(np.complex64)


class ClockRecovery(_Stage):
    """SatHelper::ClockRecovery(omega, gainOmega, mu, gainMu, omegaRelativeLimit) -- demodulator.cpp:449"""

    def __init__(self, omega, gain_omega, mu, gain_mu, omega_rel_limit, device=0):
        super().__init__()
        rc = lib().xrd_clock_recovery_create(device, omega, gain_omega, mu, gain_mu, omega_rel_limit,
                                             C.byref(self._h))
        if rc:
            raise XrdError(rc, lib().xrd_stage_last_error(None).decode())

    def Work(self, x):
        n = len(_iq(x)) // 2
        out, ns = self._work(x, 1, out_len=n + 32)
        return out[: 2 * ns].view(np.complex64).copy()


# ---- the chain ----
class Demodulator:
    """processSamples() of the reference (demodulator.cpp:100-168) as one object.

    add_samples(data, type, channel)  == onSamplesAvailable        (demodulator.cpp:54-74)
    process(sink)                     == processSamples            (sink == SymbolManager::add)
    demod(iq)                         one-shot over a host array, state carried across calls
    """

    def __init__(self, cfg=None, mode="hrit", **kw):
        self.cfg = cfg if cfg is not None else default_config(mode, **kw)
        self._h = C.c_void_p()
        rc = lib().xrd_create(C.byref(self.cfg), C.byref(self._h))
        if rc:
            raise XrdError(rc, lib().xrd_last_error(None).decode())
        self.n_channels = self.cfg.n_channels

    def _check(self, rc):
        if rc < 0:
            raise XrdError(rc, lib().xrd_last_error(self._h).decode())
        return rc

    def close(self):
        if self._h:
            lib().xrd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sps(self):
        D = max(1, self.cfg.decimation)
        return float(np.float32(np.float32(self.cfg.sample_rate) / np.float32(D)) / np.float32(self.cfg.symbol_rate))

    def symbol_capacity(self, n_complex):
        """symbols one call over n_complex samples per channel can produce at most (xrd_symbol_capacity)"""
        return int(self._check(lib().xrd_symbol_capacity(self._h, int(n_complex))))

    def set_tuning(self, **kw):
        t = Tuning()
        for k, v in kw.items():
            setattr(t, k, v)
        self._check(lib().xrd_set_tuning(self._h, C.byref(t)))

    def stats(self):
        s = Stats()
        self._check(lib().xrd_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def state(self, channel=0):
        st = LoopState()
        self._check(lib().xrd_get_state(self._h, channel, C.byref(st)))
        return st

    def diag(self, channel=0):
        """diagnostics of the last call (xrd_get_diag): DiagManager frame bytes, SNR / lock estimate"""
        g = Diag()
        self._check(lib().xrd_get_diag(self._h, channel, C.byref(g)))
        return g

    def set_state(self, st, channel=0):
        """loop variables only (xrd_set_state); checkpoint()/restore() also carry the filter histories"""
        self._check(lib().xrd_set_state(self._h, channel, C.byref(st)))

    def checkpoint(self):
        """bytes: loop variables, filter histories, M&M tail and totals of every channel (xrd_checkpoint_save)"""
        n = lib().xrd_checkpoint_size(self._h)
        buf = C.create_string_buffer(n)
        self._check(lib().xrd_checkpoint_save(self._h, buf, n))
        return buf.raw

    def restore(self, blob):
        self._check(lib().xrd_checkpoint_load(self._h, C.c_char_p(blob), len(blob)))

    def reset(self):
        self._check(lib().xrd_reset(self._h))

    @property
    def stream(self):
        """cudaStream_t (int) all work of this demodulator is issued on"""
        return lib().xrd_stream(self._h) or 0

    def demod(self, iq, type=XRD_FLOATIQ):
        """iq: [n_channels, n] (or [n] for one channel) samples of `type`; returns a list of
        complex64 symbol arrays (one per channel), or the array itself for one channel."""
        if type == XRD_FLOATIQ:
            a = _iq(iq)
        else:
            a = np.ascontiguousarray(iq, _NP_OF_TYPE[type]).reshape(-1)
        n = len(a) // 2 // self.n_channels
        cap = self.symbol_capacity(n)
        sym = np.empty((self.n_channels, 2 * cap), np.float32)
        cnt = np.zeros(self.n_channels, np.int64)
        self._check(lib().xrd_demod_batch(self._h, _p(a), n, type, _p(sym), cap, cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        outs = [sym[c, : 2 * cnt[c]].view(np.complex64).copy() for c in range(self.n_channels)]
        return outs[0] if self.n_channels == 1 else outs

    def demod_i8(self, iq, type=XRD_FLOATIQ):
        """as demod(), but returns the int8 soft symbols the reference puts on the wire (SymbolManager.cpp:43-46),
        packed by the last kernel of the chain"""
        if type == XRD_FLOATIQ:
            a = _iq(iq)
        else:
            a = np.ascontiguousarray(iq, _NP_OF_TYPE[type]).reshape(-1)
        n = len(a) // 2 // self.n_channels
        cap = self.symbol_capacity(n)
        soft = np.empty((self.n_channels, cap), np.int8)
        cnt = np.zeros(self.n_channels, np.int64)
        self._check(lib().xrd_demod_batch_i8(self._h, _p(a), n, type, _p(soft), cap, cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        outs = [soft[c, : cnt[c]].copy() for c in range(self.n_channels)]
        return outs[0] if self.n_channels == 1 else outs

    def demod_device(self, iq_ptr, n_complex, sym_ptr, cap, type=XRD_FLOATIQ):
        """device-resident call: raw device pointers (ints), returns per-channel symbol counts"""
        cnt = np.zeros(self.n_channels, np.int64)
        self._check(lib().xrd_demod_device(self._h, C.c_void_p(iq_ptr), n_complex, type, C.c_void_p(sym_ptr), cap,
                                           cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        return cnt

    def add_samples(self, data, type=XRD_FLOATIQ, channel=0):
        if type == XRD_FLOATIQ:
            a = _iq(data)
        else:
            a = np.ascontiguousarray(data, _NP_OF_TYPE[type]).reshape(-1)
        self._check(lib().xrd_add_samples(self._h, channel, _p(a), len(a) // 2, type))

    def process(self, sink=None, min_samples=32768):
        """sink(channel, symbols: complex64 array).  Returns samples consumed per channel."""
        def _cb(user, ch, ptr, n):
            if sink is not None:
                arr = np.ctypeslib.as_array(ptr, shape=(2 * n,)).view(np.complex64).copy() if n else np.empty(0, np.complex64)
                sink(ch, arr)
        cb = SYMBOLS_CB(_cb)
        return self._check(lib().xrd_process(self._h, min_samples, cb, None))

    def soft_i8(self, sym):
        s = _iq(sym)
        out = np.empty(len(s) // 2, np.int8)
        self._check(lib().xrd_soft_i8(self._h, _p(s), len(out), _p(out)))
        return out


class DecoderFront:
    """The decoder's first steps on the soft-symbol byte stream (decoder/src/newdecoder.cpp:212-290): sync-word
    correlation, frame alignment, 180-degree phase fix (LRIT), Viterbi r=1/2 k=7, NRZ-M (HRIT)."""

    def __init__(self, lrit=True, soft_mode=0, device=0):
        self._h = C.c_void_p()
        rc = lib().xrd_decoder_front_create(device, 1 if lrit else 0, soft_mode, C.byref(self._h))
        if rc:
            raise XrdError(rc, lib().xrd_last_error(None).decode())

    def _check(self, rc):
        if rc < 0:
            raise XrdError(rc, lib().xrd_last_error(None).decode())
        return rc

    def close(self):
        if self._h:
            lib().xrd_decoder_front_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._check(lib().xrd_decoder_front_reset(self._h))

    def correlate(self, data):
        """Correlator::correlate: (highest correlation, position, word)"""
        d = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(lib().xrd_correlate(self._h, _p(d), len(d), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def run(self, soft):
        """soft: int8 soft symbols; returns (frames [n, 1024] uint8, meta list of FrameMeta, consumed bytes)"""
        d = np.ascontiguousarray(soft).view(np.int8).reshape(-1)
        cap = len(d) // 16384 + 1
        frames = np.zeros((cap, 1024), np.uint8)
        meta = (FrameMeta * cap)()
        nf, cons = C.c_size_t(), C.c_size_t()
        self._check(lib().xrd_decoder_front_run(self._h, _p(d), len(d), _p(frames), meta, cap, C.byref(nf), C.byref(cons)))
        return frames[: nf.value].copy(), [meta[i] for i in range(nf.value)], cons.value
