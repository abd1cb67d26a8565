"""include/xrd_sathelper.hpp -- the C++ host side that mirrors the reference's operator interface
(SatHelper::{Filters,FirFilter,AGC,CostasLoop,ClockRecovery}, demodulator.cpp:443-450,135-157) and
xrd::Demodulator (onSamplesAvailable / processSamples / SymbolManager::add seams).  The driver
(xritdemod_b200/csrc/shim_test.cpp) is compiled with plain g++ against libxrd.so."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import assert_bitexact, make_signal


@pytest.fixture(scope="module")
def shim(xrd):
    from xritdemod_b200 import build

    so = build.build_shim_test()
    assert so and os.path.exists(so)
    L = C.CDLL(so)
    L.shim_last_error.restype = C.c_char_p
    L.shim_run_operator_chain.restype = C.c_longlong
    L.shim_run_operator_chain.argtypes = [C.c_void_p, C.c_longlong, C.c_uint, C.c_uint, C.c_float, C.c_uint, C.c_int,
                                          C.c_void_p, C.c_longlong]
    L.shim_run_demodulator.restype = C.c_longlong
    L.shim_run_demodulator.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
    L.shim_run_wired.restype = C.c_longlong
    L.shim_run_wired.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_longlong, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong)]
    return L


def test_shim_compiles_and_loads_without_a_gpu(shim):
    for name in ("shim_run_operator_chain", "shim_run_demodulator", "shim_run_wired", "shim_error_paths", "shim_last_error"):
        assert hasattr(shim, name)


def _run_chain(shim, x, fs, rs, alpha, decim, chunk):
    x = np.ascontiguousarray(x)
    out = np.empty(2 * (len(x) // 2 + 64), np.float32)
    n = shim.shim_run_operator_chain(x.ctypes.data_as(C.c_void_p), len(x), fs, rs, alpha, decim, chunk,
                                     out.ctypes.data_as(C.c_void_p), len(out) // 2)
    assert n >= 0, shim.shim_last_error().decode()
    return out[: 2 * n].view(np.complex64)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["hrit", "lrit"])
def test_reference_call_sites_run_on_the_b200_path(gpu, shim, oracle, mode):
    """five operators constructed and called exactly as demodulator.cpp does, 65535*2-sample chunks"""
    _, x = make_signal(mode, 1 << 20)
    cfg = oracle.config(mode == "hrit")
    ref = oracle.Chain(cfg).process(x)
    got = _run_chain(shim, x, cfg.sample_rate, cfg.symbol_rate, cfg.rrc_alpha, 1, 131070)
    assert_bitexact(got, ref, "operator chain through the SatHelper shim")


@pytest.mark.gpu
def test_reference_call_sites_with_decimation(gpu, shim, oracle):
    _, x = make_signal("hrit10", 1 << 20)
    kw = dict(sample_rate=10000000, decimation=4)
    ref = oracle.Chain(oracle.config(True, **kw)).process(x)
    got = _run_chain(shim, x, 10000000, 927000, float(np.float32(0.3)), 4, 262144)
    assert_bitexact(got, ref, "decimated operator chain through the shim")


@pytest.mark.gpu
@pytest.mark.parametrize("type_", [0, 1])
def test_demodulator_seams(gpu, shim, oracle, siggen, type_):
    """frontend callback (65535-sample blocks, CFileFrontend.cpp:12) -> processSamples -> sink.add"""
    _, x = make_signal("hrit", 600000, amp=(0.3, 0.6))
    raw = x if type_ == 0 else siggen.to_s16(x)
    xf = x if type_ == 0 else oracle.convert_s16(raw)
    ref = oracle.Chain(oracle.config(True)).process(xf)
    raw = np.ascontiguousarray(raw)
    out = np.empty(2 * (len(x) // 2 + 64), np.float32)
    n = shim.shim_run_demodulator(raw.ctypes.data_as(C.c_void_p), len(x), type_, 1, 65535,
                                  out.ctypes.data_as(C.c_void_p), len(out) // 2)
    assert n >= 0, shim.shim_last_error().decode()
    assert_bitexact(out[: 2 * n].view(np.complex64), ref, "xrd::Demodulator seams")


def _run_wired(shim, x, block, threads, ckpt_block):
    x = np.ascontiguousarray(x)
    out = np.empty(2 * (len(x) // 2 + 64), np.float32)
    diag = np.empty(1024 * (len(x) // block + 2), np.float32)
    nd = C.c_longlong(0)
    n = shim.shim_run_wired(x.ctypes.data_as(C.c_void_p), len(x), 0, 1, block, threads, ckpt_block,
                            out.ctypes.data_as(C.c_void_p), len(out) // 2, diag.ctypes.data_as(C.c_void_p), len(diag),
                            C.byref(nd))
    assert n >= 0, shim.shim_last_error().decode()
    return out[: 2 * n].view(np.complex64), diag[: nd.value]


@pytest.mark.gpu
def test_diag_tap_and_checkpoint_resume(gpu, shim, oracle):
    """the DiagManager tap of demodulator.cpp:161-163 -- addSamples((float*)ba, min(symbols, 1024)) after every
    chunk -- and a checkpoint taken between two chunks, restored into a NEW demodulator that carries on: symbols and
    taps equal the oracle run in the same chunks"""
    _, x = make_signal("hrit", 600000)
    block = 65535
    ch = oracle.Chain(oracle.config(True))
    per_chunk = [ch.process(x[p:p + block]) for p in range(0, len(x), block)]
    ref = np.concatenate(per_chunk)
    ref_diag = np.concatenate([s.view(np.float32)[: min(len(s), 1024)] for s in per_chunk])
    got, diag = _run_wired(shim, x, block, 0, 4)
    assert_bitexact(got, ref, "symbols across a checkpoint / restore")
    assert_bitexact(diag, ref_diag, "diag tap")


@pytest.mark.gpu
def test_frontend_and_symbol_threads(gpu, shim, oracle):
    """frontend thread and symbol-loop thread as the reference wires them (demodulator.cpp:434,475): first calls race
    to create the device side, chunk sizes are whatever the FIFO holds; the symbols are the oracle's"""
    _, x = make_signal("hrit", 1 << 20)
    ref = oracle.Chain(oracle.config(True)).process(x)
    for _ in range(3):
        got, diag = _run_wired(shim, x, 65535, 1, 0)
        assert_bitexact(got, ref, "threaded seams")
        assert 0 < len(diag) <= 1024 * 40


@pytest.mark.gpu
def test_shim_error_behaviour(gpu, shim):
    assert shim.shim_error_paths() == 3   # both unsupported constructions threw SatHelperException
