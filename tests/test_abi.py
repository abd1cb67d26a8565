"""The C-ABI library loads without a GPU and exports every symbol include/xrd.h declares;
host-only entry points (parameter defaults, tap designers) agree with the oracle; compute entry
points fail loudly (no CPU fallback) when no device is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, assert_bitexact


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "xrd.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(xrd_[a-z0-9_]+)\s*\(", hdr)) - {"xrd_symbols_cb"})


def test_every_declared_symbol_is_exported(xrd):
    names = _declared_symbols()
    assert len(names) >= 26
    L = xrd.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(xrd.ABI_SYMBOLS) == names, "demod.ABI_SYMBOLS out of sync with include/xrd.h"


def test_library_is_sm100a_only(xrd):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", xrd.LIBXRD_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def test_config_defaults_follow_parameters_h(xrd):
    h = xrd.default_config("hrit")
    assert (h.sample_rate, h.symbol_rate, h.decimation, h.rrc_taps, h.loop_order) == (2500000, 927000, 1, 63, 2)
    assert abs(h.rrc_alpha - 0.3) < 1e-7 and abs(h.pll_alpha - 0.0037) < 1e-9   # demodulator.cpp:220
    assert (h.agc_rate, h.agc_ref, h.agc_gain, h.agc_max_gain) == (np.float32(0.01), 0.5, 1.0, 4000.0)
    l = xrd.default_config("lrit")
    assert (l.symbol_rate, l.sample_rate) == (293883, 1250000) and abs(l.rrc_alpha - 0.5) < 1e-7


def test_struct_layouts_match_the_oracle_config(xrd, oracle):
    # same leading fields in the same order: the oracle config is the product config minus device/channels
    a = [f for f, _ in xrd.Config._fields_][:14]
    b = [f for f, _ in oracle.XoConfig._fields_]
    assert a == b


@pytest.mark.parametrize("args", [(2.5e6, 927000.0, 0.3, 63), (1.25e6, 293883.0, 0.5, 63), (2.5e6, 927000.0, 0.3, 255),
                                  (4.0, 1.0, 0.5, 15)])
def test_designers_equal_oracle(xrd, oracle, args):
    assert_bitexact(xrd.rrc_taps(1, *args), oracle.rrc_taps(*args), "rrc taps")


def test_lowpass_mmse_gains_equal_oracle(xrd, oracle):
    assert_bitexact(xrd.lowpass_taps(1, 10e6, 1.25e6, 100e3), oracle.lowpass_taps(10e6, 1.25e6, 100e3), "lowpass")
    assert_bitexact(xrd.mmse_table(), oracle.mmse_table(), "mmse table")
    assert xrd.costas_gains(0.0037) == oracle.costas_gains(0.0037)


def test_bad_arguments_are_rejected(xrd):
    L = xrd.lib()
    assert L.xrd_create(None, None) == -1
    cfg = xrd.default_config("hrit", loop_order=4)
    h = C.c_void_p()
    assert L.xrd_create(C.byref(cfg), C.byref(h)) == -1 and not h
    assert b"loop_order" in L.xrd_last_error(None)
    assert L.xrd_demod_batch(None, None, 0, 0, None, 0, None) == -1
    assert L.xrd_stage_work(None, None, None, 4) == -1


def test_no_cpu_fallback_without_a_device(xrd):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(xrd.XrdError) as e:
        xrd.Demodulator(mode="hrit")
    assert e.value.code == -2  # XRD_E_CUDA
    with pytest.raises(xrd.XrdError):
        xrd.AGC()
    rc, _, _, _ = xrd.device_check(0)
    assert rc != 0


def test_product_does_not_touch_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/: the package, the public headers,
    the tools a user runs and every other function of bench.py must not"""
    import ast

    words = ("oracle_ffi", "xrit_oracle", "libxrit_oracle")
    pkg = os.path.join(ROOT, "xritdemod_b200")
    files = []
    for dp, _, fs in os.walk(pkg):
        files += [os.path.join(dp, f) for f in fs if f.endswith((".py", ".cu", ".cuh", ".c", ".cpp", ".h", ".hpp"))]
    inc = os.path.join(ROOT, "include")
    files += [os.path.join(inc, f) for f in os.listdir(inc)]
    files += [os.path.join(ROOT, "tools", f) for f in ("demod_cfile.py", "decode_frames.py") if
              os.path.exists(os.path.join(ROOT, "tools", f))]
    for f in files:
        s = open(f).read()
        assert not any(w in s for w in words), f
    # bench.py: the oracle may appear in the cpu_baseline / --impl reference legs only
    allowed = {"cpu_oracle_msps", "run_reference", "cpu_baseline_leg"}
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.Expr) and isinstance(getattr(node, "value", None), ast.Constant):
            continue   # the module docstring may describe the CPU legs
        seg = ast.get_source_segment(src, node) or ""
        if any(w in seg for w in words) or "oracle" in seg.replace("cpu_oracle_msps", ""):
            assert isinstance(node, ast.FunctionDef) and node.name in allowed, "bench.py: %s touches the oracle" % getattr(
                node, "name", type(node).__name__)
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    tree = ast.parse(entry)
    for node in tree.body:
        seg = ast.get_source_segment(entry, node) or ""
        if any(w in seg for w in words):
            assert isinstance(node, ast.FunctionDef) and node.name in ("smoke", "build"), "__graft_entry__: %s" % seg[:60]
