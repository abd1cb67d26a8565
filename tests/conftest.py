import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_ffi

    oracle_ffi.lib()
    return oracle_ffi


@pytest.fixture(scope="session")
def xrd():
    """the product binding; builds libxrd.so in-tree if it is stale (nvcc needs no GPU)"""
    from xritdemod_b200 import build

    build.build_all()
    from xritdemod_b200 import demod

    demod.lib()
    return demod


@pytest.fixture(scope="session")
def siggen():
    from xritdemod_b200 import siggen as sg

    return sg


@pytest.fixture(scope="session")
def gpu(xrd):
    rc, name, sms, cc = xrd.device_check(0)
    if rc != 0:
        pytest.fail("no usable sm_100 device (rc=%d, %s)" % (rc, name))
    return name


_SIG_CACHE = {}


def make_signal(mode, n, channel=0, noise=True, ramp=None, **kw):
    from xritdemod_b200 import siggen as sg

    key = (mode, n, channel, noise, ramp, tuple(sorted(kw.items())))
    if key not in _SIG_CACHE:
        p = sg.params(mode, channel, noise=noise, n=n, ramp_len=ramp if ramp is not None else min(n, 1 << 20), **kw)
        _SIG_CACHE[key] = (p, sg.generate(p, n))
    return _SIG_CACHE[key]


def assert_bitexact(a, b, what=""):
    a = np.ascontiguousarray(a).view(np.float32).reshape(-1)
    b = np.ascontiguousarray(b).view(np.float32).reshape(-1)
    assert len(a) == len(b), "%s: length %d != %d" % (what, len(a), len(b))
    ne = np.nonzero(a != b)[0]
    if len(ne):
        i = ne[0]
        raise AssertionError("%s: %d of %d floats differ; first at float %d (sample %d): %r vs %r; max|d|=%g" % (
            what, len(ne), len(a), i, i // 2, a[i], b[i], np.abs(a - b).max()))
