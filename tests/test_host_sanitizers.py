"""ASan + UBSan over the host-side C code (the oracle and the synthetic IQ source): tools/san/host_sanitize.c drives every
public entry point of the oracle, in ragged calls, and compares the stage operators with the chain bit for bit."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_oracle_and_siggen_under_asan_ubsan():
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    probe = subprocess.run(["gcc", "-fsanitize=address,undefined", "-x", "c", "-", "-o", "/dev/null"],
                           input="int main(void){return 0;}", capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("this gcc has no sanitizer runtime")
    r = subprocess.run([os.path.join(ROOT, "tools", "san", "run.sh")], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "host sanitizer run complete" in out
    assert "AddressSanitizer" not in out and "runtime error" not in out and "LeakSanitizer" not in out, out[-4000:]
    assert "bit-equal to the stage calls" in out
