"""The reference arm of bench.py runs on host cores only, so its side of the JSON contract can be checked without a GPU:
one line per config (five for c4), the metric / unit / config of the B200 arm, `impl`, `cpu_baseline` and `e2e` as the
task statement spells them."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                       capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    return [json.loads(l) for l in r.stdout.strip().splitlines() if l.startswith("{")]


@pytest.mark.parametrize("config,lines", [("c2", 1), ("c4", 5)])
def test_reference_arm_prints_the_contract(config, lines):
    out = _run("--config", config, "--samples", "400000", "--steps", "2", "--warmup", "3")
    assert len(out) == lines
    for d in out:
        for k in REQUIRED:
            assert k in d, k
        assert d["impl"] == "reference" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
        assert d["metric"] == "IQ Msamples/s through full demod chain"
        assert d["value"] > 1.0 and d["ms_per_step"] > 0
        assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
        assert "workload" in d["config"] and "model" not in d["config"]
        cb = d["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
        e = d["e2e"]
        assert e["value"] == d["value"] and e["unit"] == d["unit"]
        assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
        assert d["steps"] == 2 and d["warmup"] == 3 and d["warmup_run"] == 1


def test_reference_arm_other_ranks_exit_quietly():
    """under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without output"""
    e = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--samples", "100000"],
                       capture_output=True, text=True, timeout=300, env=e)
    assert r.returncode == 0 and r.stdout.strip() == ""
