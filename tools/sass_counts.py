#!/usr/bin/env python
"""Mnemonic counts per kernel and the TMA FIR's hot loops, from `cuobjdump -sass xritdemod_b200/libxrd.so`.

    python tools/sass_counts.py        # rewrites profiles/r02_sass_counts.md and profiles/r02_sass_fir_tma.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "xritdemod_b200", "libxrd.so")
COLS = ["FFMA2", "FFMA", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "LDCU", "DADD", "SHFL", "MUFU"]


def kernels(text):
    cur, body = None, []
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                yield cur, body
            cur, body = m.group(1), []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            body.append(line)
    if cur:
        yield cur, body


def mnemonic(line):
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    return m.group(1) if m else ""


def main():
    text = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    rows, fir = [], None
    for name, body in kernels(text):
        ops = [mnemonic(l) for l in body]
        cnt = {c: 0 for c in COLS}
        for op in ops:
            head = op.split(".")[0]
            if head in cnt:
                cnt[head] += 1
        rows.append((name, len(ops), cnt))
        if "fir_tma_kernel" in name:
            fir = body
    out = ["# SASS evidence, round 2 (`python tools/sass_counts.py`: cuobjdump -sass xritdemod_b200/libxrd.so; sm_100a only)", "",
           "Mnemonic counts per kernel (TMA load = UTMALDG, bulk store = UBLKCP, mbarrier = SYNCS, packed FP32 = FFMA2,",
           "cp.async = LDGSTS, uniform constant load = LDCU):", "",
           "| kernel | instructions | " + " | ".join(COLS) + " |", "|---|---|" + "---|" * len(COLS)]
    for name, n, cnt in rows:
        out.append("| `%s` | %d | %s |" % (name[:70], n, " | ".join(str(cnt[c]) for c in COLS)))
    out.append("")
    open(os.path.join(ROOT, "profiles", "r02_sass_counts.md"), "w").write("\n".join(out))
    if fir:
        keep = [l.rstrip() for l in fir if re.search(r"UTMALDG|UBLKCP|SYNCS|FFMA2|LDCU|UTMACMDFLUSH|FENCE|LDS|STS|BRA|SHFL|ELECT|ARRIVE", l)]
        hdr = ["fir_tma_kernel: the instructions that carry the design (cuobjdump -sass, sm_100a), in program order.",
               "UTMALDG.2D = cp.async.bulk.tensor load by the producer warp; SYNCS = mbarrier arrive/try_wait; LDCU.64 + FFMA2 ... UR = tap pair",
               "from the constant bank used directly as a uniform operand of the packed FMA; UBLKCP = cp.async.bulk shared->global store of a",
               "warp's 2304-byte output slab.  %d instructions in the kernel, %d shown." % (len(fir), len(keep)), ""]
        open(os.path.join(ROOT, "profiles", "r02_sass_fir_tma.txt"), "w").write("\n".join(hdr + keep) + "\n")
    print("kernels:", len(rows))


if __name__ == "__main__":
    sys.exit(main())
