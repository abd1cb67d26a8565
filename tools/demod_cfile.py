"""Offline demodulation of a recorded IQ file on the B200 path: the reference's CFileFrontend use case
(demodulator/src/CFileFrontend.cpp:33-62 reads a file of complex<float> in 65535-sample blocks and paces them in
real time) without the pacing, with the reference's egress format: one int8 soft symbol per recovered symbol
(SymbolManager.cpp:37-52), the byte stream its decoder reads in 16384-byte frames (decoder/src/newdecoder.cpp:213-216).

usage: python tools/demod_cfile.py IN.cfile OUT.s8 [--mode hrit|lrit] [--sample-rate HZ] [--decimation D]
                                   [--type f32|s16|s8] [--chunk SAMPLES] [--cf32-out FILE]
The loop state is carried from chunk to chunk, so the output does not depend on --chunk."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xritdemod_b200 import demod  # noqa: E402

TYPES = {"f32": (demod.XRD_FLOATIQ, np.float32), "s16": (demod.XRD_S16IQ, np.int16), "s8": (demod.XRD_S8IQ, np.int8)}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("input")
    ap.add_argument("output", help="int8 soft symbols")
    ap.add_argument("--mode", default="hrit", choices=["hrit", "lrit"])
    ap.add_argument("--sample-rate", type=int, default=0, help="input sample rate (default: the mode's circuit rate)")
    ap.add_argument("--decimation", type=int, default=1)
    ap.add_argument("--type", default="f32", choices=sorted(TYPES))
    ap.add_argument("--chunk", type=int, default=32 << 20, help="complex samples per GPU call")
    ap.add_argument("--cf32-out", default=None, help="also write the complex soft symbols (cf32)")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)

    kw = dict(mode=args.mode, device_ordinal=args.device, decimation=args.decimation)
    if args.sample_rate:
        kw["sample_rate"] = args.sample_rate
    d = demod.Demodulator(**kw)
    xtype, npt = TYPES[args.type]
    chunk = max(args.decimation, args.chunk - args.chunk % args.decimation)
    n_in = n_sym = 0
    t0 = time.time()
    with open(args.input, "rb") as f, open(args.output, "wb") as out:
        cf = open(args.cf32_out, "wb") if args.cf32_out else None
        carry = np.empty(0, npt)
        while True:
            raw = np.fromfile(f, npt, 2 * chunk - len(carry))
            buf = np.concatenate([carry, raw]) if len(carry) else raw
            n = len(buf) // 2
            n -= n % args.decimation          # the remainder stays queued for the next call
            if n == 0:
                break
            x = buf[: 2 * n]
            carry = buf[2 * n:].copy()
            if cf is None:
                soft = d.demod_i8(x, type=xtype)
            else:
                sym = d.demod(x, type=xtype)
                sym.tofile(cf)
                soft = d.soft_i8(sym)
            soft.tofile(out)
            n_in += n
            n_sym += len(soft)
            if len(raw) < 2 * chunk - (len(buf) - len(raw)):
                break
        if cf is not None:
            cf.close()
    dt = time.time() - t0
    print("%d samples -> %d soft symbols in %.2f s (%.1f Msamples/s incl. file I/O)" % (n_in, n_sym, dt, n_in / dt / 1e6),
          file=sys.stderr)
    return n_in, n_sym


if __name__ == "__main__":
    main()
