#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the full demodulator chain (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2|c4|c5]

Workloads (config.workload), all synthetic, one step = one pass of the whole chain (decimating FIR when
decimation > 1 -> AGC -> RRC FIR -> Costas -> M&M, demodulator.cpp:135-157) from the freshly constructed loop state:

  c2 (default; BASELINE.json configs[1], and configs[2] under torchrun): one HRIT BPSK stream, 927 ksym/s at
      2.5 Msps, 125 000 000 cf32 samples (1 GB) per GPU; rank r demodulates stream r (seed 0x5EED0000 + r)
  c4 (configs[3]): 10 Msps input, decimation 4 (241-tap Hamming LPF), 125 000 000 input samples, one JSON line per
      RRC tap count in {15, 31, 63, 127, 255}
  c5 (configs[4]): 256 LRIT channels x 4 194 304 samples in one call per GPU; rank r holds channels 256 r .. 256 r + 255

N > 1: one process per GPU (torchrun); no data-path collective, NCCL carries the barrier and the
result records.

  value : device-resident (input already in HBM, symbols left in HBM), whole job, one step at a time
  e2e   : same metric through xrd_demod_batch with pinned HOST buffers (H2D + D2H inside the timed region)
  --impl reference : the CPU restatement of the reference chain on the host cores, one thread per stream like the
          reference's symbolThread (demodulator.cpp:475), on the same config.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "IQ Msamples/s through full demod chain"
UNIT = "Msamples/s"
N_STREAM = 125_000_000
SPS = {"hrit": 2.5e6 / 927000.0, "lrit": 1.25e6 / 293883.0}
C4_TAPS = (15, 31, 63, 127, 255)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload(config, taps=63, n_override=None):
    """the named BASELINE.json config: signal mode, demodulator parameters, sizes, algorithmic bytes per input sample
    (B = 8 + 8 / (D * sps), cf32 in + cf32 symbols out)"""
    if config == "c2":
        w = dict(sig="hrit", mode="hrit", kw={}, n=N_STREAM, nch=1, D=1, sps=SPS["hrit"],
                 label="HRIT BPSK 927 ksym/s, 2.5 Msps, %d-sample cf32 stream per GPU (configs[1])")
    elif config == "c4":
        w = dict(sig="hrit10", mode="hrit", kw=dict(sample_rate=10000000, decimation=4, rrc_taps=taps), n=N_STREAM, nch=1,
                 D=4, sps=SPS["hrit"],
                 label="HRIT stress: 10 Msps in, decimation 4 (241-tap LPF), %d-tap RRC, %%d input samples per GPU (configs[3])" % taps)
    elif config == "c5":
        w = dict(sig="lrit", mode="lrit", kw={}, n=1 << 22, nch=256, D=1, sps=SPS["lrit"],
                 label="256 LRIT channels (293.883 ksym/s at 1.25 Msps) x %d samples in one call per GPU (configs[4])")
    else:
        raise SystemExit("unknown --config %s" % config)
    if n_override:
        w["n"] = int(n_override)
    w["config"] = config
    w["label"] = w["label"] % w["n"]
    w["bytes_per_sample"] = 8.0 + 8.0 / (w["D"] * w["sps"])
    return w


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def generate_input(w, rank, out):
    """fills out[nch, n] (complex64 view of a pinned tensor) with the rank's streams / channels"""
    from xritdemod_b200 import shard, siggen

    for c in range(w["nch"]):
        stream = rank * w["nch"] + c
        p = siggen.params(w["sig"], stream, n=w["n"], ramp_len=1 << 20)
        assert p.seed == shard.seed_of_stream(stream)
        siggen.generate(p, w["n"], out=out[c])


def dist_setup(n_gpus):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, local, world


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_oracle_msps(w, xs, n_threads):
    """the oracle chain over the rows of xs (one independent stream each), rows dealt round-robin to n_threads host
    threads, one chain per row like one reference process per stream; returns (Msps, seconds, symbols of row 0)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ffi as o

    cfg = o.config(w["mode"] == "hrit", **w["kw"])
    res = [0] * len(xs)

    def work(k):
        for r in range(k, len(xs), n_threads):
            res[r] = len(o.Chain(cfg).process(xs[r]))   # ctypes releases the GIL

    th = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
    t = time.perf_counter()
    [a.start() for a in th]
    [a.join() for a in th]
    dt = time.perf_counter() - t
    return sum(len(x) for x in xs) / dt / 1e6, dt, res[0]


def cpu_baseline_leg(w, x_host):
    """cpu_baseline of the b200 arm: the oracle on ONE host thread (the reference runs one stream on one symbolThread,
    demodulator.cpp:475), bounded sample of the same input"""
    rows = [x_host[c] for c in range(min(w["nch"], 32))]
    v, dt, _ = cpu_oracle_msps(w, rows, 1)
    return {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
            "sample": "%d of the same stream(s), %d samples each, one pass, 1 thread (%.1f s)" % (len(rows), w["n"], dt)}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's processSamples() on host cores, same config"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from xritdemod_b200 import siggen

    ncpu = os.cpu_count() or 1
    siggen.set_threads(ncpu)
    for taps in (C4_TAPS if args.config == "c4" else (63,)):
        w = workload(args.config, taps, args.samples)
        n_full = w["n"]
        if args.config == "c4":
            w["n"] = min(n_full, 31_250_000)   # bounded sample: a quarter of the stream per step, five tap counts
        gpus = max(1, args.gpus)
        streams = gpus * w["nch"]
        xs = np.empty((streams, w["n"]), np.complex64)
        for r in range(gpus):
            generate_input(w, r, xs[r * w["nch"]:(r + 1) * w["nch"]])
        # the reference runs the chain of one stream on one thread (symbolThread, demodulator.cpp:475): N streams
        # use min(N, host cores) threads, one reference process each
        cores = min(streams, ncpu)
        # one warm-up pass is all a CPU run needs (page faults, caches); the remaining warm-up steps of the contract
        # would only add 6.7 s each at the full 125 M-sample config
        warm_run = min(args.warmup, 1)
        for _ in range(warm_run):
            cpu_oracle_msps(w, xs, cores)
        t = 0.0
        for _ in range(args.steps):
            _, dt, nsym = cpu_oracle_msps(w, xs, cores)
            t += dt
        ms = t / args.steps * 1e3
        value = streams * w["n"] / (ms * 1e-3) / 1e6
        sample = "%d stream(s) x %d samples per step (%s), 1 thread per stream on %d host threads" % (
            streams, w["n"], "the whole workload" if w["n"] == n_full else "bounded: the first %d of %d samples" % (w["n"], n_full),
            cores)
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "warmup_run": warm_run, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["label"], "streams": streams, "samples_per_stream": w["n"],
                       "parallelism": "1 stream per host thread"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "host_cores": ncpu},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "parity unpinned: libSatHelper is not vendored by the reference, so its CPU path is timed through "
                    "the C restatement in oracle/ (kind=port)",
        }
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ the B200 arm
def run_b200(args):
    import torch

    from xritdemod_b200 import build, demod, shard, siggen

    rank, local, world = dist_setup(args.gpus)
    if world != args.gpus:
        log("note: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE" % (args.gpus, world))
    build.build_all()
    rc, name, sms, cc = demod.device_check(local)
    if rc != 0:
        raise SystemExit("bench.py: no usable sm_100 device on rank %d (rc=%d); there is no CPU fallback" % (rank, rc))
    torch.cuda.set_device(local)
    ncpu = os.cpu_count() or 1
    # launchers export OMP_NUM_THREADS=1; give every rank its share of the host cores for the signal source
    siggen.set_threads(max(1, ncpu // max(1, world)))
    for taps in (C4_TAPS if args.config == "c4" else (63,)):
        line = measure(args, workload(args.config, taps, args.samples), rank, local, world, name)
        if rank == 0:
            print(json.dumps(line), flush=True)


def measure(args, w, rank, local, world, devname):
    import torch

    from xritdemod_b200 import demod, shard, siggen

    n, nch = w["n"], w["nch"]
    lib = demod.lib()
    t0 = time.time()
    h_in = torch.empty(2 * n * nch, dtype=torch.float32).pin_memory()
    x = h_in.numpy().view(np.complex64).reshape(nch, n)
    generate_input(w, rank, x)
    log("[rank %d] %s, %s: generated %d x %d samples in %.1f s" % (rank, devname, w["config"], nch, n, time.time() - t0))

    d = demod.Demodulator(mode=w["mode"], device_ordinal=local, n_channels=nch, **w["kw"])
    cap = d.symbol_capacity(n)
    x_dev = torch.empty(2 * n * nch, dtype=torch.float32, device="cuda")
    x_dev.copy_(h_in)
    sym_dev = torch.empty(2 * cap * nch, dtype=torch.float32, device="cuda")
    h_sym = torch.empty(2 * cap * nch, dtype=torch.float32).pin_memory()
    stream = torch.cuda.ExternalStream(d.stream)
    torch.cuda.synchronize()
    i64p = C.POINTER(C.c_int64)

    def step_device():
        d.reset()
        return d.demod_device(x_dev.data_ptr(), n, sym_dev.data_ptr(), cap)

    def call_host(dd, fn, in_ptr, type_, out_ptr):
        dd.reset()
        cnt = np.zeros(nch, np.int64)
        rcode = fn(dd._h, C.c_void_p(in_ptr), n, type_, C.c_void_p(out_ptr), cap, cnt.ctypes.data_as(i64p))
        if rcode:
            raise demod.XrdError(rcode, lib.xrd_last_error(dd._h).decode())
        return cnt

    # ---- device-resident timing, one step at a time
    for _ in range(args.warmup):
        counts = step_device()
    shard.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    st0 = d.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {k: 0.0 for k in ("ms_fir_dec", "ms_agc", "ms_fir_rrc", "ms_costas", "ms_mm")}
    tw = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        counts = step_device()
        s = d.stats()
        for k in stage_ms:
            stage_ms[k] += s[k]
    e1.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - tw) * 1e3
    shard.barrier()
    clocks = sampler.result()
    dev_ms = e0.elapsed_time(e1)
    st1 = d.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    nsym = int(counts.sum())
    # checksum of the symbols (parity across runs / ranks): sum of the int8 soft symbols of every channel
    checksum = 0
    for c in range(nch):
        sc = sym_dev[2 * cap * c: 2 * cap * c + 2 * int(counts[c])].view(-1, 2)[:, 0]
        checksum = (checksum + int(torch.clamp(sc * 127, -128, 127).to(torch.int32).sum().item())) & 0xFFFFFFFF

    # ---- end to end through the public host-buffer API (xrd_demod_batch: pinned host input -> host symbols)
    call_host(d, lib.xrd_demod_batch, h_in.data_ptr(), 0, h_sym.data_ptr())
    torch.cuda.synchronize()
    shard.barrier()
    te = time.perf_counter()
    ke = max(1, min(args.steps, 3))
    for _ in range(ke):
        cnt_e = call_host(d, lib.xrd_demod_batch, h_in.data_ptr(), 0, h_sym.data_ptr())
    torch.cuda.synchronize()
    e2e_seq_ms = (time.perf_counter() - te) * 1e3 / ke
    assert np.array_equal(cnt_e, counts)

    # The same calls with several in flight: N demodulator handles driven by N host threads, consecutive steps in
    # flight together, so the PCIe copies of one step overlap the kernels of the others (and its kernels fill the SMs
    # the others' latency-bound certified re-run rounds leave idle).  Every step still copies its input in and its
    # symbols out inside the timed region and starts from the freshly constructed loop state; at least `steps` calls
    # are timed.
    extras = {}
    e2e_ms, in_flight, calls_timed = e2e_seq_ms, 1, ke
    if not args.no_overlap:
        gib = 8.0 * n * nch / 2 ** 30
        nf = max(2, args.in_flight if args.in_flight else (5 if world < 4 else 2))
        nf = max(1, min(nf, int(100 // max(gib * 5, 1e-9)) or 1))   # ~5 input-sized device buffers per handle, <= 100 GB
        more = [demod.Demodulator(mode=w["mode"], device_ordinal=local, n_channels=nch, **w["kw"]) for _ in range(nf - 1)]
        handles = [(d, h_sym)] + [(dx, torch.empty(2 * cap * nch, dtype=torch.float32).pin_memory()) for dx in more]
        per_handle = max(2, math.ceil(args.steps / nf))   # at least two calls per handle: the first ones all start together
        got = [[] for _ in handles]
        api = [lib.xrd_demod_batch, h_in.data_ptr(), 0]   # entry point, input buffer, sample type

        def work(i):
            dd, hs = handles[i]
            for _ in range(per_handle):
                got[i].append(call_host(dd, api[0], api[1], api[2], hs.data_ptr()))

        def run_all(fn):
            th = [threading.Thread(target=fn, args=(i,)) for i in range(nf)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            [a.start() for a in th]
            [a.join() for a in th]
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) * 1e3

        def timed(fn, expect):
            run_all(fn)                         # warm-up (allocations of the extra handles, this format)
            for g in got:
                g.clear()
            shard.barrier()
            tot = run_all(fn)
            done = sum(len(g) for g in got)
            if expect is not None:
                assert all(np.array_equal(c, expect) for g in got for c in g), "symbol counts differ between handles"
            return tot / done, done

        e2e_ms, calls_timed = timed(work, counts)
        in_flight = nf
        for _, hs in handles[1:]:
            assert torch.equal(h_sym[: 2 * int(counts[0])], hs[: 2 * int(counts[0])])
        if w["config"] == "c2":
            # the reference's own egress format: int8 soft symbols packed by the last kernel (xrd_demod_batch_i8,
            # SymbolManager.cpp:43-46), 1 byte per symbol back over PCIe instead of 8
            api[0] = lib.xrd_demod_batch_i8
            ms_i8, _ = timed(work, counts)
            soft = h_sym.view(torch.int8)[:nsym].to(torch.int32)
            assert (int(soft.sum().item()) & 0xFFFFFFFF) == checksum, "int8 egress checksum"
            extras["e2e_int8_egress"] = {
                "value": n * nch / ms_i8 / 1e3, "unit": UNIT, "ms_per_step": ms_i8, "h2d_bytes_per_step": 8 * n * nch,
                "d2h_bytes_per_step": nsym, "steps_in_flight": nf, "scope": "rank 0",
                "note": "xrd_demod_batch_i8: int8 soft symbols (the reference's wire format) packed by the last kernel"}
            # and the reference's usual ingest format as well (S16 IQ from the SDR front ends, demodulator.cpp:57-63):
            # the same stream quantised to int16, 4 bytes per sample in, 1 byte per symbol out
            h_in16 = torch.empty(2 * n, dtype=torch.int16).pin_memory()
            for c0 in range(0, 2 * n, 1 << 24):
                c1 = min(2 * n, c0 + (1 << 24))
                h_in16[c0:c1] = torch.clamp(torch.round(h_in[c0:c1] * 32768.0), -32768, 32767).to(torch.int16)
            api[1], api[2] = h_in16.data_ptr(), 1
            ms_s16, _ = timed(work, None)
            c16 = [c for g in got for c in g]
            assert all(np.array_equal(c, c16[0]) for c in c16) and abs(int(c16[0][0]) - nsym) <= 2
            extras["e2e_s16_in_i8_out"] = {
                "value": n / ms_s16 / 1e3, "unit": UNIT, "ms_per_step": ms_s16, "h2d_bytes_per_step": 4 * n,
                "d2h_bytes_per_step": nsym, "steps_in_flight": nf, "scope": "rank 0",
                "note": "xrd_demod_batch_i8 with XRD_S16IQ input: the reference's own ingest and egress formats"}
            del h_in16
        # device-resident with the same handles in flight (input already in HBM, symbols left in HBM): what the SMs
        # deliver when the latency-bound certified re-run rounds of one step are filled by the kernels of another
        sym_devs = [sym_dev] + [torch.empty(2 * cap * nch, dtype=torch.float32, device="cuda") for _ in more]

        def work_dev(i):
            dd = handles[i][0]
            for _ in range(per_handle):
                dd.reset()
                got[i].append(dd.demod_device(x_dev.data_ptr(), n, sym_devs[i].data_ptr(), cap))

        ms_dev, _ = timed(work_dev, counts)
        for sd in sym_devs[1:]:
            assert torch.equal(sym_dev[: 2 * int(counts[0])], sd[: 2 * int(counts[0])])
        extras["device_resident_in_flight"] = {
            "value": n * nch / ms_dev / 1e3, "unit": UNIT, "ms_per_step": ms_dev, "steps_in_flight": nf, "scope": "rank 0",
            "note": "xrd_demod_device on N handles together (wall clock, synchronised both sides); `value` is one step "
                    "at a time"}
        del sym_devs
        for dx in more:
            dx.close()

    # ---- what the host link delivers with nothing else going on: plain pinned H2D copies of the same input on every
    # rank at once (the ceiling of the e2e figure on this box)
    shard.barrier()
    torch.cuda.synchronize()
    tc = time.perf_counter()
    for _ in range(3):
        x_dev.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 3 * 8.0 * n * nch / (time.perf_counter() - tc) / 1e9
    # ... and with the step's output going the other way at the same time (two streams): one step's worth of plain
    # copies, nothing computed -- the floor of an e2e step on this box
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    tc = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s_up):
            x_dev.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_sym[: 2 * nsym].copy_(sym_dev[: 2 * nsym], non_blocking=True)
    torch.cuda.synchronize()
    copy_floor_ms = (time.perf_counter() - tc) * 1e3 / 3
    shard.barrier()

    if w["config"] == "c2" and not args.no_overlap:
        extras.update(seam_and_dropout_extras(d, x, n, lib))

    rec = shard.StreamRecord(rank=rank, n_streams=nch, n_samples=n * nch * args.steps, n_symbols=nsym * args.steps,
                             elapsed_ms=max(dev_ms, 0.0), checksum=checksum)
    recs = shard.gather_records(rec)
    rec_e = shard.gather_records(shard.StreamRecord(rank, nch, n * nch, nsym, e2e_ms, checksum))
    rec_h = shard.gather_records(shard.StreamRecord(rank, nch, 0, 0, h2d_gbs, 0))
    rec_f = shard.gather_records(shard.StreamRecord(rank, nch, 0, 0, copy_floor_ms, 0))
    if rank != 0:
        return None
    agg, agg_e = shard.aggregate(recs), shard.aggregate(rec_e)
    ms_per_step = agg["elapsed_ms"] / args.steps
    value = agg["msps"]

    # ---- roofline of the dominant kernel: device time of its stage measured live (CUDA events recorded on the
    # demodulator's own stream around the stage, xrd_get_stats), algorithmic bytes of SURVEY.md 8(d)
    peak, peak_src = peaks()
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    dom_ms = stage_ms[dom] / args.steps
    kernel_of = {"ms_mm": "mm_chain32_kernel<1024> + mm_delta_kernel<512>", "ms_costas": "wn_loop_kernel<CostasLoopK,4>",
                 "ms_agc": "wn_loop_kernel<AgcLoop,4>", "ms_fir_rrc": "fir_tma_kernel", "ms_fir_dec": "fird_poly_kernel<4>"}
    alg_bytes = n * nch * w["bytes_per_sample"]
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes of that kernel from the committed ncu --set full capture (profiles/), per stage pass
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("samples") == n * nch and tr.get("config", "c2") == w["config"]:
            traffic = tr.get(kernel_of[dom])
    except Exception:
        pass
    fir_ms = stage_ms["ms_fir_rrc"] / args.steps
    roofline = {"bound": "hbm", "kernel": kernel_of[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_sample": w["bytes_per_sample"], "kernel_ms": dom_ms,
                "note": "kernel_ms = all launches of that stage's kernels in one step (first pass + certified re-runs); "
                        "the feedback loops are latency/issue-bound exact recurrences, not HBM-bound (DESIGN.md)",
                "chain_achieved": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                "chain_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                # the one streaming stage, on its own stage traffic (8 B in + 8 B out per decimated sample)
                "rrc_fir": {"kernel_ms": fir_ms, "stage_gbs": 16.0 * n * nch / w["D"] / (fir_ms * 1e-3) / 1e9 if fir_ms > 0 else None,
                            "stage_frac": 16.0 * n * nch / w["D"] / (fir_ms * 1e-3) / 1e9 / peak if fir_ms > 0 else None,
                            "tflops": 4.0 * w["kw"].get("rrc_taps", 63) * n * nch / w["D"] / (fir_ms * 1e-3) / 1e12 if fir_ms > 0 else None},
                "stage_ms": {k: v / args.steps for k, v in stage_ms.items()}}

    cpu = None if args.no_cpu else cpu_baseline_leg(w, x)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": w["label"], "name": w["config"], "streams": world * nch, "samples_per_stream": n,
                   "symbols_per_gpu": nsym,
                   "l2": "input (%.2f GB) and every intermediate exceed the 126 MB L2; no flush needed" % (8 * n * nch / 1e9),
                   "parallelism": "%d stream(s) per GPU, no data-path collective" % nch},
        "e2e": {"value": agg_e["msps"], "unit": UNIT, "h2d_bytes_per_step": 8 * n * nch, "d2h_bytes_per_step": 8 * nsym,
                "ms_per_step": agg_e["elapsed_ms"], "steps_in_flight": in_flight, "calls_timed_per_rank": calls_timed,
                "one_step_at_a_time": {"value": n * nch / e2e_seq_ms / 1e3, "ms_per_step": e2e_seq_ms},
                "h2d_ceiling": {"gbs_per_rank": [r.elapsed_ms for r in rec_h], "gbs_total": sum(r.elapsed_ms for r in rec_h),
                                "msps_if_link_bound": sum(r.elapsed_ms for r in rec_h) / 8.0 * 1e3,
                                "copies_only_ms_per_step": max(r.elapsed_ms for r in rec_f),
                                "copies_only_ms_per_rank": [r.elapsed_ms for r in rec_f],
                                "copies_only_msps": world * n * nch / max(r.elapsed_ms for r in rec_f) / 1e3,
                                "note": "plain pinned-host -> device copies of the same input on all ranks at once; "
                                        "copies_only = one step's input up and symbols down at the same time on two streams, "
                                        "nothing computed, all ranks at once, slowest rank: what the host links of this box "
                                        "allow an e2e step (copies_only_msps is the ceiling of e2e.value)"},
                "note": "xrd_demod_batch on pinned host buffers; steps_in_flight = N: consecutive steps run on N demodulator "
                        "handles (N host threads), so the PCIe copies of one step overlap the kernels of the others; every "
                        "step's H2D and D2H are inside the timed region"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "wall_ms_per_step": wall_ms / args.steps,
        "symbol_checksum": checksum,
        "per_rank_ms_per_step": [r.elapsed_ms / args.steps for r in recs],
        "fixups": {k: st1[k] - st0[k] for k in ("agc_rounds", "costas_rounds", "mm_rounds", "agc_redo", "costas_redo", "mm_redo")},
    }
    line.update(extras)
    return line


def seam_and_dropout_extras(d, x, n, lib):
    """rank-local extras on the c2 stream: the FIFO seam the reference's threads use, and streams that lose the signal"""
    import torch

    from xritdemod_b200 import demod, siggen

    out = {}
    # (1) the seam north_star names: frontend callback -> FIFO -> processSamples (demodulator.cpp:54-74,100-168) with
    # CFileFrontend's 65535-sample callbacks (CFileFrontend.cpp:12,48), a frontend thread and a symbol-loop thread
    m = min(n, 32_000_000)
    xs = x[0][:m]
    d.reset()
    consumed, nsym_seam = [0], [0]
    done = threading.Event()

    def frontend():
        pos = 0
        while pos < m:
            k = min(65535, m - pos)
            try:
                d.add_samples(xs[pos:pos + k])
                pos += k
            except demod.XrdError:      # FIFO full (the reference would drop the block): wait for the symbol loop
                time.sleep(20e-6)
        done.set()

    def sink(ch, s):
        nsym_seam[0] += len(s)

    def symbol_loop():
        while consumed[0] < m:
            finishing = done.is_set()
            c = d.process(sink, min_samples=1 if finishing else 32768)
            consumed[0] += c
            if c == 0:
                time.sleep(1e-6)            # symbolLoopFunc, demodulator.cpp:170-175

    tf, tl = threading.Thread(target=frontend), threading.Thread(target=symbol_loop)
    t0 = time.perf_counter()
    tf.start()
    tl.start()
    tf.join()
    tl.join()
    dt = time.perf_counter() - t0
    out["e2e_fifo_seam"] = {
        "value": m / dt / 1e6, "unit": UNIT, "samples": m, "symbols": nsym_seam[0], "seconds": dt,
        "note": "xrd_add_samples (65535-sample callbacks, frontend thread) -> host FIFO of FIFO_SIZE floats -> "
                "xrd_process (symbol-loop thread) -> symbol callback; a call never holds more than 512 Ki samples "
                "(Parameters.h:57), so this is bound by per-call latency, not by the kernels"}
    # (2) loss of signal: speculation cannot merge across a stretch without a lockable signal, and the certified
    # re-runs then close one segment per round
    cap = d.symbol_capacity(n)
    sym = torch.empty(2 * cap, dtype=torch.float32, device="cuda")

    def rate(buf, reps=2):
        xd = torch.from_numpy(buf.view(np.float32)).cuda()
        best = None
        for _ in range(reps):
            d.reset()
            torch.cuda.synchronize()
            t = time.perf_counter()
            d.demod_device(xd.data_ptr(), len(buf), sym.data_ptr(), cap)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        return len(buf) / best / 1e6, best

    k = min(n, 32_000_000)
    burst = np.array(x[0][:k], copy=True)
    lo = k // 2
    burst[lo:lo + 1_000_000] = siggen.noise_only(7, 1_000_000)
    v, dt = rate(burst)
    out["dropout_stream"] = {"value": v, "unit": UNIT, "samples": k, "seconds": dt,
                             "note": "the c2 stream with 1 000 000 samples of noise only in the middle (signal lost for 0.4 s)"}
    v, dt = rate(siggen.noise_only(11, 4_000_000), reps=1)
    out["noise_only_stream"] = {"value": v, "unit": UNIT, "samples": 4_000_000, "seconds": dt,
                                "note": "no signal at all: every hand-off fails, the stages degrade to serial chains"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--samples", type=int, default=0, help="samples per stream (default: the config's size)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="e2e: one step at a time only; no extras")
    ap.add_argument("--in-flight", type=int, default=0,
                    help="e2e: demodulator handles (host threads) in flight together (default 5; 2 from 4 GPUs up)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
