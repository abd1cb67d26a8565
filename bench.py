#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the full demodulator chain (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[1] -- one HRIT BPSK stream, 927 ksym/s at
2.5 Msps, 125 000 000 complex-float samples (1 GB) per GPU.  A step is one pass of the whole
chain (AGC -> RRC FIR -> Costas -> M&M; decimation 1, so the decimator is skipped exactly as the
reference does, demodulator.cpp:136) over that stream from the freshly constructed loop state.
N > 1: one process per GPU (torchrun), rank r demodulates its own stream (seed 0x5EED0000 + r);
there is no data-path collective, NCCL carries the barrier and the result records.

  value : device-resident (input already in HBM, symbols left in HBM), whole job
  e2e   : same metric through xrd_demod_batch with pinned HOST buffers (H2D + D2H inside)
  --impl reference : the CPU restatement of the reference chain (oracle/), single thread per
          stream like the reference's symbolThread (demodulator.cpp:475), bounded sample.
"""
import argparse
import ctypes as C
import json
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
SPS_HRIT = 2.5e6 / 927000.0
BYTES_PER_SAMPLE = 8.0 + 8.0 / SPS_HRIT          # cf32 in + cf32 symbols out = 10.966 B


def log(*a):
    print(*a, file=sys.stderr, flush=True)


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


def make_stream(stream_id, n):
    from xritdemod_b200 import shard, siggen

    p = siggen.params("hrit", stream_id, n=n, ramp_len=1 << 20)
    assert p.seed == shard.seed_of_stream(stream_id)
    return p


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


# ------------------------------------------------------------------------------------------
def cpu_oracle_msps(x, n_threads=1):
    """oracle chain over x (complex64), one thread per stream; returns (Msps, seconds)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ffi as o

    cfg = o.config(True)
    if n_threads == 1:
        ch = o.Chain(cfg)
        t = time.perf_counter()
        sym = ch.process(x)
        dt = time.perf_counter() - t
        return len(x) / dt / 1e6, dt, len(sym)
    res = [None] * n_threads
    chains = [o.Chain(cfg) for _ in range(n_threads)]

    def work(i):
        res[i] = len(chains[i].process(x))   # ctypes releases the GIL

    th = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    t = time.perf_counter()
    [a.start() for a in th]
    [a.join() for a in th]
    dt = time.perf_counter() - t
    return n_threads * len(x) / dt / 1e6, dt, res[0]


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's processSamples() on host cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from xritdemod_b200 import siggen

    n = 16_000_000   # bounded sample of the 125 M-sample stream (same generator, same seed)
    p = make_stream(0, N_STREAM)
    x = siggen.generate(p, n)
    ncpu = os.cpu_count() or 1
    # the arm's config is one stream per GPU; the reference runs the chain of one stream on one thread
    # (symbolThread, demodulator.cpp:475), so N streams use N host threads (one reference process each)
    streams = max(1, args.gpus)
    cores = min(streams, ncpu)
    for _ in range(args.warmup):
        cpu_oracle_msps(x, n_threads=cores)
    t = 0.0
    for _ in range(args.steps):
        _, dt, nsym = cpu_oracle_msps(x, n_threads=cores)
        t += dt
    ms = t / args.steps * 1e3
    value = cores * n / (ms * 1e-3) / 1e6
    allc, _, _ = cpu_oracle_msps(x[: 4_000_000], n_threads=ncpu)
    sample = "first %d samples of %d stream(s) per step, 1 thread per stream (the reference's symbolThread)" % (n, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "HRIT BPSK 927 ksym/s, 2.5 Msps, 125000000-sample cf32 stream (configs[1])",
                   "streams": cores, "samples_per_step": n * cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cores": ncpu, "all_cores_independent_streams_msps": allc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "parity unpinned: libSatHelper is not vendored by the reference, so its CPU path is timed through "
                "the C restatement in oracle/ (kind=port)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
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
    n = args.samples
    t0 = time.time()
    p = make_stream(rank, n)
    h_in = torch.empty(2 * n, dtype=torch.float32).pin_memory()
    x = h_in.numpy().view(np.complex64)
    siggen.generate(p, n, out=x)
    log("[rank %d] %s, generated %d samples in %.1f s" % (rank, name, n, time.time() - t0))

    d = demod.Demodulator(mode="hrit", device_ordinal=local)
    cap = d.symbol_capacity(n)
    x_dev = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    x_dev.copy_(h_in)
    sym_dev = torch.empty(2 * cap, dtype=torch.float32, device="cuda")
    h_sym = torch.empty(2 * cap, dtype=torch.float32).pin_memory()
    stream = torch.cuda.ExternalStream(d.stream)
    torch.cuda.synchronize()

    def step_device():
        d.reset()
        return int(d.demod_device(x_dev.data_ptr(), n, sym_dev.data_ptr(), cap)[0])

    def step_e2e():
        d.reset()
        cnt = np.zeros(1, np.int64)
        rcode = demod.lib().xrd_demod_batch(d._h, C.c_void_p(h_in.data_ptr()), n, 0, C.c_void_p(h_sym.data_ptr()), cap,
                                            cnt.ctypes.data_as(C.POINTER(C.c_int64)))
        if rcode:
            raise demod.XrdError(rcode, demod.lib().xrd_last_error(d._h).decode())
        return int(cnt[0])

    # ---- device-resident timing
    for _ in range(args.warmup):
        nsym = step_device()
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
        nsym = step_device()
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
    # the step includes host round trips (hand-off verdicts); the device-event span covers them
    elapsed = max(dev_ms, 0.0)
    # checksum of the symbols (parity across runs / ranks): sum of int8 soft symbols
    sym = sym_dev[: 2 * nsym].view(-1, 2)[:, 0]
    checksum = int(torch.clamp(sym * 127, -128, 127).to(torch.int32).sum().item()) & 0xFFFFFFFF

    # ---- end to end through the public host-buffer API (xrd_demod_batch: pinned host input -> host symbols)
    step_e2e()
    torch.cuda.synchronize()
    shard.barrier()
    te = time.perf_counter()
    ke = max(1, min(args.steps, 3))
    for _ in range(ke):
        nsym_e = step_e2e()
    torch.cuda.synchronize()
    e2e_seq_ms = (time.perf_counter() - te) * 1e3 / ke
    assert nsym_e == nsym

    # The same calls double-buffered: two demodulator handles, two host threads, consecutive steps in flight
    # together, so the PCIe copies of one step overlap the kernels of the other (and its kernels fill the SMs the
    # other's certified re-run rounds leave idle).  Every step still copies its 1 GB in and its symbols out inside
    # the timed region and starts from the freshly constructed loop state.
    e2e_ms, in_flight, e2e_i8_ms, dev_conc_ms, e2e_s16_ms = e2e_seq_ms, 1, None, None, None
    if not args.no_overlap:
        nf = max(2, args.in_flight)
        extra = [demod.Demodulator(mode="hrit", device_ordinal=local) for _ in range(nf - 1)]
        handles = [(d, h_sym)] + [(dx, torch.empty(2 * cap, dtype=torch.float32).pin_memory()) for dx in extra]
        per_handle = max(2, ke)
        counts = [[] for _ in handles]

        api = [demod.lib().xrd_demod_batch, h_in.data_ptr(), 0]   # entry point, input buffer, sample type

        def work(i):
            dd, hs = handles[i]
            cnt = np.zeros(1, np.int64)
            for _ in range(per_handle):
                dd.reset()
                rcode = api[0](dd._h, C.c_void_p(api[1]), n, api[2], C.c_void_p(hs.data_ptr()),
                               cap, cnt.ctypes.data_as(C.POINTER(C.c_int64)))
                counts[i].append((rcode, int(cnt[0])))

        def run_pair():
            th = [threading.Thread(target=work, args=(i,)) for i in range(nf)]
            t0 = time.perf_counter()
            [a.start() for a in th]
            [a.join() for a in th]
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) * 1e3

        run_pair()                      # warm-up (allocations of the extra handles)
        counts = [[] for _ in handles]
        shard.barrier()
        tot_ms = run_pair()
        done = sum(len(c) for c in counts)
        assert all(rc == 0 and ns == nsym for c in counts for rc, ns in c), counts
        for _, hs in handles[1:]:
            assert torch.equal(h_sym[: 2 * nsym], hs[: 2 * nsym])
        e2e_ms, in_flight = tot_ms / done, nf
        # the same with the reference's own egress format: int8 soft symbols packed by the last kernel
        # (xrd_demod_batch_i8, SymbolManager.cpp:43-46), 1 byte per symbol back over PCIe instead of 8
        api[0] = demod.lib().xrd_demod_batch_i8
        counts = [[] for _ in handles]
        run_pair()
        counts = [[] for _ in handles]
        shard.barrier()
        tot_i8 = run_pair()
        done_i8 = sum(len(c) for c in counts)
        assert all(rc == 0 and ns == nsym for c in counts for rc, ns in c), counts
        soft = h_sym.view(torch.int8)[:nsym].to(torch.int32)
        assert (int(soft.sum().item()) & 0xFFFFFFFF) == checksum, "int8 egress checksum"
        e2e_i8_ms = tot_i8 / done_i8
        # and with the reference's usual ingest format as well (S16 IQ from the SDR front ends, demodulator.cpp:57-63):
        # the same stream quantised to int16, 4 bytes per sample in, 1 byte per symbol out
        h_in16 = torch.empty(2 * n, dtype=torch.int16).pin_memory()
        for c0 in range(0, 2 * n, 1 << 24):
            c1 = min(2 * n, c0 + (1 << 24))
            h_in16[c0:c1] = torch.clamp(torch.round(h_in[c0:c1] * 32768.0), -32768, 32767).to(torch.int16)
        api[1], api[2] = h_in16.data_ptr(), 1
        counts = [[] for _ in handles]
        run_pair()
        counts = [[] for _ in handles]
        shard.barrier()
        tot_s16 = run_pair()
        done_s16 = sum(len(c) for c in counts)
        nsym_s16 = counts[0][0][1]
        assert all(rc == 0 and ns == nsym_s16 for c in counts for rc, ns in c) and abs(nsym_s16 - nsym) <= 2, counts
        e2e_s16_ms = tot_s16 / done_s16
        del h_in16
        # device-resident, the same handles in flight together (input already in HBM, symbols left in HBM): what the
        # SMs deliver when the latency-bound certified re-run rounds of one step are filled by the kernels of another
        sym_devs = [sym_dev] + [torch.empty(2 * cap, dtype=torch.float32, device="cuda") for _ in extra]

        def work_dev(i):
            dd = handles[i][0]
            for _ in range(per_handle):
                dd.reset()
                counts[i].append((0, int(dd.demod_device(x_dev.data_ptr(), n, sym_devs[i].data_ptr(), cap)[0])))

        def run_dev():
            th = [threading.Thread(target=work_dev, args=(i,)) for i in range(nf)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            [a.start() for a in th]
            [a.join() for a in th]
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) * 1e3

        counts = [[] for _ in handles]
        run_dev()
        counts = [[] for _ in handles]
        shard.barrier()
        tot_dev = run_dev()
        done_dev = sum(len(c) for c in counts)
        assert all(ns == nsym for c in counts for _, ns in c), counts
        for sd in sym_devs[1:]:
            assert torch.equal(sym_dev[: 2 * nsym], sd[: 2 * nsym])
        dev_conc_ms = tot_dev / done_dev
        del sym_devs
        for dx in extra:
            dx.close()

    rec = shard.StreamRecord(rank=rank, n_streams=1, n_samples=n * args.steps, n_symbols=nsym * args.steps,
                             elapsed_ms=elapsed, checksum=checksum)
    recs = shard.gather_records(rec)
    rec_e = shard.gather_records(shard.StreamRecord(rank, 1, n, nsym_e, e2e_ms, checksum))
    if rank != 0:
        return
    agg, agg_e = shard.aggregate(recs), shard.aggregate(rec_e)
    ms_per_step = agg["elapsed_ms"] / args.steps
    value = agg["msps"]

    # ---- roofline of the dominant kernel: device time of its stage measured live (CUDA events recorded on the
    # demodulator's own stream around the stage, xrd_get_stats), algorithmic bytes of SURVEY.md 8(d)
    peak, peak_src = peaks()
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    dom_ms = stage_ms[dom] / args.steps
    kernel_of = {"ms_mm": "mm_chain32_kernel<1024> + mm_delta_kernel<512>", "ms_costas": "wn_loop_kernel<CostasLoopK,4>",
                 "ms_agc": "wn_loop_kernel<AgcLoop,4>", "ms_fir_rrc": "fir1_kernel", "ms_fir_dec": "fird_kernel"}
    alg_bytes = n * BYTES_PER_SAMPLE
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes of that kernel from the committed ncu --set full capture (profiles/), per stage pass
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("samples") == n:
            traffic = tr.get(kernel_of[dom])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_of[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms,
                "note": "kernel_ms = all launches of that kernel in one step (first pass + certified re-runs); the "
                        "feedback loops are latency/issue-bound exact recurrences, not HBM-bound (DESIGN.md)",
                "chain_achieved": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                "chain_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                "stage_ms": {k: v / args.steps for k, v in stage_ms.items()}}

    # ---- CPU baseline beside it: the oracle on this box's host cores, bounded sample, 1 thread
    cpu = None
    if not args.no_cpu:
        ns = min(n, 64_000_000)
        v, dt, _ = cpu_oracle_msps(x[:ns])
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
               "sample": "first %d samples of the same stream, one pass, 1 thread (%.1f s)" % (ns, dt)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "HRIT BPSK 927 ksym/s, 2.5 Msps, %d-sample cf32 stream per GPU (configs[1])" % n,
                   "streams": world, "samples_per_stream": n, "symbols_per_stream": nsym,
                   "l2": "input (%.2f GB) and every intermediate exceed the 126 MB L2; no flush needed" % (8 * n / 1e9),
                   "parallelism": "1 stream per GPU, no data-path collective"},
        "e2e": {"value": agg_e["msps"], "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * nsym,
                "ms_per_step": agg_e["elapsed_ms"], "steps_in_flight": in_flight,
                "one_step_at_a_time": {"value": n / e2e_seq_ms / 1e3, "ms_per_step": e2e_seq_ms},
                "int8_egress": None if e2e_i8_ms is None else {
                    "value": n / e2e_i8_ms / 1e3, "ms_per_step": e2e_i8_ms, "d2h_bytes_per_step": nsym,
                    "note": "xrd_demod_batch_i8: int8 soft symbols (the reference's wire format) packed by the last kernel"},
                "note": "xrd_demod_batch on pinned host buffers; steps_in_flight = N: consecutive steps run on N demodulator "
                        "handles (N host threads), so the PCIe copies of one step overlap the kernels of the others; every "
                        "step's H2D and D2H are inside the timed region"},
        "e2e_s16_in_i8_out": None if e2e_s16_ms is None else {
            "value": n / e2e_s16_ms / 1e3, "unit": UNIT, "ms_per_step": e2e_s16_ms, "h2d_bytes_per_step": 4 * n,
            "d2h_bytes_per_step": nsym, "steps_in_flight": in_flight, "scope": "rank 0",
            "note": "xrd_demod_batch_i8 with XRD_S16IQ input: the reference's own ingest and egress formats"},
        "device_resident_in_flight": None if dev_conc_ms is None else {
            "value": n / dev_conc_ms / 1e3, "unit": UNIT, "ms_per_step": dev_conc_ms, "steps_in_flight": in_flight,
            "scope": "rank 0", "note": "xrd_demod_device on N handles together (wall clock, synchronised both sides); "
            "`value` above is one step at a time"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "wall_ms_per_step": wall_ms / args.steps,
        "symbol_checksum": checksum,
        "fixups": {k: st1[k] - st0[k] for k in ("agc_rounds", "costas_rounds", "mm_rounds", "agc_redo", "costas_redo", "mm_redo")},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=N_STREAM)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="e2e: one step at a time only (no double buffering)")
    ap.add_argument("--in-flight", type=int, default=4, help="e2e: demodulator handles (host threads) in flight together")
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
