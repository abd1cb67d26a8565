"""Multi-GPU plumbing of the demodulator path: independent IQ streams shard across ranks
(one process per GPU), no collective on the data path.  torch.distributed is
used only for the start barrier and for gathering one small result record per rank."""
from dataclasses import asdict, dataclass

import torch
import torch.distributed as dist


@dataclass
class StreamRecord:
    rank: int
    n_streams: int
    n_samples: int      # complex input samples this rank demodulated in the timed region
    n_symbols: int
    elapsed_ms: float   # device time of the timed region on this rank
    checksum: int       # order-independent checksum of the soft symbols (parity across ranks/runs)


def streams_of_rank(n_streams, rank, world):
    """block partition of stream ids; sizes differ by at most one"""
    base, rem = divmod(n_streams, world)
    lo = rank * base + min(rank, rem)
    return list(range(lo, lo + base + (1 if rank < rem else 0)))


def seed_of_stream(stream):
    """synthetic-signal seed of a stream (0x5EED0000 + channel)"""
    return 0x5EED0000 + stream


def is_distributed():
    return dist.is_available() and dist.is_initialized()


def barrier():
    if is_distributed():
        dist.barrier()


def gather_records(rec):
    """every rank contributes one record; returns the rank-ordered list on every rank"""
    if not is_distributed():
        return [rec]
    t = torch.tensor([rec.rank, rec.n_streams, rec.n_samples, rec.n_symbols, rec.elapsed_ms, rec.checksum],
                     dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    recs = []
    for o in out:
        v = o.cpu().tolist()
        recs.append(StreamRecord(int(v[0]), int(v[1]), int(v[2]), int(v[3]), float(v[4]), int(v[5])))
    return sorted(recs, key=lambda r: r.rank)


def aggregate(recs):
    """whole-job numbers: units of all ranks over the slowest rank's time"""
    t = max(r.elapsed_ms for r in recs)
    n = sum(r.n_samples for r in recs)
    return dict(n_samples=n, n_symbols=sum(r.n_symbols for r in recs), n_streams=sum(r.n_streams for r in recs),
                elapsed_ms=t, msps=(n / (t * 1e-3) / 1e6) if t > 0 else 0.0, per_rank=[asdict(r) for r in recs])
