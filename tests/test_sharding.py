"""Multi-GPU host logic on CPU: world_size-2 gloo run of the stream sharding + record gather
that bench.py uses (independent streams shard one per rank, no data-path
collective)."""
import os
import subprocess
import sys

from conftest import ROOT

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from xritdemod_b200 import shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
mine = shard.streams_of_rank(5, rank, world)
rec = shard.StreamRecord(rank=rank, n_streams=len(mine), n_samples=1000 * len(mine), n_symbols=370 * len(mine),
                         elapsed_ms=10.0 + rank, checksum=sum(mine))
shard.barrier()
allrec = shard.gather_records(rec)
if rank == 0:
    assert [r.rank for r in allrec] == [0, 1]
    assert sorted(sum((shard.streams_of_rank(5, r, 2) for r in range(2)), [])) == list(range(5))
    agg = shard.aggregate(allrec)
    assert agg["n_samples"] == 5000 and agg["n_symbols"] == 1850 and agg["elapsed_ms"] == 11.0, agg
    assert abs(agg["msps"] - 5000 / 11.0e-3 / 1e6) < 1e-9
    print("OK", agg)
dist.destroy_process_group()
"""


def test_two_rank_gloo_shard_and_gather(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER % dict(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES="")) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "OK" in outs[0][0]


def test_stream_assignment_is_a_partition():
    from xritdemod_b200 import shard

    for n, w in [(8, 8), (8, 4), (7, 2), (1, 8), (256, 8), (0, 2)]:
        parts = [shard.streams_of_rank(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert shard.seed_of_stream(3) == 0x5EED0000 + 3
