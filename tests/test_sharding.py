"""Multi-GPU plumbing on CPU: independent transactions are sharded across ranks with no data-path collective; the only
communication is the barrier + max-over-ranks timing.  Exercised with world_size 2 over gloo."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
mine = bench.shard_batch(list(range(1024)), rank, world)
t = torch.tensor([float(len(mine)), 1.0 + rank])
total, tmax = bench.reduce_counts_and_time(len(mine), 1.0 + rank, dist, device="cpu")
gathered = [None] * world
dist.all_gather_object(gathered, mine)
if rank == 0:
    flat = sorted(x for part in gathered for x in part)
    print(json.dumps({"total": total, "tmax": tmax, "disjoint_cover": flat == list(range(1024)),
                      "types": [sum(1 for s in part if s %% 4 == k) for part in gathered for k in range(4)]}))
dist.destroy_process_group()
'''


def test_batch_sharding_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", CHILD % dict(root=ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["total"] == 1024 and res["tmax"] == 2.0 and res["disjoint_cover"]
    assert res["types"] == [128] * 8        # 256 transactions of each type, split evenly over the two ranks


def test_shard_batch_shapes():
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 4, 8):
        parts = [bench.shard_batch(list(range(1024)), r, world) for r in range(world)]
        assert sorted(x for p in parts for x in p) == list(range(1024))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


SPLIT_CHILD = r'''
import os, sys, json, random
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from oracle import refapi as Rf          # CPU stand-in for the per-GPU partial MSM (test infrastructure); the gather + host sum is the product's
from blockmaze_b200 import api
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
n = 3000
rng = random.Random(5)
bases = Rf.g1_bases_bytes(n, 4242)
scal = b"".join(rng.randrange(1 << 253).to_bytes(32, "little") for _ in range(n))
per = n // world
first, count = rank * per, (per if rank < world - 1 else n - per * (world - 1))      # the point-range split of bench.py msm_split
part = Rf.msm_g1_bytes(bases[64 * first:64 * (first + count)], scal[32 * first:32 * (first + count)], 1)[0]
gathered = [None] * world
dist.all_gather_object(gathered, part)
if rank == 0:
    total = api.g1_sum(gathered)
    print(json.dumps({"equal": total == Rf.msm_g1_bytes(bases, scal, 1)[0], "parts": len(gathered), "distinct": len(set(gathered))}))
dist.destroy_process_group()
'''


def test_msm_point_range_split_world_size_2():
    """SURVEY.md 8e, second half: a single MSM split by point range -- one partial point per rank, gathered without a data-path collective
    and summed on the host by zkb200_g1_sum.  World size 2 over gloo; the per-rank partial comes from the CPU reference here."""
    import json
    import pytest
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")):
        pytest.skip("oracle/_ref not built")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", SPLIT_CHILD % dict(root=ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res == {"equal": True, "parts": 2, "distinct": 2}
