"""world_size-2 gloo test (CPU) of the N>1 host logic: pair partition, max-over-ranks timing, and the
optional detection gather.  The data path itself has no collective."""
import importlib
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, ROOT)
    shard = importlib.import_module("multispectral-object-detection_b200.shard")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_bounds(n_pairs, world, rank)
    # each rank "computes" z for its pairs: row value = global pair index
    z_local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(hi - lo, 4, 3).contiguous()
    shards = shard.gather_detections(z_local, n_pairs)
    z = torch.cat(shards, 0)
    t = shard.max_over_ranks(10.0 + rank)
    # batched-NMS results: image g keeps (g % 3) + 1 boxes, every value = g
    det = torch.zeros(hi - lo, 5, 6)
    cnt = torch.zeros(hi - lo, dtype=torch.int32)
    for j, gidx in enumerate(range(lo, hi)):
        cnt[j] = (gidx % 3) + 1
        det[j, : int(cnt[j])] = float(gidx)
    per_image = shard.gather_nms(det, cnt, n_pairs)
    nms_ok = len(per_image) == n_pairs and all(
        d.shape == ((g % 3) + 1, 6) and bool((d == float(g)).all()) for g, d in enumerate(per_image))
    q.put((rank, lo, hi, z[:, 0, 0].tolist(), t, nms_ok))
    dist.destroy_process_group()


def test_shard_partition_and_collectives_world2():
    world, n_pairs = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 7)]
    for r in res:
        assert r[3] == [float(i) for i in range(n_pairs)]       # pair order preserved after the gather
        assert r[4] == 11.0                                      # max over ranks
        assert r[5]                                              # NMS results gathered per image, in pair order


def test_shard_bounds_cover_exactly():
    shard = importlib.import_module("multispectral-object-detection_b200.shard")
    for n in (0, 1, 7, 32, 255, 256):
        for w in (1, 2, 4, 8):
            b = [shard.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
