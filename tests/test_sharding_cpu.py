"""CPU tests of the N>1 path with the gloo backend (world_size 2): packed per-group slab gather == the ranks' slabs, max-over-ranks timing,
stream sharding."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from orb_slam3_modified_b200 import sharding
    nb, cap, G = 3, 16, 2
    gen = torch.Generator().manual_seed(100 + rank)
    slabs = [[sharding.PackedSlab(nb, cap, torch.device('cpu')) for _ in range(2)] for _ in range(G)]
    for pair in slabs:
        for s in pair:
            s.kps.copy_(torch.rand((nb, cap, 7), generator=gen))
            s.desc.copy_(torch.randint(0, 256, (nb, cap, 32), generator=gen, dtype=torch.uint8))
            s.n.copy_(torch.randint(0, cap, (nb,), generator=gen, dtype=torch.int32))
            s.mono.copy_(torch.randint(0, cap, (nb,), generator=gen, dtype=torch.int32))
    gather = sharding.GroupSlabGather(dist, world, slabs)
    for d in range(2):                    # two rounds in flight per group, as in bench.py
        for g in range(G):
            gather.wait(g, d)
            gather.gather(g, d)
    gather.drain()
    # what this rank holds for (group 1, set 0) after the exchange, unpacked per source rank
    got = [tuple(v.numpy().copy() for v in sharding.PackedSlab.views_of(gather.out[1][0][r], nb, cap)) for r in range(world)]
    mine = tuple(v.numpy().copy() for v in (slabs[1][0].kps, slabs[1][0].desc, slabs[1][0].n, slabs[1][0].mono))
    t = sharding.max_over_ranks(dist, 1.0 + rank, torch.device('cpu'))
    q.put((rank, mine, got, t))
    dist.destroy_process_group()


def test_slab_gather_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        for src in range(world):          # every rank sees every rank's slab, field by field, bit for bit
            for a, b in zip(r[2][src], res[src][1]):
                assert a.dtype == b.dtype and a.tobytes() == b.tobytes()
        assert r[3] == 2.0      # max over ranks


def test_shard_streams():
    from orb_slam3_modified_b200 import sharding
    for n, w in ((8, 8), (10, 4), (3, 8), (256, 2)):
        got = [sharding.shard_streams(n, r, w) for r in range(w)]
        assert sorted(sum(got, [])) == list(range(n))
        assert max(len(g) for g in got) - min(len(g) for g in got) <= (n + w - 1) // w
