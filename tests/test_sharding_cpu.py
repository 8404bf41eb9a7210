"""CPU tests of the N>1 path with the gloo backend (world_size 2): slab gather == concatenation, max-over-ranks timing,
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
    B, cap = 3, 16
    g = torch.Generator().manual_seed(100 + rank)
    kps = torch.rand((B, cap, 7), generator=g)
    desc = torch.randint(0, 256, (B, cap, 32), generator=g, dtype=torch.uint8)
    n = torch.randint(0, cap, (B,), generator=g, dtype=torch.int32)
    gather = sharding.SlabGather(dist, world, kps, desc, n)
    gk, gd, gn = gather(kps, desc, n)
    t = sharding.max_over_ranks(dist, 1.0 + rank, torch.device('cpu'))
    q.put((rank, kps.numpy(), desc.numpy(), n.numpy(), gk.numpy(), gd.numpy(), gn.numpy(), t))
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
    cat_k = np.stack([r[1] for r in res]); cat_d = np.stack([r[2] for r in res]); cat_n = np.stack([r[3] for r in res])
    for r in res:
        assert r[4].tobytes() == cat_k.tobytes() and r[5].tobytes() == cat_d.tobytes() and r[6].tobytes() == cat_n.tobytes()
        assert r[7] == 2.0      # max over ranks


def test_shard_streams():
    from orb_slam3_modified_b200 import sharding
    for n, w in ((8, 8), (10, 4), (3, 8), (256, 2)):
        got = [sharding.shard_streams(n, r, w) for r in range(w)]
        assert sorted(sum(got, [])) == list(range(n))
        assert max(len(g) for g in got) - min(len(g) for g in got) <= (n + w - 1) // w
