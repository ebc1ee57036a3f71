"""world_size-2 gloo test of the sharding / all-gather host logic (CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pointdreamer_b200 import dist as pdist
    n = 6
    idx = pdist.shard_indices(n)
    local = torch.stack([torch.full((2, 3), float(i)) for i in idx])
    full = pdist.gather_stacked(local, n)
    ok = all(bool((full[i] == i).all()) for i in range(n)) and full.shape == (n, 2, 3)

    calls = []

    class FakeInpainter:
        def inpaint_batch(self, imgs, masks, chain0=None):
            calls.append((imgs.shape[0], chain0))
            off = torch.arange(imgs.shape[0], dtype=torch.float32).reshape(-1, 1, 1, 1) + chain0
            return imgs * 2 + off

    imgs = torch.arange(4 * 3 * 2 * 2, dtype=torch.float32).reshape(4, 3, 2, 2)
    out = pdist.inpaint_views_sharded(FakeInpainter(), imgs, torch.ones(4, 2, 2))
    expect = torch.stack([imgs[v] * 2 + v for v in range(4)])
    ok = ok and torch.equal(out, expect)
    ok = ok and calls == [(2, 2 * rank)]  # the local views run as ONE batch on their noise slots
    blocks = pdist.gather_blocks(torch.full((3, 2), float(rank)))
    ok = ok and torch.equal(blocks, torch.tensor([0., 0, 0, 1, 1, 1])[:, None].repeat(1, 2))
    q.put((rank, ok, idx))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    res.sort()
    assert res[0][1] and res[1][1]
    assert res[0][2] == [0, 2, 4] and res[1][2] == [1, 3, 5]
