"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from isopoints_b200.dist import all_gather_varlen, all_reduce_point_grads, shard_range, shard_views


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
    try:
        full = torch.arange(11 * 6, dtype=torch.float32).view(11, 6)
        b, e = shard_range(11, rank, world)
        got, counts = all_gather_varlen(full[b:e].clone())
        ok = torch.equal(got, full) and counts == [6, 5]
        # ragged with an empty shard
        part = full[:4] if rank == 0 else full[:0]
        got2, counts2 = all_gather_varlen(part.clone())
        ok = ok and torch.equal(got2, full[:4]) and counts2 == [4, 0]
        # equal shards take the no-copy path
        got3, _ = all_gather_varlen(full[rank * 5:(rank + 1) * 5].clone())
        ok = ok and torch.equal(got3, full[:10])
        g = torch.full((7, 3), float(rank + 1))
        all_reduce_point_grads(g)
        ok = ok and bool((g == 3.0).all())
        # the "nothing converged" early exit is a collective decision: a rank with an empty shard must not
        # leave alone (the others would wait for it in resample's all-gather)
        from isopoints_b200.dist import ShardedUniformProjection
        sp = ShardedUniformProjection()
        ok = ok and sp._nothing_converged([5] if rank == 0 else [0]) is False
        ok = ok and sp._nothing_converged([0]) is True
        try:
            sp.project_points(torch.zeros(1, 4, 3), None)          # upsampling branch is not sharded
            ok = False
        except NotImplementedError:
            pass
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_varlen_all_gather_and_grad_all_reduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 8, 200_000, 2_000_001):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    assert shard_views(16, 3, 8) == [6, 7] and shard_views(3, 3, 8) == []
