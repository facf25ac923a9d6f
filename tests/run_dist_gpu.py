"""Multi-GPU parity check (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/run_dist_gpu.py

Every rank projects + resamples its contiguous shard of one cloud with ShardedUniformProjection;
rank 0 also runs the single-GPU operator on the whole cloud and checks that the concatenation of
the shards' results equals it (positions within 1e-4 rel, masks equal up to tolerance-boundary
flips).  Prints one line 'DIST_PARITY_OK ...' on success."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isopoints_b200.dist import ShardedUniformProjection, all_gather_varlen, shard_range  # noqa: E402
from isopoints_b200.levelset_sampling import UniformProjection  # noqa: E402
from tests.helpers import Siren, SphereSDF, TinySiren  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for name, net, n in (("sphere", SphereSDF(), 40_000), ("siren", TinySiren(seed=4), 30_001),
                         ("fused_siren", Siren(256, 2, 30.0, seed=6), 20_003)):
        g = torch.Generator().manual_seed(11)
        x = ((torch.rand(1, n, 3, generator=g) - 0.5) * 1.6).to(dev)
        net = net.to(dev)
        kw = dict(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=2)
        b, e = shard_range(n, rank, world)
        out = ShardedUniformProjection(**kw).project_points(x[:, b:e].contiguous(), net, skip_upsampling=True)
        pts, _ = all_gather_varlen(out["levelset_points"][0])
        msk, _ = all_gather_varlen(out["mask"][0].float()[:, None])
        if rank == 0:
            ref = UniformProjection(**kw).project_points(x, net, skip_upsampling=True)
            same_shape = ref["levelset_points"].shape[1] == pts.shape[0]
            if same_shape:
                agree = (ref["mask"][0] == msk[:, 0].bool())
                close = torch.isclose(ref["levelset_points"][0], pts, rtol=1e-4, atol=1e-5).all(-1)
                frac = float((agree & close).float().mean())
            else:
                frac = 0.0
            print("%s: n=%d shards=%d agree=%.6f" % (name, n, world, frac), flush=True)
            ok = ok and frac > 0.999
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0 and ok:
        print("DIST_PARITY_OK world=%d" % world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
