"""Host logic of the multi-GPU path on CPU with the gloo backend, world_size 2: the flat gradient
bucket, the mean-of-views all-reduce and the view dealing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relightable3dgaussian_b200.dist import GradBucket, view_for_rank


def test_bucket_views_alias_flat_buffer_and_are_aligned():
    b = GradBucket(P=1000, S=5, M=16, device="cpu")
    assert set(b.views) == {"means3D", "features", "sh", "opacity", "scales", "rotations"}
    assert b.views["sh"].shape == (1000, 16, 3) and b.views["features"].shape == (1000, 5)
    for o in b.offsets:
        assert o % 128 == 0
    b.views["scales"].fill_(3.0)
    assert b.flat.sum().item() == 3.0 * 3000
    for v in b.views.values():
        assert v.is_contiguous()
        assert v.data_ptr() >= b.flat.data_ptr() and v.data_ptr() < b.flat.data_ptr() + b.bytes()
    # S = 0 (no feature section content) still works
    assert GradBucket(10, 0, 16, "cpu").views["features"].shape == (10, 0)


def test_view_dealing_covers_all_views():
    seen = [view_for_rank(s, r, 4, 8) for s in range(2) for r in range(4)]
    assert sorted(seen) == list(range(8))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = GradBucket(P=257, S=5, M=16, device="cpu")
    g = torch.Generator().manual_seed(100 + rank)
    for v in b.views.values():
        v.copy_(torch.randn(v.shape, generator=g))
    b.allreduce_mean()
    q.put((rank, {k: v.numpy().copy() for k, v in b.views.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_mean_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # expected: mean of the two ranks' seeded gradients, identical on both ranks
    exp = {}
    for rank in range(2):
        g = torch.Generator().manual_seed(100 + rank)
        b = GradBucket(P=257, S=5, M=16, device="cpu")
        for k, v in b.views.items():
            exp[k] = exp.get(k, 0) + torch.randn(v.shape, generator=g) / 2
    for k in exp:
        assert torch.allclose(torch.from_numpy(res[0][k]), exp[k], atol=1e-6) and (res[0][k] == res[1][k]).all()
