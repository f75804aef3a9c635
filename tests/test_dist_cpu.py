"""Host logic of the multi-GPU path on CPU with the gloo backend, world_size 2: the flat gradient
bucket, the mean-of-views all-reduce and the view dealing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relightable3dgaussian_b200.dist import FactoredGradExchange, GradBucket, view_for_rank


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


def _worker_factored(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = FactoredGradExchange(P=257, S=5, M=16, device="cpu")
    g = torch.Generator().manual_seed(200 + rank)
    for v in ex.views.values():
        v.copy_(torch.randn(v.shape, generator=g))
    ex.dense.allreduce_mean()
    gathered = ex.gather_factors().clone()
    err = None
    try:                                   # the rebuild is a CUDA kernel: no CPU path, it must refuse
        ex.rebuild_sh(torch.zeros(257, 3), torch.zeros(world, 3), 3)
    except RuntimeError as e:
        err = str(e)
    q.put((rank, {k: v.numpy().copy() for k, v in ex.dense.views.items()}, gathered.numpy(), err))
    dist.barrier()
    dist.destroy_process_group()


def test_factored_exchange_world2_gloo():
    """Collective plumbing of the factorised exchange: the dense rest is mean-reduced, every rank ends
    up with all ranks' SH factors in rank order; the SH section is NOT part of the all-reduce."""
    ex = FactoredGradExchange(P=100, S=5, M=16, device="cpu", world=1)
    assert set(ex.views) == {"means3D", "features", "opacity", "scales", "rotations", "sh_factor"} and "sh" in ex.grads
    assert ex.bytes() < GradBucket(100, 5, 16, "cpu").bytes() / 3
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_factored, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    res = {r: (d, gth, err) for r, d, gth, err in got}
    exp_dense, exp_fac = {}, []
    for rank in range(2):
        g = torch.Generator().manual_seed(200 + rank)
        ex = FactoredGradExchange(P=257, S=5, M=16, device="cpu", world=1)
        for k, v in ex.views.items():
            t = torch.randn(v.shape, generator=g)
            if k == "sh_factor":
                exp_fac.append(t)
            else:
                exp_dense[k] = exp_dense.get(k, 0) + t / 2
    for rank in range(2):
        d, gth, err = res[rank]
        for k in exp_dense:
            assert torch.allclose(torch.from_numpy(d[k]), exp_dense[k], atol=1e-6)
        assert gth.shape == (2, 257, 3) and np.array_equal(gth[0], exp_fac[0].numpy()) and np.array_equal(gth[1], exp_fac[1].numpy())
        assert err is not None and "GPU only" in err


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
