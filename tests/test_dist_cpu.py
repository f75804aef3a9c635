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


def _fake_raster(R, P, S, M, H, W, rank=0):
    """Stand-ins for the two C-ABI calls behind rasterizer._C (no GPU here): recognisable per-view values."""
    seen = {}

    def fake_fwd(*args, _defer=False, _min_capacity=None):
        z = lambda *s: torch.zeros(*s)
        seen["defer"] = _defer
        return (3, z(H, W).int(), z(3, H, W), z(1, H, W), z(1, H, W), z(S, H, W), z(3, H, W), z(3, H, W), z(P, 1), torch.ones(P).int(),
                torch.zeros(8, dtype=torch.uint8), torch.zeros(8, dtype=torch.uint8), torch.zeros(8, dtype=torch.uint8))

    def fake_bwd(*args, _out=None):
        seen["out"] = _out
        if _out is not None and "sh_factor" in _out:
            _out["sh_factor"].fill_(6.0 + rank)
        per_view = lambda shape, v: torch.full(shape, v + rank)
        # 9-tuple order of rasterize_points.cu:143-235
        return (per_view((P, 3), 10.0), per_view((P, 3), 11.0), per_view((P, 1), 3.0), per_view((P, 3), 1.0), per_view((P, S), 2.0),
                per_view((P, 6), 12.0), torch.empty(0) if _out is not None else per_view((P, M, 3), 8.0), per_view((P, 3), 4.0), per_view((P, 4), 5.0))
    return seen, fake_fwd, fake_bwd


def test_exchange_aware_backward_host_logic(monkeypatch):
    """rasterizer.set_grad_exchange: the autograd backward asks the kernels for the SH-gradient FACTOR only,
    gathers + rebuilds, and returns the averaged dL_dsh; every other gradient stays this view's (they are
    averaged at the leaves, dist.LeafGradBucket).  Host logic only: the C-ABI calls are stand-ins."""
    from relightable3dgaussian_b200 import rasterizer as R
    P, S, M, H, W = 7, 2, 16, 4, 5
    ex = FactoredGradExchange(P, S, M, "cpu", world=1)
    seen, fake_fwd, fake_bwd = _fake_raster(R, P, S, M, H, W)

    def fake_rebuild(means3D, campos_all, degree):
        seen["rebuild"] = (tuple(means3D.shape), tuple(campos_all.shape), degree)
        ex.sh.fill_(7.0)
        return ex.sh

    monkeypatch.setattr(R._C, "rasterize_gaussians", fake_fwd)
    monkeypatch.setattr(R._C, "rasterize_gaussians_backward", fake_bwd)
    monkeypatch.setattr(ex, "rebuild_sh", fake_rebuild)
    leaf = lambda *s: torch.randn(*s, requires_grad=True)
    means3D, means2D, feats, shs, opac, scales, rots = leaf(P, 3), leaf(P, 3), leaf(P, S), leaf(P, M, 3), leaf(P, 1), leaf(P, 3), leaf(P, 4)
    rs = R.GaussianRasterizationSettings(H, W, 1.0, 1.0, 0.0, 0.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3, torch.zeros(3),
                                         False, True, True, False)
    campos_all = torch.zeros(1, 3)
    R.set_grad_exchange(ex, campos_all)
    try:
        out = R.GaussianRasterizer(rs)(means3D=means3D, means2D=means2D, opacities=opac, shs=shs, scales=scales, rotations=rots, features=feats)
        (out[2].sum() + out[3].sum() + out[4].sum() + out[5].sum()).backward()
        assert set(seen["out"]) == {"sh_factor"} and seen["out"]["sh_factor"] is ex.factor and seen["rebuild"] == ((P, 3), (1, 3), 3)
        expect = {"means3D": (means3D, 1.0), "features": (feats, 2.0), "opacity": (opac, 3.0), "scales": (scales, 4.0), "rotations": (rots, 5.0),
                  "sh": (shs, 7.0), "means2D": (means2D, 10.0)}
        for name, (t, v) in expect.items():
            assert t.grad is not None and t.grad.shape == t.shape and bool((t.grad == v).all()), name
        assert shs.grad.data_ptr() != ex.sh.data_ptr()        # autograd must have copied: the buffer is reused next step
        # a backward that cannot take part (no shs) must raise, not silently skip the collectives
        out = R.GaussianRasterizer(rs)(means3D=means3D, means2D=means2D, opacities=opac, colors_precomp=leaf(P, 3), scales=scales, rotations=rots, features=feats)
        with pytest.raises(RuntimeError, match="same collectives"):
            out[2].sum().backward()
    finally:
        R.set_grad_exchange(None)
    # without an exchange installed the plain backward is used (no _out)
    out = R.GaussianRasterizer(rs)(means3D=means3D, means2D=means2D, opacities=opac, shs=shs, scales=scales, rotations=rots, features=feats)
    shs.grad = None
    out[2].sum().backward()
    assert seen["out"] is None and bool((shs.grad == 8.0).all())


def test_deferred_count_only_when_differentiated(monkeypatch):
    """set_deferred_count(True): a forward nobody will differentiate (no_grad / no input requires grad)
    must take the synchronous path — a deferred count there would never be resolved."""
    from relightable3dgaussian_b200 import rasterizer as R
    P, S, M, H, W = 5, 2, 16, 4, 5
    seen, fake_fwd, fake_bwd = _fake_raster(R, P, S, M, H, W)
    monkeypatch.setattr(R._C, "rasterize_gaussians", fake_fwd)
    rs = R.GaussianRasterizationSettings(H, W, 1.0, 1.0, 0.0, 0.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3, torch.zeros(3),
                                         False, True, True, False)
    t = lambda *s, g=False: torch.randn(*s, requires_grad=g)
    R.set_deferred_count(True)
    try:
        kw = lambda g: dict(means3D=t(P, 3, g=g), means2D=t(P, 3), opacities=t(P, 1), shs=t(P, M, 3), scales=t(P, 3), rotations=t(P, 4), features=t(P, S))
        R.GaussianRasterizer(rs)(**kw(True));  assert seen["defer"] is True
        R.GaussianRasterizer(rs)(**kw(False)); assert seen["defer"] is False
        with torch.no_grad():
            R.GaussianRasterizer(rs)(**kw(True)); assert seen["defer"] is False
    finally:
        R.set_deferred_count(False)


def _worker_leaf_exchange(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from relightable3dgaussian_b200 import rasterizer as R
    from relightable3dgaussian_b200.dist import LeafGradBucket
    P, S, M, H, W = 9, 3, 16, 4, 5
    ex = FactoredGradExchange(P, S, M, "cpu")
    seen, fake_fwd, fake_bwd = _fake_raster(R, P, S, M, H, W, rank=rank)
    R._C.rasterize_gaussians, R._C.rasterize_gaussians_backward = fake_fwd, fake_bwd

    def rebuild(means3D, campos_all, degree):          # stand-in for the CUDA rebuild: mean of the gathered factors
        ex.sh.copy_(ex.gathered.mean(0)[:, None, :].expand(P, M, 3))
        return ex.sh
    ex.rebuild_sh = rebuild
    g = torch.Generator().manual_seed(5)               # same replicated parameters on both ranks
    leaf = lambda *s: torch.randn(*s, generator=g).requires_grad_(True)
    xyz, theta, dc, rest, opac, scales, rots, env = leaf(P, 3), leaf(P, S), leaf(P, 1, 3), leaf(P, M - 1, 3), leaf(P, 1), leaf(P, 3), leaf(P, 4), leaf(4)
    view_scale = torch.tensor([2.0, -3.0])[rank]        # the view-dependent map leaves -> rasterizer inputs (render.py:91, neilf.py:120)
    feats = theta * view_scale + xyz[:, :1] * (rank + 1.0)
    rs = R.GaussianRasterizationSettings(H, W, 1.0, 1.0, 0.0, 0.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3, torch.zeros(3),
                                         False, True, True, False)
    bucket = LeafGradBucket([xyz, theta, opac, scales, rots, env])       # SH leaves are exchanged in backward
    R.set_grad_exchange(ex, torch.zeros(world, 3))
    bucket.zero()
    out = R.GaussianRasterizer(rs)(means3D=xyz, means2D=torch.zeros(P, 3, requires_grad=True), opacities=opac, shs=torch.cat([dc, rest], 1),
                                   scales=scales, rotations=rots, features=feats)
    loss = out[2].sum() + out[5].sum() + (env * (rank + 1.0)).sum()      # env: a gradient path that bypasses the rasterizer
    loss.backward()
    bucket.allreduce_mean()
    R.set_grad_exchange(None)
    q.put((rank, {k: v.grad.detach().numpy().copy() for k, v in dict(xyz=xyz, theta=theta, dc=dc, rest=rest, opac=opac, env=env).items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_leaf_level_exchange_with_view_dependent_features_world2_gloo():
    """averaging cotangents at the rasterizer INPUTS is wrong when the map from the leaves
    to those inputs depends on the view.  The exchange now averages the dense rest at the leaves
    (LeafGradBucket) and only dL_dsh inside backward: both ranks must end with mean_v(J_v^T g_v)."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_leaf_exchange, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    P, S, M = 9, 3, 16
    vs = [2.0, -3.0]
    # per-view cotangents of the stand-in backward: features 2+r, means3D 1+r, opacity 3+r, sh factor 6+r
    exp_theta = np.mean([vs[r] * (2.0 + r) for r in range(2)])                    # mean_v(J_v^T g_v), J_v = view_scale_v
    exp_xyz0 = np.mean([(1.0 + r) + S * (r + 1.0) * (2.0 + r) for r in range(2)])   # direct + through features
    exp_xyz12 = np.mean([1.0 + r for r in range(2)])
    for r in range(2):
        d = got[r]
        assert np.allclose(d["theta"], exp_theta) and np.allclose(d["xyz"][:, 0], exp_xyz0) and np.allclose(d["xyz"][:, 1:], exp_xyz12)
        assert np.allclose(d["opac"], 3.5) and np.allclose(d["env"], 1.5)
        assert np.allclose(d["dc"], 6.5) and np.allclose(d["rest"], 6.5)
    for k in got[0]:
        assert np.array_equal(got[0][k], got[1][k]), k                             # replicas stay identical


def _worker_bake(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from relightable3dgaussian_b200 import raytracer as RT

    class FakeTracer:                       # stand-in for the CUDA LBVH bake: visibility = a known function of (Gaussian, sample)
        traced = 0

        def __init__(self, *a):
            pass

        def bake_visibility(self, xyz, icov, opacity, normal, N, first_slot=0, count=None, write_dirs=True, **kw):
            P = xyz.shape[0]
            FakeTracer.traced += count
            assert not write_dirs                                  # sharded: directions are generated locally, not exchanged
            vis = torch.zeros(P, N, 1)
            order = torch.arange(P).flip(0)                        # "Morton order": slot s holds Gaussian P-1-s
            g = order[first_slot:first_slot + count]
            vis[g] = (xyz[g, :1] * 10.0)[:, None, :] + torch.arange(N, dtype=torch.float32)[None, :, None]
            return {"visibility": vis}
    RT.RayTracer = FakeTracer
    RT.sample_incident_rays = lambda n, is_training, N: (n[:, None, :].expand(-1, N, -1).contiguous(), torch.full((n.shape[0], N, 1), 6.0))
    P, N = 1003, 40                          # odd P: the last rank's range is shorter
    g = torch.Generator().manual_seed(9)
    xyz = torch.randn(P, 3, generator=g)
    nrm = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
    z = torch.zeros(P, 3)
    vis, dirs, areas = RT.update_visibility(xyz, z, z, z, z[:, 0], nrm, N, shard_group=True)
    q.put((rank, vis.numpy(), dirs.numpy(), FakeTracer.traced))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_visibility_bake_world2_gloo():
    """update_visibility(shard_group=...): each rank bakes only its range of leaf slots into a zero-initialised
    full-size tensor, one sum all-reduce restores the reference's [P,N,1] tensor on every rank (the bake kernel is
    replaced by a stand-in: no GPU here)."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_bake, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = {r: (v, d, n) for r, v, d, n in [q.get(timeout=120) for _ in range(2)]}
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    P, N = 1003, 40
    g = torch.Generator().manual_seed(9)
    xyz = torch.randn(P, 3, generator=g)
    nrm = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
    expect = ((xyz[:, :1] * 10.0)[:, None, :] + torch.arange(N, dtype=torch.float32)[None, :, None]).numpy()
    for r in range(2):
        v, d, traced = got[r]
        assert v.shape == (P, N, 1) and np.allclose(v, expect, atol=1e-6) and np.allclose(d, nrm[:, None, :].expand(-1, N, -1).numpy())
    assert got[0][2] == 502 and got[1][2] == 501          # ceil(1003 / 2) and the rest: nobody traced everything


def _worker_stats(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from relightable3dgaussian_b200.dist import allreduce_densification_stats
    g = torch.Generator().manual_seed(300 + rank)
    P = 101
    w, xg, ng = torch.rand(P, 1, generator=g), torch.rand(P, 1, generator=g), torch.rand(P, 1, generator=g)
    filt = torch.rand(P, generator=g) < 0.6
    radii = torch.randint(0, 50, (P,), generator=g, dtype=torch.int32)
    out = allreduce_densification_stats(w, xg, ng, filt, radii)
    q.put((rank, [t.numpy().copy() for t in out]))
    dist.barrier()
    dist.destroy_process_group()


def test_densification_stats_world2_gloo():
    """Two views of one step: the accumulator increments equal what the reference's per-view loop
    (gaussian_model.py:931-937, train.py:161-165) adds when it sees the two views one after the other."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_stats, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    P = 101
    acc = dict(w=torch.zeros(P, 1), x=torch.zeros(P, 1), n=torch.zeros(P, 1), d=torch.zeros(P, 1), r=torch.zeros(P, dtype=torch.int32))
    for rank in range(2):                                   # the reference's sequential accumulation
        g = torch.Generator().manual_seed(300 + rank)
        w, xg, ng = torch.rand(P, 1, generator=g), torch.rand(P, 1, generator=g), torch.rand(P, 1, generator=g)
        filt = torch.rand(P, generator=g) < 0.6
        radii = torch.randint(0, 50, (P,), generator=g, dtype=torch.int32)
        acc["w"] += w; acc["x"][filt] += xg[filt]; acc["n"][filt] += ng[filt]; acc["d"][filt] += 1
        acc["r"][filt] = torch.max(acc["r"][filt], radii[filt])
    for rank in range(2):
        w, x, n, d, r = got[rank]
        assert np.allclose(w, acc["w"].numpy(), atol=1e-6) and np.allclose(x, acc["x"].numpy(), atol=1e-6)
        assert np.allclose(n, acc["n"].numpy(), atol=1e-6) and np.array_equal(d, acc["d"].numpy()) and np.array_equal(r, acc["r"].numpy())
