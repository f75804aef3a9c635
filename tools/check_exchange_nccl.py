#!/usr/bin/env python3
"""Multi-GPU check of the factorised gradient exchange on real NCCL (tests/test_multigpu_gpu.py launches it
when the box has >= 2 GPUs).  Every rank renders its view of one step, then the SAME per-view gradients go
through both exchanges — dense all-reduce (dist.GradBucket) and factorised (dist.FactoredGradExchange)
— and the results are compared on every rank; both are timed with CUDA events (max over ranks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
         --master-port 29511 tools/check_exchange_nccl.py [--P 1000000] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as tdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relightable3dgaussian_b200 import _C_raster as C, dist as rdist, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1_000_000)
    ap.add_argument("--W", type=int, default=800)
    ap.add_argument("--H", type=int, default=800)
    ap.add_argument("--S", type=int, default=5)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    rank, local, world = rdist.init_from_env()
    assert world > 1, "launch with torchrun on >= 2 GPUs"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    P, W, H, S, M = a.P, a.W, a.H, a.S, 16
    sc = synth.make_scene(P, "shell-v1", 0, S)
    cams = [synth.make_camera(r, W, H) for r in range(world)]
    cam = cams[rank]
    d = lambda t: t.to(dev)
    E = torch.Tensor([])
    bg = torch.zeros(3, device=dev)
    means3D, scales, rots, opac, shs, feats = map(d, (sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs, sc.features))
    view, proj, pos = d(cam.viewmatrix), d(cam.projmatrix), d(cam.campos)
    g = torch.Generator().manual_seed(1234)
    cot = [torch.randn(c, H, W, generator=g).to(dev) for c in (3, 1, 1, S)]
    campos_all = torch.stack([d(c.campos) for c in cams]).contiguous()
    dense = rdist.GradBucket(P, S, M, dev)
    fact = rdist.FactoredGradExchange(P, S, M, dev)

    def fwd_bwd(out_views):
        o = C.rasterize_gaussians(bg, means3D, feats, E, opac, scales, rots, 1.0, E, view, proj, cam.tanfovx, cam.tanfovy,
                                  cam.cx, cam.cy, H, W, shs, 3, pos, False, True, False)
        C.rasterize_gaussians_backward(bg, means3D, feats, o[9], E, scales, rots, 1.0, E, view, proj, cam.tanfovx, cam.tanfovy,
                                       *cot, shs, 3, pos, o[10], o[0], o[11], o[12], True, False, _out=out_views)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); tdist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    # ---- same gradients through both exchanges -------------------------------------------------
    fwd_bwd(dense.views); dense.allreduce_mean()
    fwd_bwd(fact.views); fact.exchange(means3D, campos_all, 3)
    torch.cuda.synchronize()
    rel = {}
    for k in ("means3D", "features", "opacity", "scales", "rotations", "sh"):
        x, y = fact.grads[k].double(), dense.views[k].double()
        rel[k] = float((x - y).norm() / (y.norm() + 1e-30))
    # every rank must hold the same rebuilt SH gradient bit for bit
    chk = fact.sh.view(torch.int32).sum(dtype=torch.int64).reshape(1)
    lo, hi = chk.clone(), chk.clone()
    tdist.all_reduce(lo, op=tdist.ReduceOp.MIN); tdist.all_reduce(hi, op=tdist.ReduceOp.MAX)
    # ---- the same gradients through the ONE-kernel NVLink exchange (peer loads + NVLS all-reduce) ----------------------
    p2p_rel, p2p_info = None, "unavailable"
    try:
        p2p = rdist.P2PGradExchange(P, S, M, dev)
        p2p_info = dict(multicast=bool(p2p.multicast))
    except Exception as e:                                # no symmetric memory on this box / torch build
        p2p, p2p_info = None, f"unavailable: {type(e).__name__}: {e}"
    if p2p is not None:
        fwd_bwd(p2p.views); p2p.exchange(means3D, campos_all, 3)
        torch.cuda.synchronize()
        p2p_rel = {k: float((p2p.grads[k].double() - dense.views[k].double()).norm() / (dense.views[k].double().norm() + 1e-30))
                   for k in ("means3D", "features", "opacity", "scales", "rotations", "sh")}
        c2 = torch.stack([p2p.sh.view(torch.int32).sum(dtype=torch.int64), p2p.flat[:p2p.n_dense].view(torch.int32).sum(dtype=torch.int64)])
        lo2, hi2 = c2.clone(), c2.clone()
        tdist.all_reduce(lo2, op=tdist.ReduceOp.MIN); tdist.all_reduce(hi2, op=tdist.ReduceOp.MAX)
        p2p_info["identical_on_all_ranks"] = bool((lo2 == hi2).all())
    # ---- timing: exchange only, and the whole step -----------------------------------------------
    t = dict(dense_exchange_ms=timed(dense.allreduce_mean), factored_exchange_ms=timed(lambda: fact.exchange(means3D, campos_all, 3)),
             dense_step_ms=timed(lambda: (fwd_bwd(dense.views), dense.allreduce_mean())),
             factored_step_ms=timed(lambda: (fwd_bwd(fact.views), fact.exchange(means3D, campos_all, 3))))
    if p2p is not None:
        from relightable3dgaussian_b200.dist import _nvls_mean_inplace
        t.update(p2p_barrier_pair_ms=timed(lambda: (p2p.hdl.barrier(channel=0), p2p.hdl.barrier(channel=1))),
                 p2p_sh_only_ms=timed(lambda: p2p.exchange(means3D, campos_all, 3, dense=False)),
                 p2p_dense_only_ms=timed(lambda: _nvls_mean_inplace(p2p.flat[:p2p.n_dense], p2p.hdl, p2p.world, p2p.rank, p2p.multicast, p2p.peer_ptrs)),
                 nccl_allreduce_dense_rest_ms=timed(fact.dense.allreduce_mean), nccl_allgather_factors_ms=timed(fact.gather_factors),
                 local_rebuild_ms=timed(lambda: fact.rebuild_sh(means3D, campos_all, 3)))
        t.update(p2p_exchange_ms=timed(lambda: p2p.exchange(means3D, campos_all, 3)),
                 p2p_step_ms=timed(lambda: (fwd_bwd(p2p.views), p2p.exchange(means3D, campos_all, 3))))
    t["single_gpu_step_ms"] = timed(lambda: fwd_bwd(dense.views))
    # ---- hardware correctness gate: the N-view step's gradients == the mean of N single-view backward
    # passes of the REFERENCE's own kernels (oracle/_ref), run sequentially on this rank ---------------------------------
    ref_rel = None
    from oracle import ref_gpu
    if ref_gpu.available():
        ref = ref_gpu.RefRasterizer()
        acc = None
        for c in cams:
            kw = dict(means3D=means3D, shs=shs, scales=scales, rotations=rots, features=feats)
            cam_kw = dict(viewmatrix=d(c.viewmatrix), projmatrix=d(c.projmatrix), campos=d(c.campos))
            o = ref.forward(bg=bg, W=W, H=H, tan_fovx=c.tanfovx, tan_fovy=c.tanfovy, cx=c.cx, cy=c.cy, opacities=opac, **cam_kw, **kw)
            gr = ref.backward(o, bg=bg, tan_fovx=c.tanfovx, tan_fovy=c.tanfovy, dL_dcolor=cot[0], dL_dopacity=cot[1], dL_ddepth=cot[2],
                              dL_dfeature=cot[3], **cam_kw, **kw)
            acc = {k: v.double() / world for k, v in gr.items()} if acc is None else {k: acc[k] + v.double() / world for k, v in gr.items()}
        names = dict(means3D="dL_dmeans3D", features="dL_dfeatures", opacity="dL_dopacity", scales="dL_dscales", rotations="dL_drotations", sh="dL_dsh")
        ref_rel = {k: float((fact.grads[k].double() - acc[n]).norm() / (acc[n].norm() + 1e-30)) for k, n in names.items()}
        worst = torch.tensor([max(ref_rel.values())], device=dev)
        tdist.all_reduce(worst, op=tdist.ReduceOp.MAX)
        ref_rel["max_over_ranks"] = float(worst.item())
    if rank == 0:
        ok = all(v < 1e-4 for v in rel.values()) and int(lo) == int(hi)      # atomics: run-to-run summation order in the two backwards
        if ref_rel is not None:
            ok = ok and ref_rel["max_over_ranks"] < 1e-3
        if p2p_rel is not None:
            ok = ok and all(v < 1e-4 for v in p2p_rel.values()) and p2p_info["identical_on_all_ranks"]
        print(json.dumps(dict(what="exchange check", world=world, P=P, rel_l2_factored_vs_dense=rel, identical_on_all_ranks=int(lo) == int(hi),
                              dense_bytes=dense.bytes(), factored_bytes_per_rank=fact.bytes(),
                              rel_l2_vs_mean_of_reference_single_view_backwards=ref_rel, rel_l2_p2p_vs_dense=p2p_rel, p2p=p2p_info, ok=ok, **t)), flush=True)
        if not ok:
            sys.exit(1)
    tdist.barrier()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
