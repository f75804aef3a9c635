#!/usr/bin/env python3
"""Ad-hoc GPU parity + timing probe (development aid; the real checks live in tests/).
Compares our kernels against the unmodified reference kernels (oracle/_ref) and the CPU oracle."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relightable3dgaussian_b200 import _C_raster as C, synth  # noqa: E402
from oracle import ref_gpu  # noqa: E402


def run_ours(sc, cam, bg, S, pseudo=True, colors_precomp=False, cov_precomp=None):
    dev = "cuda"
    E = torch.Tensor([])
    feats = sc.features.to(dev) if S > 0 else torch.empty((sc.means3D.shape[0], 0), device=dev)
    out = C.rasterize_gaussians(
        bg, sc.means3D.to(dev), feats, E, sc.opacities.to(dev), sc.scales.to(dev),
        sc.rotations.to(dev), 1.0, E, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.tanfovx,
        cam.tanfovy, cam.cx, cam.cy, cam.image_height, cam.image_width, sc.shs.to(dev), 3,
        cam.campos.to(dev), False, pseudo, False)
    return out


def cmp(name, a, b, exact=False):
    a = a.detach().cpu(); b = b.detach().cpu()
    if exact:
        neq = int((a != b).sum())
        print(f"  {name:18s} exact mismatches: {neq} / {a.numel()}")
        return neq
    d = (a.double() - b.double()).abs()
    bits = int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum()) if a.dtype == torch.float32 else -1
    rel = float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
    print(f"  {name:18s} max-abs {float(d.max()) if d.numel() else 0:.3e}  rel-l2 {rel:.3e}  bit-mismatch {bits}/{a.numel()}")
    return float(d.max()) if d.numel() else 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=100000)
    ap.add_argument("--W", type=int, default=800)
    ap.add_argument("--H", type=int, default=800)
    ap.add_argument("--S", type=int, default=5)
    ap.add_argument("--view", type=int, default=1)
    ap.add_argument("--recipe", default="shell-v1")
    ap.add_argument("--time", type=int, default=0)
    args = ap.parse_args()
    dev = "cuda"
    sc = synth.make_scene(args.P, args.recipe, 0, args.S)
    cam = synth.make_camera(args.view, args.W, args.H)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    P, S, W, H = args.P, args.S, args.W, args.H
    ours = run_ours(sc, cam, bg, S)
    torch.cuda.synchronize()
    (R, ncon, color, opac, depth, feat, normal, xyz, weights, radii, gB, bB, iB) = ours
    print(f"ours: R={R}")
    ref = ref_gpu.RefRasterizer()
    d = lambda t: None if t is None else t.to(dev)
    kw = dict(means3D=d(sc.means3D), opacities=d(sc.opacities), viewmatrix=d(cam.viewmatrix),
              projmatrix=d(cam.projmatrix), campos=d(cam.campos), bg=bg, W=W, H=H,
              tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, cx=cam.cx, cy=cam.cy, shs=d(sc.shs),
              scales=d(sc.scales), rotations=d(sc.rotations), features=d(sc.features) if S else None)
    ro = ref.forward(**kw)
    torch.cuda.synchronize()
    print(f"ref : R={ro['num_rendered']}")
    cmp("radii", radii, ro["radii"], True)
    for nm in ("tiles_touched", "point_offsets"):
        cmp(nm, C.debug_intermediate(nm, P, S, W, H, gB, iB, bB), ref.intermediate(nm), True)
    vis = (ro["radii"] > 0)
    for nm in ("depths", "means2D", "conic_opacity", "rgb"):
        a = C.debug_intermediate(nm, P, S, W, H, gB, iB, bB)
        b = ref.intermediate(nm)
        m = vis if a.dim() == 1 else vis[:, None].expand_as(a)
        cmp(nm, torch.where(m, a, torch.zeros_like(a)), torch.where(m, b, torch.zeros_like(b)))
    if R == ro["num_rendered"]:
        cmp("point_list_keys", C.debug_intermediate("point_list_keys", P, S, W, H, gB, iB, bB, R), ref.intermediate("point_list_keys"), True)
        cmp("point_list", C.debug_intermediate("point_list", P, S, W, H, gB, iB, bB, R), ref.intermediate("point_list"), True)
    cmp("ranges", C.debug_intermediate("ranges", P, S, W, H, gB, iB, bB), ref.intermediate("ranges"), True)
    cmp("n_contrib", ncon.reshape(-1), ref.intermediate("n_contrib"), True)
    cmp("final_T", C.debug_intermediate("final_T", P, S, W, H, gB, iB, bB).reshape(-1), ref.intermediate("final_T"))
    for nm, t in (("color", color), ("opacity", opac), ("depth", depth), ("feature", feat), ("normal", normal), ("surface_xyz", xyz), ("weights", weights)):
        cmp(nm, t, ro[nm])
    # backward
    g = torch.Generator().manual_seed(1)
    dc = torch.randn(3, H, W, generator=g).to(dev); do = torch.randn(1, H, W, generator=g).to(dev)
    dd = torch.randn(1, H, W, generator=g).to(dev); df = torch.randn(S, H, W, generator=g).to(dev)
    E = torch.Tensor([])
    feats = d(sc.features) if S else torch.empty((P, 0), device=dev)
    og = C.rasterize_gaussians_backward(bg, d(sc.means3D), feats, radii, E, d(sc.scales), d(sc.rotations), 1.0, E,
                                        d(cam.viewmatrix), d(cam.projmatrix), cam.tanfovx, cam.tanfovy, dc, do, dd, df,
                                        d(sc.shs), 3, d(cam.campos), gB, R, bB, iB, True, False)
    torch.cuda.synchronize()
    rg = ref.backward(ro, means3D=d(sc.means3D), viewmatrix=d(cam.viewmatrix), projmatrix=d(cam.projmatrix),
                      campos=d(cam.campos), bg=bg, tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, dL_dcolor=dc,
                      dL_dopacity=do, dL_ddepth=dd, dL_dfeature=df, shs=d(sc.shs), scales=d(sc.scales),
                      rotations=d(sc.rotations), features=d(sc.features) if S else None)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dfeatures", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for nm, t in zip(names, og):
        cmp(nm, t, rg[nm])
    if args.time:
        def timeit(fn, n=args.time):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        t_of = timeit(lambda: run_ours(sc_dev, cam_dev, bg, S))
        def ours_fb():
            o = run_ours(sc_dev, cam_dev, bg, S)
            C.rasterize_gaussians_backward(bg, sc_dev.means3D, feats, o[9], E, sc_dev.scales, sc_dev.rotations, 1.0, E,
                                           cam_dev.viewmatrix, cam_dev.projmatrix, cam.tanfovx, cam.tanfovy, dc, do, dd, df,
                                           sc_dev.shs, 3, cam_dev.campos, o[10], o[0], o[11], o[12], True, False)
        t_ofb = timeit(ours_fb)
        t_rf = timeit(lambda: ref.forward(**kw))
        bkw = {k: v for k, v in kw.items() if k in ("means3D", "viewmatrix", "projmatrix", "campos", "bg", "tan_fovx", "tan_fovy", "shs", "scales", "rotations", "features")}
        def ref_fb():
            r = ref.forward(**kw)
            ref.backward(r, dL_dcolor=dc, dL_dopacity=do, dL_ddepth=dd, dL_dfeature=df, **bkw)
        t_rfb = timeit(ref_fb)
        print(json.dumps(dict(P=P, R=R, ours_fwd_ms=t_of, ours_fwdbwd_ms=t_ofb, ref_fwd_ms=t_rf, ref_fwdbwd_ms=t_rfb)))


if __name__ == "__main__":
    a = sys.argv
    # device-resident copies for timing
    main_args = None
    import types
    # build device copies lazily inside main via globals
    ap_P = int(a[a.index("--P") + 1]) if "--P" in a else 100000
    ap_S = int(a[a.index("--S") + 1]) if "--S" in a else 5
    ap_W = int(a[a.index("--W") + 1]) if "--W" in a else 800
    ap_H = int(a[a.index("--H") + 1]) if "--H" in a else 800
    ap_v = int(a[a.index("--view") + 1]) if "--view" in a else 1
    ap_r = a[a.index("--recipe") + 1] if "--recipe" in a else "shell-v1"
    _sc = synth.make_scene(ap_P, ap_r, 0, ap_S)
    _cam = synth.make_camera(ap_v, ap_W, ap_H)
    sc_dev = synth.Scene(*[None if t is None else t.cuda() for t in _sc])
    cam_dev = types.SimpleNamespace(**{k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in _cam._asdict().items()})
    main()
