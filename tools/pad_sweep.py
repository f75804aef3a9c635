#!/usr/bin/env python3
"""Development probe: stage times of the headline step for several values of an environment tuning
knob (default R3DG_PAD_GRAD, bytes inserted in front of the gradient rows in the geometry buffer).
usage: pad_sweep.py [ENV_NAME] v0 v1 ...   -> one JSON line per value."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from relightable3dgaussian_b200 import _C_raster as C, _lib, dist as rdist  # noqa: E402


def main():
    args = sys.argv[1:]
    name = "R3DG_PAD_GRAD"
    if args and not args[0].lstrip("-").isdigit():
        name, args = args[0], args[1:]
    vals = args or ["0"]
    cfg = bench.HEADLINE
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    sc, cams, cot, _ = bench.make_inputs(cfg, dev)
    P, S, W, H = cfg["P"], cfg["S"], cfg["W"], cfg["H"]
    d = lambda t: t.to(dev)
    means3D, scales, rots, opac, shs, feats = map(d, (sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs, sc.features))
    bg = torch.zeros(3, device=dev)
    camd = [dict(view=d(c.viewmatrix), proj=d(c.projmatrix), pos=d(c.campos), c=c) for c in cams]
    dcot = {k: d(v) for k, v in cot.items()}
    E = torch.Tensor([])
    bucket = rdist.GradBucket(P, S, 16, dev)

    def step(i):
        cam = camd[i % cfg["views"]]
        c = cam["c"]
        out = C.rasterize_gaussians(bg, means3D, feats, E, opac, scales, rots, 1.0, E, cam["view"], cam["proj"],
                                    c.tanfovx, c.tanfovy, c.cx, c.cy, H, W, shs, 3, cam["pos"], False, True, False)
        C.rasterize_gaussians_backward(bg, means3D, feats, out[9], E, scales, rots, 1.0, E, cam["view"], cam["proj"],
                                       c.tanfovx, c.tanfovy, dcot["color"], dcot["opacity"], dcot["depth"],
                                       dcot["feature"], shs, 3, cam["pos"], out[10], out[0], out[11], out[12],
                                       True, False, _out=bucket.views)

    for v in vals:
        os.environ[name] = v
        for i in range(5):
            step(i)
        torch.cuda.synchronize()
        steps = 40
        lib.r3dg_prof_begin(steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(5 + i)
        e1.record()
        torch.cuda.synchronize()
        arr = (ctypes.c_float * 9)(); nf = ctypes.c_int(); nb = ctypes.c_int()
        lib.r3dg_prof_end(arr, ctypes.byref(nf), ctypes.byref(nb))
        st = {n: round(arr[i] / steps, 4) for i, n in enumerate(bench.STAGES)}
        print(json.dumps({name: v, "ms_per_step": round(e0.elapsed_time(e1) / steps, 4), **st}), flush=True)


if __name__ == "__main__":
    main()
