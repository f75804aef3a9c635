#!/usr/bin/env python3
"""Generates tests/golden/raster_*.npz on the GPU box from the UNMODIFIED reference CUDA kernels
(oracle/_ref/libref_raster.so, built from /root/reference by oracle/build_ref.sh):

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy into tests/golden/

Each fixture holds the inputs (so the test does not depend on RNG reproducibility), every
observable intermediate of the reference forward (radii, tiles_touched, depths, means2D,
conic_opacity, rgb, sorted keys / point_list, tile ranges, n_contrib, final_T), its outputs and,
for seeded cotangents, its nine gradient tensors.  These pin the CPU oracle (tests -m "not gpu")
and are a second, reference-owned check for our kernels (tests -m gpu)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import case_inputs, npy  # noqa: E402
from oracle import ref_gpu  # noqa: E402

CASES = {
    # name: (P, W, H, S, view, recipe, scale_boost, center_shift, mode, sh_degree)
    "shell_s5": (900, 72, 40, 5, 1, "shell-v1", 4.0, False, "sh_sr", 3),
    "cube_s0_precomp": (700, 64, 48, 0, 3, "cube-v1", 5.0, True, "col_cov", 3),
    "shell_s16_deg1": (600, 50, 34, 16, 6, "shell-v1", 4.0, True, "sh_sr", 1),
}


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda")
    for name, (P, W, H, S, view, recipe, boost, shift, mode, deg) in CASES.items():
        sc, cam = case_inputs(P, W, H, S, view, recipe, seed=7, scale_boost=boost, center_shift=shift)
        bg = torch.tensor([0.3, 0.1, 0.2])
        d = lambda t: None if t is None else t.to(dev)
        ref = ref_gpu.RefRasterizer()
        kw = dict(means3D=d(sc.means3D), opacities=d(sc.opacities), viewmatrix=d(cam.viewmatrix),
                  projmatrix=d(cam.projmatrix), campos=d(cam.campos), bg=d(bg), W=W, H=H, tan_fovx=cam.tanfovx,
                  tan_fovy=cam.tanfovy, cx=cam.cx, cy=cam.cy, features=d(sc.features) if S else None, sh_degree=deg)
        extra = {}
        if mode == "sh_sr":
            kw.update(shs=d(sc.shs), scales=d(sc.scales), rotations=d(sc.rotations))
        else:
            # precomputed colours and covariances: take the reference's own cov3D of a first pass
            r0 = ref_gpu.RefRasterizer()
            r0.forward(**dict(kw, shs=d(sc.shs), scales=d(sc.scales), rotations=d(sc.rotations)))
            cov = r0.intermediate("cov3D").clone()
            g = torch.Generator().manual_seed(3)
            col = torch.rand(P, 3, generator=g)
            kw.update(colors_precomp=d(col), cov3D_precomp=cov)
            extra = dict(in_colors_precomp=npy(col), in_cov3D_precomp=npy(cov))
        fo = ref.forward(**kw)
        R = fo["num_rendered"]
        inter = {k: npy(ref.intermediate(k)) for k in
                 ("depths", "means2D", "conic_opacity", "rgb", "clamped", "tiles_touched", "point_offsets", "cov3D",
                  "point_list", "point_list_keys", "ranges", "n_contrib", "final_T")}
        g = torch.Generator().manual_seed(11)
        cot = dict(color=torch.randn(3, H, W, generator=g), opacity=torch.randn(1, H, W, generator=g),
                   depth=torch.randn(1, H, W, generator=g), feature=torch.randn(S, H, W, generator=g))
        bkw = {k: v for k, v in kw.items() if k in ("means3D", "viewmatrix", "projmatrix", "campos", "bg", "tan_fovx",
                                                     "tan_fovy", "shs", "colors_precomp", "scales", "rotations",
                                                     "cov3D_precomp", "features", "sh_degree")}
        gr = ref.backward(fo, dL_dcolor=d(cot["color"]), dL_dopacity=d(cot["opacity"]), dL_ddepth=d(cot["depth"]),
                          dL_dfeature=d(cot["feature"]), **bkw)
        torch.cuda.synchronize()
        rec = dict(meta=np.array([P, W, H, S, view, deg, R], np.int64), mode=np.array(mode),
                   tanfov=np.array([cam.tanfovx, cam.tanfovy, cam.cx, cam.cy], np.float64),
                   in_means3D=npy(sc.means3D), in_scales=npy(sc.scales), in_rotations=npy(sc.rotations),
                   in_opacities=npy(sc.opacities), in_shs=npy(sc.shs), in_bg=npy(bg),
                   in_viewmatrix=npy(cam.viewmatrix), in_projmatrix=npy(cam.projmatrix), in_campos=npy(cam.campos),
                   **extra)
        if S:
            rec["in_features"] = npy(sc.features)
        for k in ("color", "opacity", "depth", "feature", "normal", "surface_xyz", "weights", "radii"):
            rec["out_" + k] = npy(fo[k])
        for k, v in inter.items():
            rec["mid_" + k] = v
        for k, v in cot.items():
            rec["cot_" + k] = npy(v)
        for k, v in gr.items():
            rec["grad_" + k] = npy(v)
        path = os.path.join(outdir, f"raster_{name}.npz")
        np.savez_compressed(path, **rec)
        print(f"wrote {path}: R={R}, visible={(fo['radii'] > 0).sum().item()}/{P}, {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE))
