#!/usr/bin/env python3
"""Stage-2 ("neilf") training step on one B200, assembled like gaussian_renderer/neilf.py:87-147 +
scene/gaussian_model.py:495-497 from this repo's operators (development / evidence for
BASELINE.json configs #4/#5, not the headline bench):

    rendering_equation (fused shading, baked visibility)  ->  16-channel feature pack
    -> GaussianRasterizer fwd (S = 16)  ->  un-premultiply + L1 loss  ->  backward (raster + shading)
    -> FusedAdam over the 13 parameter groups

Prints one JSON line with the per-step time and its split.  The baked tensors are random but
well-formed (Fibonacci directions around the normals, visibility in {0} u [0.9, 1]): the BVH bake is
a one-off and is timed by tools/bench_stage3.py.

  python tools/bench_stage2_step.py [--P 1500000 --W 1600 --H 1200 --N 32 --steps 20]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relightable3dgaussian_b200 import shading, synth  # noqa: E402
from relightable3dgaussian_b200.optim import FusedAdam  # noqa: E402
from relightable3dgaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer, set_deferred_count  # noqa: E402
from relightable3dgaussian_b200.raytracer import fibonacci_sphere_sampling  # noqa: E402


class Light:                     # duck-types scene/direct_light_map.py:DirectLightMap
    def __init__(self, env_raw):
        self.env = env_raw

    @property
    def get_env(self):
        return F.softplus(self.env)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1_500_000)       # DTU_scan24 shape (BASELINE.json config #4)
    ap.add_argument("--W", type=int, default=1600)
    ap.add_argument("--H", type=int, default=1200)
    ap.add_argument("--N", type=int, default=32)              # script/run_dtu.sh: --sample_num 32
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam instead of FusedAdam")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    P, W, H, N = a.P, a.W, a.H, a.N
    sc = synth.make_scene(P, "shell-v1", 0, 0)
    cams = [synth.make_camera(k, W, H) for k in range(8)]
    g = torch.Generator().manual_seed(7)
    leaf = lambda t: t.to(dev).requires_grad_(True)
    # raw (pre-activation) parameters, activations as in scene/gaussian_model.py:76-110
    xyz = leaf(sc.means3D); normal = leaf(sc.normals)
    rot_raw = leaf(sc.rotations); scale_raw = leaf(sc.scales.log()); opac_raw = leaf(torch.logit(sc.opacities.clamp(1e-4, 1 - 1e-4)))
    shs_dc = leaf(sc.shs[:, :1].contiguous()); shs_rest = leaf(sc.shs[:, 1:].contiguous())
    base_raw = leaf(torch.randn(P, 3, generator=g)); rough_raw = leaf(torch.randn(P, 1, generator=g))
    inc_dc = leaf(torch.randn(P, 1, 3, generator=g) * 0.3); inc_rest = leaf(torch.randn(P, 15, 3, generator=g) * 0.1)
    env_raw = leaf(torch.randn(1, 16, 32, 3, generator=g))
    groups = [("xyz", xyz, 1.6e-4), ("normal", normal, 1e-3), ("rotation", rot_raw, 1e-3), ("scaling", scale_raw, 5e-3),
              ("opacity", opac_raw, 5e-2), ("f_dc", shs_dc, 2.5e-3), ("f_rest", shs_rest, 1.25e-4), ("base_color", base_raw, 1e-2),
              ("roughness", rough_raw, 1e-2), ("incidents_dc", inc_dc, 2e-3), ("incidents_rest", inc_rest, 1e-4), ("env", env_raw, 1e-2)]
    Opt = torch.optim.Adam if a.torch_adam else FusedAdam
    opt = Opt([{"params": [p], "lr": lr, "name": n} for n, p, lr in groups], lr=0.0, eps=1e-15)
    # baked tensors (gaussian_model.py:312-342), in chunks to bound the CPU-side generation
    dirs = torch.empty((P, N, 3), device=dev); areas = torch.empty((P, N, 1), device=dev)
    for o in range(0, P, 200_000):
        d_, a_ = fibonacci_sphere_sampling(sc.normals[o:o + 200_000].to(dev), N, random_rotate=False)
        dirs[o:o + 200_000], areas[o:o + 200_000] = d_, a_
    u = torch.rand(P, N, 1, device=dev)
    vis = torch.where(u < 0.4, torch.zeros_like(u), 0.9 + 0.1 * torch.rand(P, N, 1, device=dev))
    gts = [torch.rand(3, H, W, device=dev) for _ in cams]
    camd = [dict(view=c.viewmatrix.to(dev), proj=c.projmatrix.to(dev), pos=c.campos.to(dev)) for c in cams]
    bg = torch.zeros(3, device=dev)
    light = Light(env_raw)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []

    def step(i, record=False):
        c, cd = cams[i % 8], camd[i % 8]
        e = [ev() for _ in range(5)] if record else None
        mark = (lambda k: e[k].record()) if record else (lambda k: None)
        mark(0)
        base_color, roughness = torch.sigmoid(base_raw) * 0.77 + 0.03, torch.sigmoid(rough_raw) * 0.9 + 0.09
        incidents = torch.cat([inc_dc, inc_rest], dim=1)
        viewdirs = F.normalize(cd["pos"] - xyz, dim=-1)
        nrm = F.normalize(normal, dim=-1)
        brdf, extra = shading.rendering_equation(base_color, roughness, nrm.detach(), viewdirs, incidents, light,
                                                 visibility_precompute=vis, incident_dirs_precompute=dirs, incident_areas_precompute=areas)
        mark(1)
        view = cd["view"]
        depths = (torch.cat([xyz, torch.ones_like(xyz[:, :1])], dim=-1) @ view)[:, 2:3]
        feats = torch.cat([depths, depths.square(), brdf, nrm, base_color, roughness, extra["diffuse_light"], vis.mean(-2)], dim=-1)   # neilf.py:115-118
        rs = GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, c.cx, c.cy, bg, 1.0, view, cd["proj"], 3, cd["pos"],
                                           False, True, True, False)
        means2D = torch.zeros_like(xyz, requires_grad=True)
        out = GaussianRasterizer(rs)(means3D=xyz, means2D=means2D, opacities=torch.sigmoid(opac_raw), shs=torch.cat([shs_dc, shs_rest], dim=1),
                                     scales=torch.exp(scale_raw), rotations=F.normalize(rot_raw), features=feats)
        num_contrib, color, opacity, feature = out[1], out[2], out[3], out[5]
        feature = feature / opacity.clamp_min(1e-5) * (num_contrib > 0)                                    # neilf.py:136-137
        pbr = feature[2:5] * opacity + (1 - opacity) * bg[:, None, None]
        loss = (color - gts[i % 8]).abs().mean() + (pbr - gts[i % 8]).abs().mean() + 0.01 * feature[12:15].mean()
        mark(2)
        loss.backward()
        mark(3)
        opt.step(); opt.zero_grad()
        mark(4)
        if record:
            marks.append(e)

    set_deferred_count(True)
    for i in range(a.warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(a.steps):
        step(a.warmup + i, record=True)
    e1.record(); torch.cuda.synchronize()
    set_deferred_count(False)
    ms = e0.elapsed_time(e1) / a.steps
    split = [sum(m[k].elapsed_time(m[k + 1]) for m in marks) / len(marks) for k in range(4)]
    print(json.dumps(dict(what="stage-2 training step", P=P, W=W, H=H, N=N, S=16, optimizer="torch.optim.Adam" if a.torch_adam else "FusedAdam",
                          ms_per_step=ms, steps_per_s=1e3 / ms, shading_fwd_ms=split[0], pack_raster_fwd_loss_ms=split[1],
                          backward_ms=split[2], optimizer_ms=split[3], peak_GB=torch.cuda.max_memory_allocated() / 1e9)), flush=True)


if __name__ == "__main__":
    main()
