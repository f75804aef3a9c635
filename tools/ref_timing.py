#!/usr/bin/env python3
"""Development probe: what does the reference backward's run time depend on?"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relightable3dgaussian_b200 import synth
from oracle import ref_gpu
dev = "cuda"
P, W, H, S = 1_000_000, 800, 800, 5
sc = synth.make_scene(P, "shell-v1", 0, S)
d = lambda t: t.to(dev)
kw = dict(means3D=d(sc.means3D), opacities=d(sc.opacities), shs=d(sc.shs), scales=d(sc.scales), rotations=d(sc.rotations), features=d(sc.features))
ref = ref_gpu.RefRasterizer()
def run(view, bgv, seed, scale=1.0, n=5):
    cam = synth.make_camera(view, W, H)
    cd = dict(viewmatrix=d(cam.viewmatrix), projmatrix=d(cam.projmatrix), campos=d(cam.campos))
    bg = torch.tensor(bgv, device=dev)
    g = torch.Generator().manual_seed(seed)
    cot = [torch.randn(c, H, W, generator=g).to(dev) * scale for c in (3, 1, 1, S)]
    ts = []
    for i in range(n + 2):
        o = ref.forward(bg=bg, W=W, H=H, tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, cx=cam.cx, cy=cam.cy, **cd, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ref.backward(o, bg=bg, tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, dL_dcolor=cot[0], dL_dopacity=cot[1], dL_ddepth=cot[2], dL_dfeature=cot[3], **cd, **{k: v for k, v in kw.items() if k != "opacities"})
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts[2:]) / n
for view in (1, 5):
    for bgv in ([0.0, 0.0, 0.0], [0.1, 0.2, 0.3]):
        for seed in (1, 1234):
            print(json.dumps(dict(view=view, bg=bgv, seed=seed, ref_bwd_ms=run(view, bgv, seed))), flush=True)
