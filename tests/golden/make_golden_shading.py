#!/usr/bin/env python3
"""Generates tests/golden/shading_*.npz by EXECUTING the reference's own function bodies in this
container (CPU, fp32): `rendering_equation` + `GGX_specular`
(/root/reference/gaussian_renderer/neilf.py:339-406) and `DirectLightMap.direct_light` / `get_env`
(/root/reference/scene/direct_light_map.py:70-83,104-106).  The modules themselves cannot be
imported here (kornia / plyfile / nvdiffrast / cv2 are absent), so the function sources are cut
out of the files with `ast` and compiled unmodified into a namespace that provides torch, numpy,
F and the reference's eval_sh (utils/sh_utils.py, imported by path).  Gradients come from autograd
through those reference functions.  /root/reference is absent on the GPU box: only the outputs
are committed."""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import shading_case  # noqa: E402

REF = "/root/reference"


def cut(path, names, cls=None):
    src = open(path).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    out = []
    for n in body:
        if isinstance(n, ast.FunctionDef) and n.name in names:
            seg = ast.get_source_segment(src, n)
            decos = "".join("@" + ast.get_source_segment(src, d) + "\n" for d in n.decorator_list)
            out.append(decos + seg)
    assert len(out) == len(names), (path, names)
    return out


spec = importlib.util.spec_from_file_location("ref_sh_utils", os.path.join(REF, "utils/sh_utils.py"))
sh_utils = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sh_utils)
ns = dict(torch=torch, np=np, F=F, eval_sh=sh_utils.eval_sh)
for s in cut(os.path.join(REF, "gaussian_renderer/neilf.py"), ["rendering_equation", "GGX_specular"]):
    exec(compile(s, "neilf.py", "exec"), ns)
import textwrap
methods = cut(os.path.join(REF, "scene/direct_light_map.py"), ["direct_light", "get_env"], cls="DirectLightMap")
exec(compile("class RefLight:\n" + "\n".join(textwrap.indent(m, "    ") for m in methods), "direct_light_map.py", "exec"), ns)


def run(name, P, N, He, seed):
    c = shading_case(P, N, He, seed)
    light = ns["RefLight"]()
    light.env = c["env_raw"].clone().requires_grad_(True)          # [1,He,2He,3], softplus applied by get_env
    leaves = {k: c[k].clone().requires_grad_(True) for k in ("base_color", "roughness", "viewdirs", "incidents")}
    pbr, extra = ns["rendering_equation"](leaves["base_color"], leaves["roughness"], c["normals"].detach(), leaves["viewdirs"],
                                          leaves["incidents"], light, visibility_precompute=c["visibility"],
                                          incident_dirs_precompute=c["incident_dirs"], incident_areas_precompute=c["incident_areas"])
    loss = (pbr * c["cot_pbr"]).sum() + (extra["diffuse_light"] * c["cot_diffuse"]).sum() + (extra["specular"] * c["cot_specular"]).sum()
    loss.backward()
    rec = {"in_" + k: v.detach().numpy() for k, v in c.items()}
    rec.update(pbr=pbr.detach().numpy(), **{"x_" + k: v.detach().numpy() for k, v in extra.items()})
    rec.update(**{"grad_" + k: v.grad.numpy() for k, v in leaves.items()}, grad_env_raw=light.env.grad.numpy())
    path = os.path.join(HERE, f"shading_{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"wrote {path}: pbr mean {pbr.mean().item():.4f}, {os.path.getsize(path) / 1e3:.0f} kB")


run("n24", 240, 24, 16, 0)
run("n40_env8", 96, 40, 8, 1)
