"""Shared test utilities: seeded cases, conversions, comparison helpers."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from relightable3dgaussian_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def sh_case(P, seed):
    """Inputs of BASELINE.json config #1: sh [P,3,16] (channel-major, utils/sh_utils.py:77), unit dirs."""
    g = torch.Generator().manual_seed(1000 + seed)
    sh = torch.randn(P, 3, 16, generator=g) * 0.3
    dirs = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
    return sh, dirs


def npy(t):
    return None if t is None else t.detach().cpu().numpy()


def case_inputs(P, W, H, S, view=1, recipe="shell-v1", seed=0, scale_boost=1.0, center_shift=False):
    sc = synth.make_scene(P, recipe, seed, S)
    if scale_boost != 1.0:
        sc = sc._replace(scales=sc.scales * scale_boost)
    cam = synth.make_camera(view, W, H, center_shift=center_shift)
    return sc, cam


def oracle_kwargs(sc, cam, bg):
    return dict(means3D=npy(sc.means3D), opacities=npy(sc.opacities), viewmatrix=npy(cam.viewmatrix),
                projmatrix=npy(cam.projmatrix), campos=npy(cam.campos), bg=npy(bg),
                W=cam.image_width, H=cam.image_height, tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy,
                cx=cam.cx, cy=cam.cy)


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def max_rel_above_floor(a, b, floor=1e-6):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    m = np.abs(b) > floor * (np.abs(b).max() + 1e-30)
    if not m.any():
        return 0.0
    return float((np.abs(a - b)[m] / np.abs(b)[m]).max())
