"""TEST INFRASTRUCTURE — not product code.

PyTorch (fp32, device-agnostic) restatement of the reference's incident-direction sampling and of the
inverse-covariance helper, the checker for r3dg_sample_incident_dirs / r3dg_bvh_bake_visibility:
  rotation_between_z            utils/sh_utils.py:36-68
  fibonacci_sphere_sampling     utils/graphics_utils.py:9-37
  sample_incident_rays          scene/gaussian_model.py:20-28
  inverse_covariance            scene/gaussian_model.py:257-260 + utils/general_utils.py:66-79,82-103,151-160
Parity status: PINNED — tests/test_oracle_cpu.py compares these functions with the reference's own
(loaded from /root/reference/utils/*.py in the build container) and with tests/golden/sampling.npz
(generated from the reference functions by tests/golden/make_golden_sampling.py).
Only tests/, smoke() and bench.py's reference arm may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F


def rotation_between_z(vec):
    """Rotation taking +z to `vec` ([...,3] -> [...,3,3]); -I when vec.z + 1 <= 0 (utils/sh_utils.py:36-68).

    Rodrigues' formula about the axis (-n.y, n.x, 0).  Because the axis has no z component, the reference's nine
    expressions collapse exactly in IEEE arithmetic (x + 0 == x, -0 - x == -x, 0 / c == 0), which is the form written
    here — and in the CUDA kernel (csrc/bvh.cu rotate_from_z); tests/golden/sampling.npz pins the equality bit for bit."""
    nx, ny, nz = vec.unbind(-1)
    ax, ay = -ny, nx
    c = (nz + 1).clamp_min(1e-7)
    xy = (ax * ay) / c
    rows = torch.stack([torch.stack([1 + (-(ay * ay)) / c, xy, ay], -1),
                        torch.stack([xy, 1 + (-(ax * ax)) / c, -ax], -1),
                        torch.stack([-ay, ax, 1 + (-(ay * ay) - ax * ax) / c], -1)], -2)
    flip = -torch.eye(3, dtype=torch.float32, device=vec.device).expand_as(rows)
    return torch.where((nz + 1 > 0)[..., None, None], rows, flip)


def fibonacci_sphere_sampling(normals, sample_num, random_rotate=True, phase=None):
    """Hemisphere directions around `normals` on a Fibonacci spiral (utils/graphics_utils.py:9-37): sample i sits at height
    z_i = max(1 - 2 i / (2N - 1), sin 10 deg) and azimuth i * pi (3 - sqrt 5) (+ a per-normal random phase when
    `random_rotate`; `phase` replaces the internal torch.rand draw so a test can feed the kernel the same numbers).
    Returns (dirs [...,N,3], areas [...,N,1] = 2 pi)."""
    lead = normals.shape[:-1]
    n = normals.reshape(-1, 3)
    i = torch.arange(sample_num, dtype=torch.float, device=n.device)[None]
    z = (1 - 2 * i / (2 * sample_num - 1)).clamp_min(np.sin(10 / 180 * np.pi))
    ring = torch.sqrt(1 - z ** 2)
    azimuth = (np.pi * (3.0 - np.sqrt(5.0))) * i
    if random_rotate:
        u = torch.rand(n.shape[0], 1, device=n.device) if phase is None else phase.reshape(-1, 1)
        azimuth = u * 2 * np.pi + azimuth
    canonical = torch.stack([torch.sin(azimuth) * ring, torch.cos(azimuth) * ring, z.expand_as(azimuth)], dim=-2)     # [*,3,N]
    dirs = F.normalize(rotation_between_z(n) @ canonical, dim=-2).transpose(-1, -2)
    areas = torch.ones_like(dirs)[..., 0:1] * 2 * np.pi
    return dirs.reshape(*lead, sample_num, 3), areas.reshape(*lead, sample_num, 1)


def sample_incident_rays(normals, is_training=False, sample_num=24):
    """scene/gaussian_model.py:20-28."""
    return fibonacci_sphere_sampling(normals, sample_num, random_rotate=bool(is_training))


def build_rotation(r):
    """utils/general_utils.py:82-103."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def inverse_covariance(scaling, rotation, scaling_modifier=1.0):
    """get_inverse_covariance: covariance_activation(1 / scaling, 1 / modifier, rotation) = strip_symmetric(L L^T),
    L = build_scaling_rotation(modifier' * scaling', rotation) (scene/gaussian_model.py:33-38,257-260)."""
    s = (1.0 / scaling_modifier) * (1.0 / scaling)
    L = torch.zeros((s.shape[0], 3, 3), dtype=torch.float, device=s.device)
    L[:, 0, 0], L[:, 1, 1], L[:, 2, 2] = s[:, 0], s[:, 1], s[:, 2]
    L = build_rotation(rotation) @ L
    cov = L @ L.transpose(1, 2)
    out = torch.zeros((s.shape[0], 6), dtype=torch.float, device=s.device)
    out[:, 0], out[:, 1], out[:, 2] = cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2]
    out[:, 3], out[:, 4], out[:, 5] = cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]
    return out
