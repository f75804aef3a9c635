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
    """utils/sh_utils.py:36-68: rotation taking +z to `vec` ([...,3] -> [...,3,3]); -I when vec.z + 1 <= 0."""
    v1, v2 = -vec[..., 1], vec[..., 0]
    v3 = torch.zeros_like(v1)
    v11, v22, v33 = v1 * v1, v2 * v2, v3 * v3
    v12, v13, v23 = v1 * v2, v1 * v3, v2 * v3
    cos_p_1 = (vec[..., 2] + 1).clamp_min(1e-7)
    R = torch.zeros(vec.shape[:-1] + (3, 3), dtype=torch.float32, device=vec.device)
    R[..., 0, 0] = 1 + (-v33 - v22) / cos_p_1
    R[..., 0, 1] = -v3 + v12 / cos_p_1
    R[..., 0, 2] = v2 + v13 / cos_p_1
    R[..., 1, 0] = v3 + v12 / cos_p_1
    R[..., 1, 1] = 1 + (-v33 - v11) / cos_p_1
    R[..., 1, 2] = -v1 + v23 / cos_p_1
    R[..., 2, 0] = -v2 + v13 / cos_p_1
    R[..., 2, 1] = v1 + v23 / cos_p_1
    R[..., 2, 2] = 1 + (-v22 - v11) / cos_p_1
    return torch.where((vec[..., 2] + 1 > 0)[..., None, None], R,
                       -torch.eye(3, dtype=torch.float32, device=vec.device).expand_as(R))


def fibonacci_sphere_sampling(normals, sample_num, random_rotate=True, phase=None):
    """utils/graphics_utils.py:9-37.  `phase` replaces the internal torch.rand draw (so that a test can feed
    the same random numbers to the kernel)."""
    pre_shape = normals.shape[:-1]
    if len(pre_shape) > 1:
        normals = normals.reshape(-1, 3)
    delta = np.pi * (3.0 - np.sqrt(5.0))
    idx = torch.arange(sample_num, dtype=torch.float, device=normals.device)[None]
    z = (1 - 2 * idx / (2 * sample_num - 1)).clamp_min(np.sin(10 / 180 * np.pi))
    rad = torch.sqrt(1 - z ** 2)
    theta = delta * idx
    if random_rotate:
        u = torch.rand(*pre_shape, 1, device=normals.device) if phase is None else phase.reshape(-1, 1)
        theta = u * 2 * np.pi + theta
    y = torch.cos(theta) * rad
    x = torch.sin(theta) * rad
    z_samples = torch.stack([x, y, z.expand_as(y)], dim=-2)
    incident_dirs = rotation_between_z(normals) @ z_samples
    incident_dirs = F.normalize(incident_dirs, dim=-2).transpose(-1, -2)
    incident_areas = torch.ones_like(incident_dirs)[..., 0:1] * 2 * np.pi
    if len(pre_shape) > 1:
        incident_dirs = incident_dirs.reshape(*pre_shape, sample_num, 3)
        incident_areas = incident_areas.reshape(*pre_shape, sample_num, 1)
    return incident_dirs, incident_areas


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
