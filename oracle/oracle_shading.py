"""TEST INFRASTRUCTURE — not product code.

PyTorch (fp32, device-agnostic) restatement of the reference's BRDF shading step:
  rendering_equation      gaussian_renderer/neilf.py:339-371
  GGX_specular            gaussian_renderer/neilf.py:374-406
  direct_light / get_env  scene/direct_light_map.py:70-83,104-106 (scene/envmap.py:35-53 with transform)
  eval_sh                 utils/sh_utils.py:71-128
Gradients come from autograd.  Parity status: PINNED against tests/golden/shading_*.npz, which hold
outputs and autograd gradients of the reference's own function bodies executed in the build
container (tests/golden/make_golden_shading.py).  Only tests/, smoke() and bench.py may import it.
"""
import numpy as np
import torch
import torch.nn.functional as F

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def eval_sh3(sh, dirs):
    """sh [..., C, 16], dirs [..., 3] -> [..., C] (degree 3)."""
    x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    r = C0 * sh[..., 0]
    r = r - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
    r = (r + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] +
         C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
    r = (r + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10] +
         C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] +
         C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14] +
         C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return r


def direct_light(env_tex, dirs, transform=None):
    """env_tex [H,W,3] (already activated); lat-long lookup with bilinear grid_sample, align_corners=True."""
    shape = dirs.shape
    d = dirs.reshape(-1, 3)
    if transform is not None:
        d = d @ transform.T
    envir_map = env_tex.permute(2, 0, 1).unsqueeze(0)
    phi = torch.arccos(d[:, 2]).reshape(-1) - 1e-6
    theta = torch.atan2(d[:, 1], d[:, 0]).reshape(-1)
    query_y = (phi / np.pi) * 2 - 1
    query_x = -theta / np.pi
    grid = torch.stack((query_x, query_y)).permute(1, 0).unsqueeze(0).unsqueeze(0)
    light = F.grid_sample(envir_map, grid, align_corners=True).squeeze().permute(1, 0).reshape(-1, 3)
    return light.reshape(*shape)


def ggx_specular(normal, pts2c, pts2l, roughness, fresnel=0.04):
    L = F.normalize(pts2l, dim=-1)
    V = F.normalize(pts2c, dim=-1)
    H = F.normalize((L + V[:, None, :]) / 2.0, dim=-1)
    N = F.normalize(normal, dim=-1)
    NoV = torch.sum(V * N, dim=-1, keepdim=True)
    N = N * NoV.sign()
    NoL = torch.sum(N[:, None, :] * L, dim=-1, keepdim=True).clamp(1e-6, 1)
    NoV = torch.sum(N * V, dim=-1, keepdim=True).clamp(1e-6, 1)
    NoH = torch.sum(N[:, None, :] * H, dim=-1, keepdim=True).clamp(1e-6, 1)
    VoH = torch.sum(V[:, None, :] * H, dim=-1, keepdim=True).clamp(1e-6, 1)
    alpha = roughness * roughness
    alpha2 = alpha * alpha
    k = (alpha + 2 * roughness + 1.0) / 8.0
    FMi = ((-5.55473) * VoH - 6.98316) * VoH
    frac0 = fresnel + (1 - fresnel) * torch.pow(2.0, FMi)
    frac = frac0 * alpha2[:, None, :]
    nom0 = NoH * NoH * (alpha2[:, None, :] - 1) + 1
    nom1 = NoV * (1 - k) + k
    nom2 = NoL * (1 - k[:, None, :]) + k[:, None, :]
    nom = (4 * np.pi * nom0 * nom0 * nom1[:, None, :] * nom2).clamp(1e-6, 4 * np.pi)
    return frac / nom


def rendering_equation(base_color, roughness, normals, viewdirs, incidents, env_tex, visibility, incident_dirs,
                       incident_areas, transform=None):
    global_l = direct_light(env_tex, incident_dirs, transform) * visibility
    local_l = eval_sh3(incidents.transpose(1, 2).reshape(-1, 1, 3, 16), incident_dirs).clamp_min(0)
    lights = local_l + global_l
    n_d_i = (normals[:, None] * incident_dirs).sum(-1, keepdim=True).clamp(min=0)
    f_d = base_color[:, None] / np.pi
    f_s = ggx_specular(normals, viewdirs, incident_dirs, roughness, fresnel=0.04)
    transport = lights * incident_areas * n_d_i
    specular = (f_s * transport).mean(dim=-2)
    pbr = ((f_d + f_s) * transport).mean(dim=-2)
    diffuse_light = transport.mean(dim=-2)
    extra = {"incident_dirs": incident_dirs, "incident_lights": lights, "local_incident_lights": local_l,
             "global_incident_lights": global_l, "incident_visibility": visibility, "diffuse_light": diffuse_light,
             "specular": specular}
    return pbr, extra
