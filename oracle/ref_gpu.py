"""TEST INFRASTRUCTURE — not product code.

ctypes driver for oracle/_ref/libref_raster.so: the UNMODIFIED reference rasterizer kernels
(/root/reference/r3dg-rasterization/cuda_rasterizer/*.cu) compiled for sm_100 by
oracle/build_ref.sh behind the raw-pointer shim oracle/ref_shim_raster.cu.  It is the parity
oracle on the GPU box (bit-exact checks of radii / point_list / ranges / n_contrib, tolerance
checks of images and gradients) and the timed `--impl reference` arm of bench.py.

Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_raster.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH) and torch.cuda.is_available()


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ref_ctx_create.restype = ctypes.c_void_p
        _lib.ref_copy_out.restype = ctypes.c_longlong
    return _lib


def _p(t):
    return None if t is None or t.numel() == 0 else ctypes.c_void_p(t.data_ptr())


_IDS = {"depths": (0, torch.float32, 1), "clamped": (1, torch.uint8, 3), "radii": (2, torch.int32, 1),
        "means2D": (3, torch.float32, 2), "cov3D": (4, torch.float32, 6),
        "conic_opacity": (5, torch.float32, 4), "rgb": (6, torch.float32, 3),
        "tiles_touched": (7, torch.int32, 1), "point_offsets": (8, torch.int32, 1)}


class RefRasterizer:
    """One reference rasterizer context (three growable device buffers, like the torch glue of
    rasterize_points.cu:28-34,83-90)."""

    def __init__(self):
        self.ctx = ctypes.c_void_p(lib().ref_ctx_create())

    def __del__(self):
        try:
            lib().ref_ctx_destroy(self.ctx)
        except Exception:
            pass

    def forward(self, *, means3D, opacities, viewmatrix, projmatrix, campos, bg, W, H, tan_fovx,
                tan_fovy, cx, cy, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, features=None, sh_degree=3, scale_modifier=1.0,
                prefiltered=False, computer_pseudo_normal=True, debug=False):
        dev = means3D.device
        P = means3D.shape[0]
        S = 0 if features is None else features.shape[1]
        M = 0 if shs is None else shs.shape[1]
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)   # rasterize_points.cu:74-81
        out = dict(color=z(3, H, W), opacity=z(1, H, W), depth=z(1, H, W), feature=z(S, H, W),
                   normal=z(3, H, W), surface_xyz=z(3, H, W), weights=z(P, 1),
                   radii=torch.zeros(P, dtype=torch.int32, device=dev))
        torch.cuda.synchronize()
        R = lib().ref_raster_forward(
            self.ctx, P, S, int(sh_degree), M, _p(bg), W, H, _p(means3D), _p(shs), _p(colors_precomp),
            _p(features), _p(opacities), _p(scales), ctypes.c_float(scale_modifier), _p(rotations),
            _p(cov3D_precomp), _p(viewmatrix), _p(projmatrix), _p(campos), ctypes.c_float(tan_fovx),
            ctypes.c_float(tan_fovy), ctypes.c_float(cx), ctypes.c_float(cy), int(prefiltered),
            int(computer_pseudo_normal), _p(out["color"]), _p(out["opacity"]), _p(out["depth"]),
            _p(out["feature"]), _p(out["normal"]), _p(out["surface_xyz"]), _p(out["weights"]),
            _p(out["radii"]), int(debug))
        if R < 0:
            raise RuntimeError("reference forward failed")
        self.P, self.S, self.M, self.W, self.H, self.R = P, S, M, W, H, R
        out["num_rendered"] = R
        return out

    def intermediate(self, name):
        P, R, HW = self.P, self.R, self.W * self.H
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        dev = torch.device("cuda")
        if name in _IDS:
            i, dt, k = _IDS[name]
            t = torch.empty((P, k) if k > 1 else (P,), dtype=dt, device=dev)
        elif name == "point_list":
            i, t = 9, torch.empty(R, dtype=torch.int32, device=dev)
        elif name == "point_list_keys":
            i, t = 10, torch.empty(R, dtype=torch.int64, device=dev)
        elif name == "point_list_unsorted":
            i, t = 11, torch.empty(R, dtype=torch.int32, device=dev)
        elif name == "point_list_keys_unsorted":
            i, t = 12, torch.empty(R, dtype=torch.int64, device=dev)
        elif name == "final_T":
            i, t = 13, torch.empty(HW, dtype=torch.float32, device=dev)
        elif name == "n_contrib":
            i, t = 14, torch.empty(HW, dtype=torch.int32, device=dev)
        elif name == "ranges":
            i, t = 15, torch.empty((T, 2), dtype=torch.int32, device=dev)
        else:
            raise KeyError(name)
        if t.numel():
            torch.cuda.synchronize()
            n = lib().ref_copy_out(self.ctx, i, _p(t), ctypes.c_longlong(t.numel() * t.element_size()))
            if n < 0:
                raise RuntimeError(f"ref_copy_out({name}) failed: {n}")
            torch.cuda.synchronize()
        return t

    def backward(self, fwd_out, *, means3D, viewmatrix, projmatrix, campos, bg, tan_fovx, tan_fovy,
                 dL_dcolor, dL_dopacity, dL_ddepth, dL_dfeature, shs=None, colors_precomp=None,
                 scales=None, rotations=None, cov3D_precomp=None, features=None, sh_degree=3,
                 scale_modifier=1.0, backward_geometry=True, debug=False):
        dev = means3D.device
        P, S, M, W, H = self.P, self.S, self.M, self.W, self.H
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)   # rasterize_points.cu:183-192
        g = dict(dL_dmeans3D=z(P, 3), dL_dmeans2D=z(P, 3), dL_dfeatures=z(P, S), dL_dcolors=z(P, 3),
                 dL_dconic=z(P, 2, 2), dL_dopacity=z(P, 1), dL_dcov3D=z(P, 6), dL_dsh=z(P, M, 3),
                 dL_dscales=z(P, 3), dL_drotations=z(P, 4))
        torch.cuda.synchronize()
        rc = lib().ref_raster_backward(
            self.ctx, P, S, int(sh_degree), M, self.R, _p(bg), W, H, _p(means3D), _p(shs),
            _p(features), _p(colors_precomp), _p(scales), ctypes.c_float(scale_modifier),
            _p(rotations), _p(cov3D_precomp), _p(viewmatrix), _p(projmatrix), _p(campos),
            ctypes.c_float(tan_fovx), ctypes.c_float(tan_fovy), _p(fwd_out["radii"]),
            _p(dL_dcolor.contiguous()), _p(dL_dopacity.contiguous()), _p(dL_ddepth.contiguous()),
            _p(dL_dfeature.contiguous()), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
            _p(g["dL_dopacity"]), _p(g["dL_dcolors"]), _p(g["dL_dfeatures"]), _p(g["dL_dmeans3D"]),
            _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]),
            int(backward_geometry), int(debug))
        if rc != 0:
            raise RuntimeError("reference backward failed")
        return g


# ----------------------------------------------------------------------------------------------
# reference BVH (oracle/_ref/libref_bvh.so)
# ----------------------------------------------------------------------------------------------
BVH_LIB_PATH = os.path.join(_HERE, "_ref", "libref_bvh.so")
_bvh = None


def bvh_available():
    return os.path.exists(BVH_LIB_PATH) and torch.cuda.is_available()


def bvh_lib():
    global _bvh
    if _bvh is None:
        _bvh = ctypes.CDLL(BVH_LIB_PATH)
    return _bvh


def ref_leaf_init(means3D, scales, rotations):
    """bvh/__init__.py:29-57 in PyTorch (build_rotation: utils/general_utils.py:82-103), verbatim op order."""
    P = means3D.shape[0]
    r = rotations
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    rot = torch.zeros((q.size(0), 3, 3), device=r.device)
    rr, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rot[:, 0, 0] = 1 - 2 * (y * y + z * z); rot[:, 0, 1] = 2 * (x * y - rr * z); rot[:, 0, 2] = 2 * (x * z + rr * y)
    rot[:, 1, 0] = 2 * (x * y + rr * z); rot[:, 1, 1] = 1 - 2 * (x * x + z * z); rot[:, 1, 2] = 2 * (y * z - rr * x)
    rot[:, 2, 0] = 2 * (x * z - rr * y); rot[:, 2, 1] = 2 * (y * z + rr * x); rot[:, 2, 2] = 1 - 2 * (x * x + y * y)
    nodes = torch.full((2 * P - 1, 5), -1, device=r.device).int()
    nodes[:P - 1, 4] = 0
    nodes[P - 1:, 4] = 1
    aabbs = torch.zeros(2 * P - 1, 6, device=r.device).float()
    aabbs[:, :3] = 100000
    aabbs[:, 3:] = -100000
    a, b, c = rot[:, :, 0], rot[:, :, 1], rot[:, :, 2]
    m = 3
    sa, sb, sc = m * scales[:, 0], m * scales[:, 1], m * scales[:, 2]
    xs = [means3D + s0 * a * sa[:, None] + s1 * b * sb[:, None] + s2 * c * sc[:, None]
          for s0 in (1, -1) for s1 in (1, -1) for s2 in (1, -1)]
    # (means +- a*sa) +- b*sb +- c*sc : multiplying by +-1 is exact, so this is the reference's rounding
    aabb_min = xs[0]
    aabb_max = xs[0]
    for t in xs[1:]:
        aabb_min = torch.minimum(aabb_min, t)
        aabb_max = torch.maximum(aabb_max, t)
    aabbs[P - 1:] = torch.cat([aabb_min, aabb_max], dim=-1)
    return nodes.contiguous(), aabbs.contiguous()


def ref_bvh_create(means3D, scales, rotations):
    nodes, aabbs = ref_leaf_init(means3D, scales, rotations)
    P = means3D.shape[0]
    morton = torch.zeros(P, dtype=torch.int64, device=means3D.device)
    torch.cuda.synchronize()
    rc = bvh_lib().ref_bvh_create(P, _p(means3D.contiguous()), _p(scales.contiguous()), _p(rotations.contiguous()),
                                  _p(nodes), _p(aabbs), _p(morton))
    if rc != 0:
        raise RuntimeError("reference create_bvh failed")
    return nodes, aabbs, morton


def ref_bvh_trace_opacity(nodes, aabbs, rays_o, rays_d, means3D, covs3D, opacities, normals):
    shape = rays_d.shape[:-1]
    n = rays_d.numel() // 3
    contrib = torch.zeros(shape, dtype=torch.int32, device=rays_d.device)
    opa = torch.ones(shape, dtype=torch.float32, device=rays_d.device)
    ts = [t.contiguous() for t in (rays_o, rays_d, means3D, covs3D, opacities, normals)]
    torch.cuda.synchronize()
    rc = bvh_lib().ref_bvh_trace_opacity(n, _p(nodes), _p(aabbs), *[_p(t) for t in ts], _p(contrib), _p(opa))
    if rc != 0:
        raise RuntimeError("reference trace_bvh_opacity failed")
    return contrib, opa
