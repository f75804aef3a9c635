"""TEST INFRASTRUCTURE — not product code.

numpy/ctypes front-end of the CPU oracle (oracle_raster.c) plus a numpy restatement of
utils/sh_utils.py:eval_sh (reference lines 71-128, BASELINE.json config #1).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Parity status: pinned against the unmodified reference CUDA kernels
(oracle/_ref) and the Python reference's eval_sh — see tests/golden/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_fp = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    so = os.path.join(_HERE, "liboracle_raster.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_raster.c", "oracle_bvh.c", "build.sh")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["bash", os.path.join(_HERE, "build.sh")], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_bin.restype = ctypes.c_int64
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return int(lib().oracle_num_threads())


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def preprocess(means3D, scales, rotations, opacities, shs, viewmatrix, projmatrix, campos, W, H,
               tan_fovx, tan_fovy, sh_degree=3, scale_modifier=1.0, cov3D_precomp=None,
               colors_precomp=None):
    """forward.cu:156-258.  Returns dict of per-Gaussian intermediates."""
    P = means3D.shape[0]
    M = 0 if shs is None else shs.shape[1]
    out = dict(
        radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32),
        depths=np.zeros(P, np.float32), cov3D=np.zeros((P, 6), np.float32),
        conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32),
        clamped=np.zeros((P, 3), np.uint8), tiles_touched=np.zeros(P, np.uint32))
    a = [_f32(x) for x in (means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp,
                           viewmatrix, projmatrix, campos)]
    lib().oracle_preprocess(
        ctypes.c_int(P), ctypes.c_int(sh_degree), ctypes.c_int(M), _p(a[0]), _p(a[1]),
        ctypes.c_float(scale_modifier), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6]),
        _p(a[7]), _p(a[8]), _p(a[9]), ctypes.c_int(W), ctypes.c_int(H), ctypes.c_float(tan_fovx),
        ctypes.c_float(tan_fovy), _p(out["radii"]), _p(out["means2D"]), _p(out["depths"]),
        _p(out["cov3D"]), _p(out["conic_opacity"]), _p(out["rgb"]), _p(out["clamped"]),
        _p(out["tiles_touched"]))
    return out


def mark_visible(means3D, viewmatrix):
    P = means3D.shape[0]
    out = np.zeros(P, np.uint8)
    lib().oracle_mark_visible(ctypes.c_int(P), _p(_f32(means3D)), _p(_f32(viewmatrix)), _p(out))
    return out.astype(bool)


def bin_and_sort(pre, W, H):
    """rasterizer_impl.cu:287-327: offsets, keys, stable sort, tile ranges."""
    P = pre["radii"].shape[0]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    offsets = np.zeros(P, np.uint32)
    args = (ctypes.c_int(P), ctypes.c_int(W), ctypes.c_int(H), _p(pre["radii"]), _p(pre["means2D"]),
            _p(pre["depths"]), _p(pre["tiles_touched"]), _p(offsets))
    R = int(lib().oracle_bin(*args, None, None, None))
    keys = np.zeros(max(R, 1), np.uint64)
    plist = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((T, 2), np.uint32)
    lib().oracle_bin(*args, _p(keys), _p(plist), _p(ranges))
    return dict(num_rendered=R, point_offsets=offsets, keys=keys[:R], point_list=plist[:R],
                ranges=ranges)


def render_forward(pre, binned, colors, features, bg, W, H):
    """forward.cu:263-395."""
    P = pre["radii"].shape[0]
    S = 0 if features is None else features.shape[1]
    HW = H * W
    out = dict(final_T=np.zeros(HW, np.float32), n_contrib=np.zeros(HW, np.uint32),
               color=np.zeros((3, H, W), np.float32), opacity=np.zeros((1, H, W), np.float32),
               depth=np.zeros((1, H, W), np.float32), feature=np.zeros((S, H, W), np.float32),
               weights=np.zeros((P, 1), np.float32))
    feats = _f32(features) if S else np.zeros((P, 1), np.float32)
    lib().oracle_render_forward(
        ctypes.c_int(W), ctypes.c_int(H), ctypes.c_int(S), _p(binned["ranges"]),
        _p(binned["point_list"] if binned["num_rendered"] else np.zeros(1, np.uint32)),
        _p(pre["means2D"]), _p(pre["depths"]), _p(feats), _p(_f32(colors)),
        _p(pre["conic_opacity"]), _p(_f32(bg)), _p(out["final_T"]), _p(out["n_contrib"]),
        _p(out["color"]), _p(out["opacity"]), _p(out["depth"]), _p(out["feature"]),
        _p(out["weights"]))
    return out


def surface_normal(opacity, depth, viewmatrix, W, H, tan_fovx, tan_fovy, cx, cy):
    """forward.cu:398-491."""
    normal = np.zeros((3, H, W), np.float32)
    xyz = np.zeros((3, H, W), np.float32)
    lib().oracle_surface_normal(
        ctypes.c_int(W), ctypes.c_int(H), _p(_f32(viewmatrix)), ctypes.c_float(tan_fovx),
        ctypes.c_float(tan_fovy), ctypes.c_float(cx), ctypes.c_float(cy), _p(_f32(opacity)),
        _p(_f32(depth)), _p(normal), _p(xyz))
    return normal, xyz


def rasterize_forward(means3D, opacities, viewmatrix, projmatrix, campos, bg, W, H, tan_fovx,
                      tan_fovy, cx, cy, shs=None, colors_precomp=None, scales=None, rotations=None,
                      cov3D_precomp=None, features=None, sh_degree=3, scale_modifier=1.0,
                      computer_pseudo_normal=True):
    """Whole reference forward (rasterizer_impl.cu:199-380) on the CPU."""
    pre = preprocess(means3D, scales, rotations, opacities, shs, viewmatrix, projmatrix, campos, W,
                     H, tan_fovx, tan_fovy, sh_degree, scale_modifier, cov3D_precomp,
                     colors_precomp)
    binned = bin_and_sort(pre, W, H)
    colors = colors_precomp if colors_precomp is not None else pre["rgb"]
    img = render_forward(pre, binned, colors, features, bg, W, H)
    if computer_pseudo_normal:
        normal, xyz = surface_normal(img["opacity"], img["depth"], viewmatrix, W, H, tan_fovx,
                                     tan_fovy, cx, cy)
    else:
        normal = np.zeros((3, H, W), np.float32)
        xyz = np.zeros((3, H, W), np.float32)
    return dict(pre=pre, binned=binned, img=img, normal=normal, surface_xyz=xyz, colors=colors)


def rasterize_backward(fwd, means3D, viewmatrix, projmatrix, campos, bg, W, H, tan_fovx, tan_fovy,
                       dL_dcolor, dL_dopacity, dL_ddepth, dL_dfeature, shs=None, scales=None,
                       rotations=None, cov3D_precomp=None, features=None, sh_degree=3,
                       scale_modifier=1.0, backward_geometry=True):
    """Whole reference backward (rasterizer_impl.cu:384-491) on the CPU."""
    pre, binned, img = fwd["pre"], fwd["binned"], fwd["img"]
    P = means3D.shape[0]
    S = 0 if features is None else features.shape[1]
    M = 0 if shs is None else shs.shape[1]
    g = dict(dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 4), np.float32),
             dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
             dL_dfeatures=np.zeros((P, max(S, 0)), np.float32),
             dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
             dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
             dL_drotations=np.zeros((P, 4), np.float32))
    feats = _f32(features) if S else np.zeros((P, 1), np.float32)
    dfeat_out = g["dL_dfeatures"] if S else np.zeros((P, 1), np.float32)
    dpf = _f32(dL_dfeature) if S else np.zeros((1, H, W), np.float32)
    lib().oracle_render_backward(
        ctypes.c_int(W), ctypes.c_int(H), ctypes.c_int(S), _p(binned["ranges"]),
        _p(binned["point_list"] if binned["num_rendered"] else np.zeros(1, np.uint32)),
        _p(_f32(bg)), _p(pre["means2D"]), _p(pre["depths"]), _p(pre["conic_opacity"]),
        _p(_f32(fwd["colors"])), _p(feats), _p(img["final_T"]), _p(img["n_contrib"]),
        _p(_f32(dL_dcolor)), _p(_f32(dL_dopacity)), _p(_f32(dL_ddepth)), _p(dpf),
        ctypes.c_int(1 if backward_geometry else 0), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
        _p(g["dL_dopacity"]), _p(g["dL_dcolors"]), _p(dfeat_out))
    cov3D = _f32(cov3D_precomp) if cov3D_precomp is not None else pre["cov3D"]
    lib().oracle_preprocess_backward(
        ctypes.c_int(P), ctypes.c_int(sh_degree), ctypes.c_int(M), _p(_f32(means3D)),
        _p(pre["radii"]), _p(_f32(shs)), _p(pre["clamped"]), _p(_f32(scales)), _p(_f32(rotations)),
        ctypes.c_float(scale_modifier), _p(cov3D), _p(_f32(viewmatrix)), _p(_f32(projmatrix)),
        ctypes.c_int(W), ctypes.c_int(H), ctypes.c_float(tan_fovx), ctypes.c_float(tan_fovy),
        _p(_f32(campos)), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dcolors"]),
        _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]),
        _p(g["dL_drotations"]))
    return g


# ----------------------------------------------------------------------------------------------
# utils/sh_utils.py:71-128 eval_sh — BASELINE.json config #1 (degree-3 SH -> RGB on the CPU).
# sh: [..., C, (deg+1)^2], dirs: [..., 3] -> [..., C].  Evaluated in float32 in the same
# left-to-right order as the reference's torch expression.
# ----------------------------------------------------------------------------------------------
_C0 = np.float32(0.28209479177387814)
_C1 = np.float32(0.4886025119029199)
_C2 = np.array([1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
                -1.0925484305920792, 0.5462742152960396], np.float32)
_C3 = np.array([-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                -0.4570457994644658, 1.445305721320277, -0.5900435899266435], np.float32)
_C4 = np.array([2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892,
                0.10578554691520431, -0.6690465435572892, 0.47308734787878004,
                -1.7701307697799304, 0.6258357354491761], np.float32)


def sh_basis(deg, dirs):
    """Real SH basis values [..., (deg+1)^2] with the reference's constants and signs
    (auxiliary.h:22-39 == utils/sh_utils.py:71-128), float32."""
    d = np.asarray(dirs, np.float32)
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    f = np.float32
    w = [np.full(x.shape, 0.28209479177387814, np.float32)]
    if deg > 0:
        w += [f(-0.4886025119029199) * y, f(0.4886025119029199) * z, f(-0.4886025119029199) * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        w += [f(1.0925484305920792) * xy, f(-1.0925484305920792) * yz, f(0.31539156525252005) * (f(2) * zz - xx - yy),
              f(-1.0925484305920792) * xz, f(0.5462742152960396) * (xx - yy)]
    if deg > 2:
        w += [f(-0.5900435899266435) * y * (f(3) * xx - yy), f(2.890611442640554) * xy * z,
              f(-0.4570457994644658) * y * (f(4) * zz - xx - yy), f(0.3731763325901154) * z * (f(2) * zz - f(3) * xx - f(3) * yy),
              f(-0.4570457994644658) * x * (f(4) * zz - xx - yy), f(1.445305721320277) * z * (xx - yy),
              f(-0.5900435899266435) * x * (xx - f(3) * yy)]
    return np.stack(w, axis=-1).astype(np.float32)


def sh_grad_factor(bwd, fwd):
    """The per-view rank-1 factor of dL_dsh: dL_dRGB gated by the forward's clamp flags
    (backward.cu:20-139 `dL_dRGB.x *= clamped[3*idx+0] ? 0 : 1`), zero for culled Gaussians."""
    pre = fwd["pre"]
    gate = (pre["clamped"].reshape(-1, 3) == 0) & (pre["radii"].reshape(-1, 1) > 0)
    return (bwd["dL_dcolors"] * gate).astype(np.float32)


def sh_grad_from_factors(means3D, campos_all, factors, deg, M, scale):
    """Restatement of the multi-GPU exchange identity (include/r3dg_b200.h: r3dg_sh_grad_from_factors):
    dL_dsh[g,k,:] = scale * sum_v basis_k(normalize(mean_g - campos_v)) * factors[v][g,:]; the
    per-view term is backward.cu:20-139's `dL_dsh[k] = basis_k * dL_dRGB`."""
    means3D = np.asarray(means3D, np.float32)
    P = means3D.shape[0]
    out = np.zeros((P, M, 3), np.float32)
    n = (deg + 1) ** 2
    for v in range(len(campos_all)):
        d = means3D - np.asarray(campos_all[v], np.float32)[None]
        d = d / np.sqrt((d * d).sum(-1, keepdims=True), dtype=np.float32)
        out[:, :n] += sh_basis(deg, d)[:, :, None] * np.asarray(factors[v], np.float32)[:, None, :]
    return (out * np.float32(scale)).astype(np.float32)


def eval_sh(deg, sh, dirs):
    assert 0 <= deg <= 4
    sh = np.asarray(sh, np.float32)
    dirs = np.asarray(dirs, np.float32)
    assert sh.shape[-1] >= (deg + 1) ** 2
    result = _C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = (result - _C1 * y * sh[..., 1] + _C1 * z * sh[..., 2] - _C1 * x * sh[..., 3])
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + _C2[0] * xy * sh[..., 4] + _C2[1] * yz * sh[..., 5] +
                      _C2[2] * (np.float32(2.0) * zz - xx - yy) * sh[..., 6] +
                      _C2[3] * xz * sh[..., 7] + _C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + _C3[0] * y * (3 * xx - yy) * sh[..., 9] +
                          _C3[1] * xy * z * sh[..., 10] +
                          _C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] +
                          _C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] +
                          _C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] +
                          _C3[5] * z * (xx - yy) * sh[..., 14] +
                          _C3[6] * x * (xx - 3 * yy) * sh[..., 15])
                if deg > 3:
                    result = (result + _C4[0] * xy * (xx - yy) * sh[..., 16] +
                              _C4[1] * yz * (3 * xx - yy) * sh[..., 17] +
                              _C4[2] * xy * (7 * zz - 1) * sh[..., 18] +
                              _C4[3] * yz * (7 * zz - 3) * sh[..., 19] +
                              _C4[4] * (zz * (35 * zz - 30) + 3) * sh[..., 20] +
                              _C4[5] * xz * (7 * zz - 3) * sh[..., 21] +
                              _C4[6] * (xx - yy) * (7 * zz - 1) * sh[..., 22] +
                              _C4[7] * xz * (xx - 3 * yy) * sh[..., 23] +
                              _C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy)) * sh[..., 24])
    return result.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# BVH visibility (oracle_bvh.c): leaf boxes, LBVH build, opacity trace
# ----------------------------------------------------------------------------------------------
def bvh_build(means3D, scales, rotations):
    """bvh/__init__.py:29-59 + construct.cu:147-266.  Returns nodes i32[2P-1,5], aabbs f32[2P-1,6],
    morton u64[P]."""
    P = means3D.shape[0]
    nodes = np.full((2 * P - 1, 5), -1, np.int32)
    nodes[:P - 1, 4] = 0
    nodes[P - 1:, 4] = 1
    aabbs = np.zeros((2 * P - 1, 6), np.float32)
    aabbs[:, :3] = 100000
    aabbs[:, 3:] = -100000
    leaf = np.zeros((P, 6), np.float32)
    lib().oracle_bvh_leaf_aabbs(ctypes.c_int(P), _p(_f32(means3D)), _p(_f32(scales)), _p(_f32(rotations)), _p(leaf))
    aabbs[P - 1:] = leaf
    morton = np.zeros(P, np.uint64)
    lib().oracle_bvh_build(ctypes.c_int(P), _p(nodes), _p(aabbs), _p(morton))
    return nodes, aabbs, morton


def bvh_trace_opacity(nodes, aabbs, rays_o, rays_d, means3D, covs3D, opacities, normals):
    """trace.cu:196-287.  rays_o / rays_d [..., 3] contiguous; returns (contribute i32[...], opacity f32[...])."""
    shape = rays_d.shape[:-1]
    n = int(np.prod(shape)) if len(shape) else 1
    contrib = np.zeros(shape, np.int32)
    opa = np.ones(shape, np.float32)
    lib().oracle_bvh_trace_opacity(ctypes.c_int64(n), _p(np.ascontiguousarray(nodes, np.int32)), _p(_f32(aabbs)),
                                   _p(_f32(rays_o)), _p(_f32(rays_d)), _p(_f32(means3D)), _p(_f32(covs3D)),
                                   _p(_f32(opacities)), _p(_f32(normals)), _p(contrib), _p(opa))
    return contrib, opa
