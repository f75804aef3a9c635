"""Host-side mirror of the reference's operator surface for the rasterizer
(gaussian_renderer/r3dg_rasterization.py:32-262): `GaussianRasterizationSettings`,
`GaussianRasterizer`, `rasterize_gaussians`, with the same field names, argument meaning, return
tuple and error behaviour, on top of the B200 kernels (through `_C_raster`, our stand-in for
`r3dg_rasterization._C`).
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C_raster as _C


class GaussianRasterizationSettings(NamedTuple):
    # field order == gaussian_renderer/r3dg_rasterization.py:188-204
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    cx: float
    cy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    backward_geometry: bool
    computer_pseudo_normal: bool
    debug: bool


_DEFER_COUNT = False


def set_deferred_count(enabled: bool):
    """Opt-in host run-ahead (not part of the reference surface).  The reference forward blocks on
    a D2H copy of `num_rendered` (rasterizer_impl.cu:291); by default so do we, once, after the
    whole forward is enqueued.  With deferral on, a forward that will be differentiated (some input
    requires grad) returns `num_rendered` as a `_C_raster.DeferredCount` (int()-able) that is
    resolved in backward, after the backward kernels have been enqueued, so the loss and the
    backward follow the forward on the GPU without a host round trip in between.

    It cannot fail and cannot silently truncate: the first `_C_raster._LEARN` forwards of a
    (P, W, H) shape always take the synchronous path (they learn the instance counts); afterwards the
    binning buffer holds `_C_raster._HEADROOM` x the largest count seen; a view that still
    overflows is detected in backward, which then re-runs the forward with the exact size and
    differentiates that (the cotangents autograd hands over were computed from the truncated images
    of the first attempt — one warning is issued); forwards that are not differentiated
    (`torch.no_grad()`, eval renders) always take the synchronous path."""
    global _DEFER_COUNT
    _DEFER_COUNT = bool(enabled)


_EXCHANGE = None


def set_grad_exchange(exchange, campos_all=None):
    """Opt-in multi-GPU data parallelism over views (not part of the single-GPU reference surface;
    SURVEY.md §8e): the factorised exchange of the SH gradient inside `loss.backward()`.

    With a `dist.FactoredGradExchange` installed, the backward of `GaussianRasterizer` does not write
    the dense `dL_dsh` [P,M,3] at all: it writes this view's rank-1 factor [P,3], all-gathers the
    factors of the ranks' views, rebuilds mean_v(dL_dsh_v) locally and hands THAT to autograd as the
    cotangent of `shs`.  This is only valid for `shs` because the reference's render functions feed the
    rasterizer `pc.get_shs` — a view-independent (concatenation) function of the leaf parameters — so
    J^T(mean_v g_v) == mean_v(J^T g_v).  Every other cotangent (means3D, features, opacity, scales,
    rotations, ...) stays THIS VIEW's: `features` / `means3D` reach the leaves through view-dependent
    maps (depth channels, brdf_color(viewdirs): render.py:91, neilf.py:88-120), so they must be
    averaged at the leaf parameters after backward — `dist.LeafGradBucket` / `dist.average_leaf_grads`,
    skipping the SH leaves.  `campos_all` [world,3] holds the camera centres of the views the ranks
    render this step, in rank order (call again per step).  All ranks must take the same path:
    a backward that cannot use an installed exchange (debug mode, no `shs`) raises instead of silently
    skipping its collectives.  `set_grad_exchange(None)` turns it off."""
    global _EXCHANGE
    _EXCHANGE = None if exchange is None else (exchange, campos_all)


def _snapshot(args):
    return tuple(a.cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


class _RasterizeGaussians(torch.autograd.Function):
    """Autograd node (reference :58-185).  Inputs/outputs and the set of differentiable inputs
    are the reference's; cotangents of normal / surface_xyz / weights / n_contrib / radii are
    ignored exactly as there (:123-124 vs :132-157)."""

    @staticmethod
    def forward(ctx, means3D, means2D, features, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        args = (rs.bg, means3D, features, colors_precomp, opacities, scales, rotations,
                rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                rs.tanfovy, rs.cx, rs.cy, rs.image_height, rs.image_width, sh, rs.sh_degree,
                rs.campos, rs.prefiltered, rs.computer_pseudo_normal, rs.debug)
        if rs.debug:
            saved = _snapshot(args)
            try:
                out = _C.rasterize_gaussians(*args)
            except Exception:
                torch.save(saved, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
        else:
            out = _C.rasterize_gaussians(*args, _defer=_DEFER_COUNT and _CALL_GRAD_MODE and any(ctx.needs_input_grad))
        (num_rendered, num_contrib, color, opacity, depth, feature, normal, surface_xyz, weights,
         radii, geomBuffer, binningBuffer, imgBuffer) = out
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        # cotangents of outputs the reference's backward ignores (num_contrib, normal, surface_xyz, weights, radii)
        # stay None instead of being materialised as zero tensors (5 fill kernels per step); the four image
        # cotangents are filled in below when a loss does not touch one of them
        ctx.set_materialize_grads(False)
        ctx.out_shapes = (color.shape, opacity.shape, depth.shape, feature.shape)
        ctx.save_for_backward(colors_precomp, means3D, features, scales, rotations, cov3Ds_precomp,
                              radii, sh, geomBuffer, binningBuffer, imgBuffer, opacities)
        return (num_rendered, num_contrib, color, opacity, depth, feature, normal, surface_xyz,
                weights, radii)

    @staticmethod
    def backward(ctx, g_num_rendered, g_num_contrib, g_color, g_opacity, g_depth, g_feature,
                 g_normal, g_surface_xyz, g_weights, g_radii):
        rs = ctx.raster_settings
        (colors_precomp, means3D, features, scales, rotations, cov3Ds_precomp, radii, sh,
         geomBuffer, binningBuffer, imgBuffer, opacities) = ctx.saved_tensors
        zeros = lambda g, shape: g if g is not None else torch.zeros(shape, dtype=torch.float32, device=means3D.device)
        g_color, g_opacity = zeros(g_color, ctx.out_shapes[0]), zeros(g_opacity, ctx.out_shapes[1])
        g_depth, g_feature = zeros(g_depth, ctx.out_shapes[2]), zeros(g_feature, ctx.out_shapes[3])

        def bwd_args(geom, num_rendered, binning, img):
            return (rs.bg, means3D, features, radii, colors_precomp, scales, rotations,
                    rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                    rs.tanfovy, g_color, g_opacity, g_depth, g_feature, sh, rs.sh_degree, rs.campos,
                    geom, num_rendered, binning, img, rs.backward_geometry, rs.debug)
        args = bwd_args(geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer)
        ex = None
        if _EXCHANGE is not None:
            if rs.debug or sh.numel() == 0:
                raise RuntimeError("a gradient exchange is installed (set_grad_exchange) but this backward cannot take "
                                   "part in it (debug mode or no `shs` input): every rank must run the same collectives")
            ex = _EXCHANGE
        kw = {} if ex is None else {"_out": {"sh_factor": ex[0].factor}}
        if rs.debug:
            saved = _snapshot(args)
            try:
                grads = _C.rasterize_gaussians_backward(*args)
            except Exception:
                torch.save(saved, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise
        else:
            grads = _C.rasterize_gaussians_backward(*args, **kw)
        if isinstance(ctx.num_rendered, _C.DeferredCount):
            ctx.num_rendered.resolve()          # after the backward kernels are in flight
            if ctx.num_rendered.overflowed:     # rare: re-run the forward with the exact size, differentiate that
                import warnings
                warnings.warn(f"deferred rasterization overflowed its binning buffer ({int(ctx.num_rendered)} instances); "
                              "forward re-run inside backward (cotangents come from the truncated first attempt)")
                out = _C.rasterize_gaussians(rs.bg, means3D, features, colors_precomp, opacities, scales, rotations,
                                             rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                                             rs.tanfovy, rs.cx, rs.cy, rs.image_height, rs.image_width, sh, rs.sh_degree,
                                             rs.campos, rs.prefiltered, rs.computer_pseudo_normal, False,
                                             _min_capacity=int(ctx.num_rendered) + 4096)
                grads = _C.rasterize_gaussians_backward(*bwd_args(out[10], out[0], out[11], out[12]), **kw)
        if ex is not None:                      # SH gradient: gather the views' factors, rebuild the mean locally
            exchange, campos_all = ex
            exchange.gather_factors()
            g_sh_avg = exchange.rebuild_sh(means3D, campos_all, rs.sh_degree)
            grads = grads[:6] + (g_sh_avg,) + grads[7:]
        (g_means2D, g_colors_precomp, g_opacities, g_means3D, g_features, g_cov3Ds, g_sh, g_scales,
         g_rotations) = grads
        return (g_means3D, g_means2D, g_features, g_sh, g_colors_precomp, g_opacities, g_scales,
                g_rotations, g_cov3Ds, None)


_CALL_GRAD_MODE = True      # grad mode of the caller (inside autograd.Function.forward it is always off)


def rasterize_gaussians(means3D, means2D, features, sh, colors_precomp, opacities, scales,
                        rotations, cov3Ds_precomp, raster_settings):
    global _CALL_GRAD_MODE
    _CALL_GRAD_MODE = torch.is_grad_enabled()
    return _RasterizeGaussians.apply(means3D, means2D, features, sh, colors_precomp, opacities,
                                     scales, rotations, cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, features=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        have_sr = scales is not None or rotations is not None
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                (have_sr and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        if features is None:
            features = torch.empty_like(means3D[..., :0])
        return rasterize_gaussians(means3D, means2D, features, shs, colors_precomp, opacities,
                                   scales, rotations, cov3D_precomp, self.raster_settings)


class _Unpremultiply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature, opacity, num_contrib):
        from . import _lib
        lib = _lib.load()
        S, HW = feature.shape[0], feature.shape[1] * feature.shape[2]
        f, o = feature.detach().float().contiguous(), opacity.detach().float().contiguous()
        n = num_contrib.detach().to(torch.int32).contiguous()
        out = torch.empty_like(f)
        if S > 0 and HW > 0:
            with torch.cuda.device(f.device):
                _lib.check(lib.r3dg_unpremultiply_forward(S, HW, f.data_ptr(), o.data_ptr(), n.data_ptr(), out.data_ptr(),
                                                          torch.cuda.current_stream(f.device).cuda_stream), "unpremultiply")
        ctx.save_for_backward(f, o, n)
        return out

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        lib = _lib.load()
        f, o, n = ctx.saved_tensors
        S, HW = f.shape[0], f.shape[1] * f.shape[2]
        g = g.float().contiguous()
        d_f, d_o = torch.empty_like(f), torch.empty_like(o)
        with torch.cuda.device(f.device):
            _lib.check(lib.r3dg_unpremultiply_backward(S, HW, f.data_ptr(), o.data_ptr(), n.data_ptr(), g.data_ptr(), d_f.data_ptr(),
                                                       d_o.data_ptr(), torch.cuda.current_stream(f.device).cuda_stream),
                       "unpremultiply backward")
        return d_f, d_o, None


def unpremultiply(rendered_feature, rendered_opacity, num_contrib):
    """Optional fused form of the epilogue the reference's render functions apply to the
    rasterizer's feature image (gaussian_renderer/neilf.py:135-137, render.py:106-108):
        rendered_feature / rendered_opacity.clamp_min(1e-5) * (num_contrib > 0)
    one HBM pass forward and one backward instead of 3 + ~6 PyTorch kernels.  The reference's own files keep
    the PyTorch expression; callers opt in by calling this instead."""
    return _Unpremultiply.apply(rendered_feature, rendered_opacity, num_contrib)


class _PackFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, viewmatrix, *tensors):
        import ctypes
        from . import _lib
        lib = _lib.load()
        P = tensors[0].shape[0] if tensors else means3D.shape[0]
        dev = (tensors[0] if tensors else means3D).device
        srcs = [t.detach().float().contiguous().view(P, -1) for t in tensors]
        depth = means3D is not None
        S = (2 if depth else 0) + sum(t.shape[1] for t in srcs)
        out = torch.empty((P, S), dtype=torch.float32, device=dev)
        m = means3D.detach().float().contiguous() if depth else None
        v = viewmatrix.detach().float().contiguous() if depth else None
        arr = (_lib.PackSrc * max(len(srcs), 1))()
        for k, t in enumerate(srcs):
            arr[k].ptr, arr[k].width = t.data_ptr(), t.shape[1]
        if P > 0:
            with torch.cuda.device(dev):
                _lib.check(lib.r3dg_pack_features_forward(P, S, m.data_ptr() if depth else None, v.data_ptr() if depth else None, len(srcs), arr,
                                                          out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "pack_features")
        ctx.meta = (P, S, [tuple(t.shape) for t in tensors], [t.shape[1] for t in srcs], depth)
        ctx.save_for_backward(*([m, v] if depth else []))
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes
        from . import _lib
        lib = _lib.load()
        P, S, shapes, widths, depth = ctx.meta
        m, v = ctx.saved_tensors if depth else (None, None)
        dev = g.device
        g = g.float().contiguous()
        need = ctx.needs_input_grad
        outs = [torch.empty((P, w), dtype=torch.float32, device=dev) if need[2 + k] else None for k, w in enumerate(widths)]
        d_m = torch.empty((P, 3), dtype=torch.float32, device=dev) if depth and need[0] else None
        arr = (_lib.PackSrc * max(len(widths), 1))()
        for k, w in enumerate(widths):
            arr[k].ptr, arr[k].width = (outs[k].data_ptr() if outs[k] is not None else None), w
        if P > 0:
            with torch.cuda.device(dev):
                _lib.check(lib.r3dg_pack_features_backward(P, S, m.data_ptr() if depth else None, v.data_ptr() if depth else None, g.data_ptr(),
                                                           len(widths), arr, d_m.data_ptr() if d_m is not None else None,
                                                           torch.cuda.current_stream(dev).cuda_stream), "pack_features backward")
        return (d_m, None) + tuple(o.view(s) if o is not None else None for o, s in zip(outs, shapes))


def pack_features(tensors, means3D=None, viewmatrix=None):
    """Optional fused form of the feature pack the reference's render functions build in front of
    the rasterizer (gaussian_renderer/neilf.py:110-126, render.py:88-93):
        depths = (cat([means3D, 1]) @ viewmatrix)[:, 2:3];  features = cat([depths, depths.square(), *tensors], -1)
    (without means3D / viewmatrix: a plain fused cat).  One pass forward, one backward that writes every source
    gradient contiguous.  No gradient flows to `viewmatrix` (a camera constant in the reference)."""
    return _PackFeatures.apply(means3D, viewmatrix, *tensors)
