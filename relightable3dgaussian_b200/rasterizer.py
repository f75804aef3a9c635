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
    whole forward is enqueued.  With deferral on, `num_rendered` is returned as a
    `_C_raster.DeferredCount` (int()-able) that is resolved in backward, after the backward kernels
    have been enqueued, so the loss and the backward follow the forward on the GPU without a host
    round trip in between.  The binning buffer is then sized 1.5x the last count; an overflow
    raises at resolve time instead of being retried transparently.  The very first forward on a
    device (no count seen yet) always takes the synchronous path."""
    global _DEFER_COUNT
    _DEFER_COUNT = bool(enabled)


_EXCHANGE = None


def set_grad_exchange(exchange, campos_all=None):
    """Opt-in multi-GPU data parallelism over views (not part of the single-GPU reference surface;
    SURVEY.md §8e).  With a `dist.FactoredGradExchange` installed, the backward of
    `GaussianRasterizer` writes its per-Gaussian parameter gradients straight into the exchange
    buffers, runs the step's collectives (one all-reduce of the dense rest, one all-gather of the
    SH-gradient factors, the local rebuild) and hands autograd the gradients ALREADY AVERAGED over
    the ranks' views — `loss.backward()` is the whole step.  `campos_all` [world,3] holds the camera
    centres of the views the ranks render this step, in rank order (call again per step).
    `means2D` / `colors_precomp` / `cov3D` gradients stay per-view.  `set_grad_exchange(None)` turns it off."""
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
            out = _C.rasterize_gaussians(*args, _defer=_DEFER_COUNT)
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
                              radii, sh, geomBuffer, binningBuffer, imgBuffer)
        return (num_rendered, num_contrib, color, opacity, depth, feature, normal, surface_xyz,
                weights, radii)

    @staticmethod
    def backward(ctx, g_num_rendered, g_num_contrib, g_color, g_opacity, g_depth, g_feature,
                 g_normal, g_surface_xyz, g_weights, g_radii):
        rs = ctx.raster_settings
        (colors_precomp, means3D, features, scales, rotations, cov3Ds_precomp, radii, sh,
         geomBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
        zeros = lambda g, shape: g if g is not None else torch.zeros(shape, dtype=torch.float32, device=means3D.device)
        g_color, g_opacity = zeros(g_color, ctx.out_shapes[0]), zeros(g_opacity, ctx.out_shapes[1])
        g_depth, g_feature = zeros(g_depth, ctx.out_shapes[2]), zeros(g_feature, ctx.out_shapes[3])
        args = (rs.bg, means3D, features, radii, colors_precomp, scales, rotations,
                rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                rs.tanfovy, g_color, g_opacity, g_depth, g_feature, sh, rs.sh_degree, rs.campos,
                geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, rs.backward_geometry,
                rs.debug)
        if rs.debug:
            saved = _snapshot(args)
            try:
                grads = _C.rasterize_gaussians_backward(*args)
            except Exception:
                torch.save(saved, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise
        elif _EXCHANGE is not None and sh.numel() != 0 and scales.numel() != 0:
            ex, campos_all = _EXCHANGE
            grads = _C.rasterize_gaussians_backward(*args, _out=ex.views)
            ex.exchange(means3D, campos_all, rs.sh_degree)
            g = ex.grads
            grads = (grads[0], grads[1], g["opacity"], g["means3D"], g["features"], grads[5], g["sh"], g["scales"], g["rotations"])
        else:
            grads = _C.rasterize_gaussians_backward(*args)
        if isinstance(ctx.num_rendered, _C.DeferredCount):
            ctx.num_rendered.resolve()          # after the backward kernels are in flight
        (g_means2D, g_colors_precomp, g_opacities, g_means3D, g_features, g_cov3Ds, g_sh, g_scales,
         g_rotations) = grads
        return (g_means3D, g_means2D, g_features, g_sh, g_colors_precomp, g_opacities, g_scales,
                g_rotations, g_cov3Ds, None)


def rasterize_gaussians(means3D, means2D, features, sh, colors_precomp, opacities, scales,
                        rotations, cov3Ds_precomp, raster_settings):
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
