"""§8(f)2 fused un-premultiply epilogue vs the reference's PyTorch expression (neilf.py:136-137), and the
deferred instance count: overflow of the speculative binning buffer is repaired inside backward."""
import warnings

import numpy as np
import pytest
import torch

from helpers import case_inputs, npy, rel_l2

pytestmark = pytest.mark.gpu


def test_unpremultiply_matches_pytorch_expression():
    from relightable3dgaussian_b200.rasterizer import unpremultiply
    g = torch.Generator().manual_seed(0)
    S, H, W = 16, 123, 77
    feat = torch.randn(S, H, W, generator=g).cuda().requires_grad_(True)
    opac = torch.rand(1, H, W, generator=g).cuda()
    opac[0, :5] = 0.0; opac[0, 5:8] = 5e-6                         # below the clamp
    opac.requires_grad_(True)
    ncon = torch.randint(0, 3, (H, W), generator=g).int().cuda()
    cot = torch.randn(S, H, W, generator=g).cuda()
    ref = feat / opac.clamp_min(1e-5) * (ncon > 0)
    (ref * cot).sum().backward()
    gf, go = feat.grad.clone(), opac.grad.clone()
    feat.grad = None; opac.grad = None
    out = unpremultiply(feat, opac, ncon)
    assert torch.equal(out, ref.detach())                          # same op order: bit-identical
    (out * cot).sum().backward()
    assert rel_l2(npy(feat.grad), npy(gf)) < 1e-6 and rel_l2(npy(opac.grad), npy(go)) < 1e-5


def test_deferred_count_overflow_is_repaired_in_backward():
    from relightable3dgaussian_b200 import _C_raster
    from relightable3dgaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer, set_deferred_count
    P, W, H, S = 3000, 160, 120, 3
    bg = torch.zeros(3).cuda()
    d = lambda t: t.cuda()

    def run(boost, view):
        sc, cam = case_inputs(P, W, H, S, view=view, scale_boost=boost)
        leaves = [d(t).requires_grad_(True) for t in (sc.means3D, sc.opacities, sc.shs, sc.scales, sc.rotations, sc.features)]
        rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, cam.cx, cam.cy, bg, 1.0, d(cam.viewmatrix), d(cam.projmatrix), 3,
                                           d(cam.campos), False, True, True, False)
        out = GaussianRasterizer(rs)(means3D=leaves[0], means2D=torch.zeros_like(leaves[0], requires_grad=True), opacities=leaves[1],
                                     shs=leaves[2], scales=leaves[3], rotations=leaves[4], features=leaves[5])
        g = torch.Generator().manual_seed(9)
        cots = [torch.randn(c, H, W, generator=g).cuda() for c in (3, 1, 1, S)]
        sum((o * c).sum() for o, c in zip((out[2], out[3], out[4], out[5]), cots)).backward()
        torch.cuda.synchronize()
        return out, [t.grad.clone() for t in leaves]

    _C_raster._state.clear()
    ref_out, ref_grads = run(20.0, 2)                               # synchronous path: the truth for the big view
    _C_raster._state.clear()
    set_deferred_count(True)
    old = _C_raster._HEADROOM
    try:
        _C_raster._HEADROOM = 1.0
        for i in range(_C_raster._LEARN + 1):                      # learn the counts on SMALL views
            out, _ = run(1.0, i % 8)
        assert isinstance(out[0], _C_raster.DeferredCount)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            out, grads = run(20.0, 2)                               # many more instances than anything seen: overflows
        assert out[0].overflowed and int(out[0]) == int(ref_out[0]) and any("re-run" in str(x.message) for x in w)
        for a, b in zip(grads, ref_grads):                          # linear loss: cotangents do not depend on the images
            assert rel_l2(npy(a), npy(b)) < 1e-3
        # a forward nobody differentiates takes the synchronous path even with deferral on
        with torch.no_grad():
            out, _ = (lambda: (GaussianRasterizer(GaussianRasterizationSettings(
                H, W, 1.0, 1.0, W / 2, H / 2, bg, 1.0, torch.eye(4).cuda(), torch.eye(4).cuda(), 3, torch.zeros(3).cuda(), False, True, True, False))(
                means3D=torch.zeros(4, 3).cuda(), means2D=torch.zeros(4, 3).cuda(), opacities=torch.ones(4, 1).cuda(),
                shs=torch.zeros(4, 16, 3).cuda(), scales=torch.ones(4, 3).cuda(), rotations=torch.ones(4, 4).cuda()), None))()
        assert isinstance(out[0], int)
    finally:
        _C_raster._HEADROOM = old
        set_deferred_count(False)
        _C_raster._state.clear()


def test_pack_features_matches_reference_expression():
    """rasterizer.pack_features == neilf.py:110-118: depths via the homogeneous matmul, squares, 8-tensor cat; gradients
    of every source and of means3D through the depth channels."""
    from relightable3dgaussian_b200.rasterizer import pack_features
    from relightable3dgaussian_b200 import synth
    g = torch.Generator().manual_seed(1)
    P = 10_007
    cam = synth.make_camera(3, 320, 200)
    view = cam.viewmatrix.cuda()
    leaf = lambda *s: torch.randn(*s, generator=g).cuda().requires_grad_(True)
    xyz, brdf, nrm, base, rough, diff, vis = leaf(P, 3), leaf(P, 3), leaf(P, 3), leaf(P, 3), leaf(P, 1), leaf(P, 3), leaf(P, 1)
    srcs = [brdf, nrm, base, rough, diff, vis]
    depths = (torch.cat([xyz, torch.ones_like(xyz[:, :1])], dim=-1) @ view)[:, 2:3]
    ref = torch.cat([depths, depths.square()] + srcs, dim=-1)
    cot = torch.randn(P, 16, generator=g).cuda()
    (ref * cot).sum().backward()
    want = [t.grad.clone() for t in [xyz] + srcs]
    for t in [xyz] + srcs:
        t.grad = None
    out = pack_features(srcs, means3D=xyz, viewmatrix=view)
    assert out.shape == (P, 16) and (out - ref.detach()).abs().max().item() <= 1e-5 and torch.equal(out[:, 2:], ref.detach()[:, 2:])
    (out * cot).sum().backward()
    for t, w in zip([xyz] + srcs, want):
        assert t.grad is not None and t.grad.is_contiguous() and rel_l2(npy(t.grad), npy(w)) < 1e-6
    # plain fused cat (no depth channels), one source without grad
    a, b = leaf(P, 2), torch.randn(P, 5, generator=g).cuda()
    out = pack_features([a, b])
    assert torch.equal(out, torch.cat([a, b], -1).detach())
    out.sum().backward()
    assert torch.equal(a.grad, torch.ones_like(a))
