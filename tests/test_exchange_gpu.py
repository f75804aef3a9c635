"""GPU parity of the factorised SH-gradient exchange: on ONE GPU, two views are
rendered in turn; the gradient the ranks of a 2-GPU step would end up with — mean of the dense
per-view dL_dsh — must equal what r3dg_sh_grad_from_factors rebuilds from the two views' factors,
and equal the numpy oracle; all other gradients are unchanged by asking for the factor."""
import numpy as np
import pytest
import torch

from helpers import case_inputs, npy, rel_l2
from test_raster_gpu import GRADS, dev

pytestmark = pytest.mark.gpu


def _backward(C, sc, cam, S, deg, cots, out=None):
    bg = torch.tensor([0.1, 0.2, 0.3]).cuda()
    E = torch.Tensor([])
    d = lambda t: t.cuda()
    feats = d(sc.features) if S else torch.empty((sc.means3D.shape[0], 0), device="cuda")
    W, H = cam.image_width, cam.image_height
    o = C.rasterize_gaussians(bg, d(sc.means3D), feats, E, d(sc.opacities), d(sc.scales), d(sc.rotations), 1.0, E, d(cam.viewmatrix),
                              d(cam.projmatrix), cam.tanfovx, cam.tanfovy, cam.cx, cam.cy, H, W, d(sc.shs), deg, d(cam.campos),
                              False, True, False)
    g = C.rasterize_gaussians_backward(bg, d(sc.means3D), feats, o[9], E, d(sc.scales), d(sc.rotations), 1.0, E, d(cam.viewmatrix),
                                       d(cam.projmatrix), cam.tanfovx, cam.tanfovy, *[dev(c) for c in cots], d(sc.shs), deg,
                                       d(cam.campos), o[10], o[0], o[11], o[12], True, False, _out=out)
    torch.cuda.synchronize()
    return dict(zip(GRADS, g)), o


@pytest.mark.parametrize("P,W,H,S,deg", [(20_000, 320, 200, 5, 3), (5_000, 128, 96, 0, 1), (300, 64, 48, 2, 0)])
def test_rebuilt_sh_gradient_matches_dense_mean_and_oracle(P, W, H, S, deg):
    from relightable3dgaussian_b200 import _C_raster as C
    from relightable3dgaussian_b200.dist import FactoredGradExchange
    from oracle import oracle
    rng = np.random.default_rng(11)
    views = (1, 4)
    dense_mean, factors, campos, dense_other = 0, [], [], []
    for v in views:
        sc, cam = case_inputs(P, W, H, S, view=v)
        cots = [rng.standard_normal((c, H, W)).astype(np.float32) for c in (3, 1, 1, S)]
        gd, o = _backward(C, sc, cam, S, deg, cots)                       # reference-shaped dense backward
        ex = FactoredGradExchange(P, S, 16, "cuda", world=1)
        gf, _ = _backward(C, sc, cam, S, deg, cots, out=ex.views)         # factor instead of dL_dsh
        assert gf["dL_dsh"].numel() == 0
        for k in GRADS:
            if k != "dL_dsh" and gd[k].numel():                           # nothing else changes (atomics: run-to-run summation order only)
                assert rel_l2(npy(gf[k]), npy(gd[k])) < 1e-5, k
        # one backward producing BOTH the dense tensor and the factor: the one-view "exchange" rebuilds the dense one
        both = dict(ex.views, sh=torch.empty((P, 16, 3), device="cuda"))
        gb, _ = _backward(C, sc, cam, S, deg, cots, out=both)
        ex.rebuild_sh(sc.means3D.cuda(), cam.campos.cuda().view(1, 3).contiguous(), deg)
        assert rel_l2(npy(ex.sh), npy(gb["dL_dsh"])) < 5e-7
        assert torch.equal(ex.sh == 0, gb["dL_dsh"] == 0)                 # culled rows / inactive degrees are exact zeros
        assert bool((ex.sh[o[9] == 0] == 0).all()) and bool((ex.sh[:, (deg + 1) ** 2:] == 0).all())
        gd = gb
        dense_mean = dense_mean + npy(gd["dL_dsh"]) / len(views)
        factors.append(ex.factor.clone()); campos.append(cam.campos.clone())
    # the 2-view step, rebuilt from the gathered factors
    ex2 = FactoredGradExchange(P, S, 16, "cuda", world=2)
    ex2.gathered.copy_(torch.stack(factors))                              # what all_gather_into_tensor delivers
    ex2.rebuild_sh(sc.means3D.cuda(), torch.stack(campos).cuda().contiguous(), deg)
    assert np.abs(dense_mean).max() > 0
    assert rel_l2(npy(ex2.sh), dense_mean) < 1e-6
    ref = oracle.sh_grad_from_factors(npy(sc.means3D), [npy(c) for c in campos], [npy(f) for f in factors], deg, 16, 0.5)
    np.testing.assert_allclose(npy(ex2.sh), ref, rtol=2e-5, atol=1e-6 * np.abs(ref).max())


def test_argument_checks():
    from relightable3dgaussian_b200 import _lib
    lib = _lib.load()
    assert lib.r3dg_sh_grad_from_factors(10, 3, 9, 1, None, None, None, 1.0, None, None) == -10001      # 16 coefficients need M >= 16
    assert lib.r3dg_sh_grad_from_factors(10, 3, 16, 0, None, None, None, 1.0, None, None) == -10001
    assert lib.r3dg_sh_grad_from_factors(10, 3, 16, 65, None, None, None, 1.0, None, None) == -10002
    assert lib.r3dg_sh_grad_from_factors(0, 3, 16, 2, None, None, None, 1.0, None, None) == 0
