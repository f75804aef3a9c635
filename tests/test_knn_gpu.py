"""simple_knn._C.distCUDA2 drop-in: exact 3-NN mean squared distance vs brute force."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P", [5, 1000, 40_000])
def test_dist2_matches_bruteforce(P):
    from relightable3dgaussian_b200._C_knn import distCUDA2
    g = torch.Generator().manual_seed(P)
    pts = torch.randn(P, 3, generator=g).cuda()
    if P >= 1000:
        pts[: P // 4] *= 0.01           # a dense cluster: many boxes must be pruned / scanned
    got = distCUDA2(pts)
    ref = torch.empty(P, device="cuda")
    for s in range(0, P, 4096):
        d2 = torch.cdist(pts[s:s + 4096].double(), pts.double()).square()
        d2[torch.arange(d2.shape[0]), torch.arange(s, s + d2.shape[0])] = float("inf")
        ref[s:s + 4096] = d2.topk(min(3, P - 1), largest=False).values.float().sum(-1) / 3.0
    assert got.shape == (P,) and got.dtype == torch.float32
    torch.testing.assert_close(got, ref, rtol=2e-5, atol=1e-12)


def test_dropin_modules_importable():
    import importlib, os, sys
    from helpers import ROOT
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    assert callable(importlib.import_module("simple_knn._C").distCUDA2)
    bt = importlib.import_module("bvh_tracing")
    assert callable(bt._C.create_bvh) and callable(bt._C.trace_bvh_opacity) and callable(bt._C.trace_bvh)
