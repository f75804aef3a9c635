"""GPU parity of the fused optimiser step (r3dg_adam_step through FusedAdam) against the numpy oracle
(oracle/oracle_adam.py, pinned to torch.optim.Adam on CPU) and against torch.optim.Adam on the same
device, with the reference's parameter-group layout, per-step lr edits and optimiser-state surgery
(scene/gaussian_model.py:465-505, 667-750)."""
import numpy as np
import pytest
import torch

from helpers import npy

pytestmark = pytest.mark.gpu

# reference group layout for P Gaussians (gaussian_model.py:465-487): name, trailing shape, lr
GROUPS = [("xyz", (3,), 1.6e-4), ("normal", (3,), 1e-3), ("rotation", (4,), 1e-3), ("scaling", (3,), 5e-3),
          ("opacity", (1,), 5e-2), ("f_dc", (1, 3), 2.5e-3), ("f_rest", (15, 3), 2.5e-3 / 20),
          ("base_color", (3,), 1e-2), ("roughness", (1,), 1e-2), ("incidents_dc", (1, 3), 2e-3),
          ("incidents_rest", (15, 3), 1e-4), ("visibility_dc", (1, 1), 2.5e-3), ("visibility_rest", (3, 1), 1.25e-4)]


def make(P, seed, cls, **kw):
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn((P,) + shp, generator=g).cuda().requires_grad_(True) for _, shp, _ in GROUPS]
    opt = cls([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(params, GROUPS)], lr=0.0, eps=1e-15)
    return params, opt


def grads_for(P, step, seed):
    g = torch.Generator().manual_seed(1000 * seed + step)
    out = []
    for i, (_, shp, _) in enumerate(GROUPS):
        t = torch.randn((P,) + shp, generator=g) * 1e-3
        t[::3] = 0.0                                       # Gaussians outside the view: exact zero gradient
        out.append(t.cuda())
    return out


@pytest.mark.parametrize("P", [1, 1365, 70_001])            # 1365*3 = 4095: one short of a 4096-element block
def test_matches_torch_adam_and_oracle(P):
    from relightable3dgaussian_b200.optim import FusedAdam
    from oracle import oracle_adam
    pa, oa = make(P, 0, FusedAdam)
    pb, ob = make(P, 0, torch.optim.Adam)
    p0 = [npy(p) for p in pa]
    orc = [(npy(p), np.zeros(p.shape, np.float32), np.zeros(p.shape, np.float32)) for p in pa]
    lrs = [lr for _, _, lr in GROUPS]
    for step in range(1, 7):
        if step == 4:                                       # update_learning_rate edits group['lr'] in place
            oa.param_groups[0]["lr"] = ob.param_groups[0]["lr"] = lrs[0] = 0.9e-4
        gr = grads_for(P, step, 0)
        for p, q, g in zip(pa, pb, gr):
            p.grad, q.grad = g.clone(), g.clone()
        oa.step(); ob.step()
        oa.zero_grad(); ob.zero_grad()
        orc = [oracle_adam.adam_step(o[0], npy(g), o[1], o[2], step, lr, eps=1e-15) for o, g, lr in zip(orc, gr, lrs)]
    for p, q, o, q0, (name, _, _) in zip(pa, pb, orc, p0, GROUPS):
        sa, sb = oa.state[p], ob.state[q]
        assert float(sa["step"]) == float(sb["step"]) == 6.0
        for ref_p, ref_m, ref_v, tol in ((npy(q), npy(sb["exp_avg"]), npy(sb["exp_avg_sq"]), 2e-6), (o[0], o[1], o[2], 1e-6)):
            # parameters are O(1): compare the accumulated update, tolerance = a few last-ulp differences of p
            np.testing.assert_allclose(npy(p), ref_p, rtol=1e-6, atol=2e-7, err_msg=name)
            assert np.linalg.norm(npy(p) - ref_p) <= 2e-5 * np.linalg.norm(ref_p - q0) + 1e-12, name
            np.testing.assert_allclose(npy(sa["exp_avg"]), ref_m, rtol=tol, atol=tol * float(np.abs(ref_m).max()), err_msg=name)
            np.testing.assert_allclose(npy(sa["exp_avg_sq"]), ref_v, rtol=tol, atol=1e-18, err_msg=name)
    # the oracle is the tighter bar: same operation order, only the division / sqrt roundings are shared
    assert all(np.array_equal(npy(oa.state[p]["exp_avg_sq"]), o[2]) for p, o in zip(pa, orc)), "exp_avg_sq is a pure mul/add chain: bit-exact"


def test_state_surgery_and_state_dict_interchange():
    """The densification code replaces parameters and edits exp_avg / exp_avg_sq through optimizer.state
    (gaussian_model.py:667-750); checkpoints carry optimizer.state_dict() (:129)."""
    from relightable3dgaussian_b200.optim import FusedAdam
    P = 5000
    pa, oa = make(P, 1, FusedAdam)
    pb, ob = make(P, 1, torch.optim.Adam)
    for step in (1, 2):
        for p, q, g in zip(pa, pb, grads_for(P, step, 1)):
            p.grad, q.grad = g.clone(), g.clone()
        oa.step(); ob.step()
    # _prune_optimizer (gaussian_model.py:682-698): mask parameter and both moments of every group
    mask = (torch.arange(P) % 5 != 0).cuda()
    for opt in (oa, ob):
        for group in opt.param_groups:
            st = opt.state.get(group["params"][0], None)
            st["exp_avg"] = st["exp_avg"][mask]
            st["exp_avg_sq"] = st["exp_avg_sq"][mask]
            del opt.state[group["params"][0]]
            group["params"][0] = torch.nn.Parameter(group["params"][0][mask].requires_grad_(True))
            opt.state[group["params"][0]] = st
    Pn = int(mask.sum())
    for ga, gb, g in zip(oa.param_groups, ob.param_groups, grads_for(Pn, 3, 1)):
        ga["params"][0].grad, gb["params"][0].grad = g.clone(), g.clone()
    oa.step(); ob.step()
    for ga, gb in zip(oa.param_groups, ob.param_groups):
        np.testing.assert_allclose(npy(ga["params"][0]), npy(gb["params"][0]), rtol=1e-6, atol=2e-7, err_msg=ga["name"])
    # torch.optim.Adam's checkpoint loads into FusedAdam and vice versa
    # (deepcopy = the torch.save / torch.load round trip: load_state_dict itself aliases same-device tensors)
    import copy
    pc, oc = make(Pn, 2, FusedAdam)
    oc.load_state_dict(copy.deepcopy(ob.state_dict()))
    pd, od = make(Pn, 2, torch.optim.Adam)
    od.load_state_dict(copy.deepcopy(oa.state_dict()))
    for gc, gd, gb, g in zip(oc.param_groups, od.param_groups, ob.param_groups, grads_for(Pn, 4, 1)):
        with torch.no_grad():
            gc["params"][0].copy_(gb["params"][0]); gd["params"][0].copy_(gb["params"][0])
        gc["params"][0].grad, gd["params"][0].grad, gb["params"][0].grad = g.clone(), g.clone(), g.clone()
    oc.step(); od.step(); ob.step()
    for gc, gd, gb in zip(oc.param_groups, od.param_groups, ob.param_groups):
        np.testing.assert_allclose(npy(gc["params"][0]), npy(gb["params"][0]), rtol=1e-6, atol=2e-7, err_msg=gc["name"])
        np.testing.assert_allclose(npy(gd["params"][0]), npy(gb["params"][0]), rtol=1e-6, atol=2e-7, err_msg=gd["name"])


def test_rejects_what_it_does_not_implement_and_skips_missing_grads():
    from relightable3dgaussian_b200.optim import FusedAdam
    p = torch.zeros(10, device="cuda", requires_grad=True)
    with pytest.raises(ValueError):
        FusedAdam([p], weight_decay=0.1)
    with pytest.raises(ValueError):
        FusedAdam([p], amsgrad=True)
    opt = FusedAdam([p], lr=1e-2)
    opt.step()                                              # no gradient anywhere: nothing to do, no state
    assert len(opt.state) == 0
    q = torch.zeros(10, requires_grad=True)                 # CPU parameter: no fallback
    q.grad = torch.ones(10)
    with pytest.raises(RuntimeError):
        FusedAdam([q]).step()
    # unaligned views take the scalar path
    base = torch.randn(4099, device="cuda")
    r = base[1:].detach().requires_grad_(True)
    r.grad = torch.randn(4098, device="cuda")
    ref = r.detach().clone().requires_grad_(True); ref.grad = r.grad.clone()
    FusedAdam([r], lr=1e-2).step(); torch.optim.Adam([ref], lr=1e-2).step()
    np.testing.assert_allclose(npy(r), npy(ref), rtol=1e-6, atol=2e-7)


def test_fused_compaction_matches_boolean_indexing():
    """§8(f)3: compact_rows / prune_optimizer == the reference's `tensor[mask]` chain (gaussian_model.py:682-729)."""
    from relightable3dgaussian_b200 import optim as O
    g = torch.Generator().manual_seed(0)
    for P in (1, 7, 2048, 2049, 100_003):
        keep = (torch.rand(P, generator=g) < 0.6).cuda()
        shapes = [(3,), (1,), (15, 3), (4,), ()]
        ts = [torch.randn((P,) + s, generator=g).cuda() for s in shapes] + [torch.randint(0, 100, (P,), generator=g, dtype=torch.int32).cuda()]
        outs = O.compact_rows(ts, keep)
        for t, o in zip(ts, outs):
            assert torch.equal(o, t[keep]), (P, t.shape)
    # all kept / none kept
    t = torch.randn(1000, 3).cuda()
    assert torch.equal(O.compact_rows([t], torch.ones(1000, dtype=torch.bool).cuda())[0], t)
    assert O.compact_rows([t], torch.zeros(1000, dtype=torch.bool).cuda())[0].shape == (0, 3)
    # optimiser surgery: same parameters / moments as the reference's per-group loop, FusedAdam keeps stepping
    P = 5000
    names = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]
    shapes = [(3,), (1, 3), (15, 3), (1,), (3,), (4,)]
    params = [torch.nn.Parameter(torch.randn((P,) + s, generator=g).cuda()) for s in shapes]
    env = torch.nn.Parameter(torch.randn(1, 16, 32, 3).cuda())               # not per-Gaussian: must be left alone
    opt = O.FusedAdam([{"params": [p], "lr": 1e-3, "name": n} for p, n in zip(params, names)] + [{"params": [env], "lr": 1e-2, "name": "env"}], lr=0.0, eps=1e-15)
    for p in params + [env]:
        p.grad = torch.randn_like(p)
    opt.step()
    keep = (torch.rand(P, generator=g) < 0.5).cuda()
    expect = {n: (p.data[keep].clone(), opt.state[p]["exp_avg"][keep].clone(), opt.state[p]["exp_avg_sq"][keep].clone()) for p, n in zip(params, names)}
    accum = torch.rand(P, 1).cuda()
    new, extras = O.prune_optimizer(opt, keep, extra=[accum])
    assert set(new) == set(names) and torch.equal(extras[0], accum[keep])
    for n in names:
        p = new[n]
        assert isinstance(p, torch.nn.Parameter) and torch.equal(p.data, expect[n][0])
        assert torch.equal(opt.state[p]["exp_avg"], expect[n][1]) and torch.equal(opt.state[p]["exp_avg_sq"], expect[n][2])
    assert opt.param_groups[-1]["params"][0] is env and env in opt.state
    for p in list(new.values()) + [env]:
        p.grad = torch.randn_like(p)
    opt.step()                                                                  # state shapes are consistent
