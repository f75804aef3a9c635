#!/usr/bin/env python3
"""Generates tests/golden/adam_steps.npz by running the optimiser the reference trains with —
`torch.optim.Adam(l, lr=0.0, eps=1e-15)` with per-group learning rates, scene/gaussian_model.py:465-505
— on CPU in this container (torch 2.11; the reference pins 1.12.1, whose `_single_tensor_adam` differs
from 2.x only by `mul_().add_()` vs `lerp_()` for exp_avg: a last-ulp difference).  Inputs are
regenerated from the seeds by the test (tests/helpers.py:adam_case); the parameters, moments and
step counts after 7 steps are the fixture."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import adam_case  # noqa: E402

params, lrs, grads_for_step, lr_edits = adam_case()
tparams = [torch.from_numpy(p.copy()).requires_grad_(True) for p in params]
opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(tparams, lrs)], lr=0.0, eps=1e-15, foreach=False)
for step in range(1, 8):
    for gi, lr in lr_edits.get(step, []):
        opt.param_groups[gi]["lr"] = lr                     # update_learning_rate (gaussian_model.py:499-505)
    for p, g in zip(tparams, grads_for_step(step)):
        p.grad = torch.from_numpy(g.copy())
    opt.step()
out = {"torch_version": np.array(torch.__version__)}
for i, p in enumerate(tparams):
    st = opt.state[p]
    out[f"p{i}"] = p.detach().numpy()
    out[f"m{i}"] = st["exp_avg"].numpy()
    out[f"v{i}"] = st["exp_avg_sq"].numpy()
    out[f"step{i}"] = np.array(float(st["step"]))
np.savez_compressed(os.path.join(HERE, "adam_steps.npz"), **out)
print("wrote adam_steps.npz", {k: v.shape for k, v in out.items() if k[0] in "pmv"})
