"""TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py baselines) — never imported by the product.

CPU restatement (numpy float32) of the optimiser step the reference trains with:
`torch.optim.Adam(l, lr=0.0, eps=1e-15).step()` (scene/gaussian_model.py:489,495-497).  The
arithmetic lives in a third-party dependency that is not under /root/reference: torch, pinned to
1.12.1 by the reference (readme.md:26).  Its published algorithm, torch/optim/adam.py
`_single_tensor_adam` (weight_decay=0, amsgrad=False, maximize=False), per parameter tensor:

    step += 1
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bias_correction1 = 1 - beta1 ** step ;  bias_correction2 = 1 - beta2 ** step      (Python floats)
    step_size = lr / bias_correction1
    denom = (exp_avg_sq.sqrt() / math.sqrt(bias_correction2)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-step_size)

Pinned by tests/test_oracle_cpu.py against torch.optim.Adam executed here on CPU (the installed
torch 2.11 computes exp_avg with lerp_, a last-ulp difference; tolerance 2e-6 relative on the
update).
"""
import math

import numpy as np

f32 = np.float32


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One update; `step` is the 1-based count of THIS update.  Returns new (param, exp_avg, exp_avg_sq)."""
    p, g, m, v = (np.asarray(a, dtype=f32) for a in (param, grad, exp_avg, exp_avg_sq))
    m = m * f32(beta1) + g * f32(1 - beta1)
    v = v * f32(beta2) + (f32(1 - beta2) * g) * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / f32(math.sqrt(bc2)) + f32(eps)
    p = p + f32(-step_size) * (m / denom)
    return p.astype(f32), m.astype(f32), v.astype(f32)
