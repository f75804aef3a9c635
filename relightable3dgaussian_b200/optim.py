"""Host-side mirror of the optimiser the reference trains with: `torch.optim.Adam(l, lr=0.0,
eps=1e-15)` over 7-13 per-Gaussian parameter groups (scene/gaussian_model.py:465-497), backed by
ONE fused sm_100a launch for all groups (r3dg_adam_step, csrc/adam.cu) instead of ~8 elementwise
kernels per group.

`FusedAdam` keeps torch.optim.Adam's observable behaviour for the configuration the reference
uses: same constructor arguments, `param_groups` (the reference edits `group['lr']` per step,
gaussian_model.py:499-505), and the same per-parameter state layout {step, exp_avg, exp_avg_sq} —
the densification surgery reads and replaces those entries (`_prune_optimizer`,
`cat_tensors_to_optimizer`, `replace_tensor_to_optimizer`, gaussian_model.py:667-750) and
checkpoints store `optimizer.state_dict()` (:129), so state dicts are interchangeable with
torch.optim.Adam in both directions.  weight_decay / amsgrad / maximize are not used by the
reference and are rejected.  There is no CPU path: parameters must be CUDA float32.
"""
import ctypes

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *,
                 maximize=False, **unsupported):
        if weight_decay != 0 or amsgrad or maximize:
            raise ValueError("FusedAdam implements the reference's configuration only: weight_decay=0, amsgrad=False, maximize=False")
        for k in ("foreach", "capturable", "differentiable", "fused"):      # accepted and ignored (torch >= 2 keyword noise)
            unsupported.pop(k, None)
        if unsupported:
            raise TypeError(f"unexpected arguments {sorted(unsupported)}")
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"Invalid beta parameters: {betas}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries, keep = [], []
        device = None
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("FusedAdam: parameters and gradients must be dense CUDA float32 tensors (no CPU path)")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous")
                device = device or p.device
                if p.device != device:
                    raise RuntimeError("FusedAdam: all parameters must live on one device")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                step = st["step"]
                if torch.is_tensor(step):
                    step += 1
                    t = int(step.item())
                else:                                   # state dicts written by torch < 1.12 hold a Python int
                    t = st["step"] = step + 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                m, v = st["exp_avg"], st["exp_avg_sq"]
                if not (m.is_contiguous() and v.is_contiguous()):
                    m = st["exp_avg"] = m.contiguous()
                    v = st["exp_avg_sq"] = v.contiguous()
                keep.append(g)
                e = _lib.AdamTensor()
                e.param, e.grad, e.exp_avg, e.exp_avg_sq = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                e.n, e.step = p.numel(), t
                e.lr, e.beta1, e.beta2, e.eps = float(group["lr"]), float(beta1), float(beta2), float(group["eps"])
                entries.append(e)
        if entries:
            arr = (_lib.AdamTensor * len(entries))(*entries)
            with torch.cuda.device(device):
                _lib.check(_lib.load().r3dg_adam_step(len(entries), arr, torch.cuda.current_stream(device).cuda_stream),
                           "FusedAdam.step")
        return loss


_torch_adam = None


def install():
    """Make the unmodified reference pick up the fused step: rebinds `torch.optim.Adam` (the name
    scene/gaussian_model.py:489 resolves at call time) to FusedAdam.  `uninstall()` restores it."""
    global _torch_adam
    if _torch_adam is None:
        _torch_adam = torch.optim.Adam
        torch.optim.Adam = FusedAdam
    return FusedAdam


def uninstall():
    global _torch_adam
    if _torch_adam is not None:
        torch.optim.Adam = _torch_adam
        _torch_adam = None
