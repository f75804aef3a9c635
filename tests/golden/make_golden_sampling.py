#!/usr/bin/env python3
"""Generates tests/golden/sampling.npz from the REFERENCE's own functions (run in the build container, where
/root/reference exists): utils/sh_utils.py:rotation_between_z and utils/graphics_utils.py:fibonacci_sphere_sampling
are imported from the reference tree and executed on the CPU.  The reference hard-codes device="cuda" in
rotation_between_z (sh_utils.py:45,65); the import below hands those modules a `torch` proxy that drops that one
keyword so the unmodified function bodies run here without a GPU."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("R3DG_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampling.npz")


class _TorchCPU(types.ModuleType):
    """`torch` with factory functions that ignore device="cuda"."""

    def __getattr__(self, name):
        attr = getattr(torch, name)
        if name in ("zeros", "eye", "ones", "arange", "rand", "empty", "tensor"):
            def f(*a, **k):
                if k.get("device") == "cuda":
                    k["device"] = "cpu"
                return attr(*a, **k)
            return f
        return attr


def load(path, name, extra=None):
    src = open(path).read().replace("from .sh_utils import", "from ref_sh_utils import")
    mod = types.ModuleType(name)
    mod.__dict__["__name__"] = name
    sys.modules[name] = mod
    code = compile(src, path, "exec")
    ns = mod.__dict__
    exec(code, ns)
    ns["torch"] = _TorchCPU("torch")
    return mod


def main():
    sh = load(os.path.join(REF, "utils", "sh_utils.py"), "ref_sh_utils")
    gu = load(os.path.join(REF, "utils", "graphics_utils.py"), "ref_graphics_utils")
    gu.rotation_between_z = sh.rotation_between_z
    g = torch.Generator().manual_seed(0)
    n = torch.nn.functional.normalize(torch.randn(257, 3, generator=g), dim=-1)
    n[0] = torch.tensor([0.0, 0.0, -1.0]); n[1] = torch.tensor([0.0, 0.0, 1.0]); n[2] = torch.tensor([1.0, 0.0, 0.0])
    out = {"normals": n.numpy(), "R": sh.rotation_between_z(n).numpy()}
    for N in (24, 32, 100):
        d, a = gu.fibonacci_sphere_sampling(n, N, random_rotate=False)
        out[f"dirs_{N}"], out[f"areas_{N}"] = d.numpy(), a.numpy()
    torch.manual_seed(3)
    phase = torch.rand(257, 1)
    torch.manual_seed(3)
    d, _ = gu.fibonacci_sphere_sampling(n, 32, random_rotate=True)
    out["phase"], out["dirs_32_random"] = phase.numpy(), d.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
