#!/usr/bin/env python3
"""Generates tests/golden/sh_eval.npz by IMPORTING the Python reference
(/root/reference/utils/sh_utils.py:eval_sh, lines 71-128) in this container — BASELINE.json
config #1 (10k random Gaussians, degree-3 SH -> RGB on CPU PyTorch).  /root/reference does not
exist on the GPU box, so the outputs are committed as a fixture; inputs are regenerated from the
seed by tests (tests/helpers.py:sh_case)."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import sh_case  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_sh_utils", "/root/reference/utils/sh_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
for deg in (0, 1, 2, 3):
    sh, dirs = sh_case(10_000 if deg == 3 else 1_000, seed=deg)
    res = ref.eval_sh(deg, sh, dirs)                        # [P,3]
    out[f"deg{deg}"] = res.numpy().astype(np.float32)
    if deg == 3:
        out["deg3_rgb"] = torch.clamp_min(res + 0.5, 0.0).numpy().astype(np.float32)   # render.py:76-77
sh4 = torch.randn(256, 3, 25, generator=torch.Generator().manual_seed(9))
d4 = torch.nn.functional.normalize(torch.randn(256, 3, generator=torch.Generator().manual_seed(10)), dim=-1)
out["deg4_small"] = ref.eval_sh(4, sh4, d4).numpy().astype(np.float32)
np.savez_compressed(os.path.join(HERE, "sh_eval.npz"), **out)
print("wrote sh_eval.npz", {k: v.shape for k, v in out.items()})
