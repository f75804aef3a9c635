"""Multi-GPU hardware gate (needs >= 2 GPUs on the box; skipped otherwise): the N-view step's exchanged
gradients equal (a) the dense all-reduce and (b) the mean of N single-view backward passes of the
reference's own kernels, to 1e-3 (SURVEY.md §8e definition of an N-view step)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_factored_exchange_equals_mean_of_reference_single_view_backwards():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_exchange_nccl.py"), "--P", "200000", "--W", "640", "--H", "480",
           "--iters", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["identical_on_all_ranks"]
    if res["rel_l2_vs_mean_of_reference_single_view_backwards"] is not None:
        assert res["rel_l2_vs_mean_of_reference_single_view_backwards"]["max_over_ranks"] < 1e-3
