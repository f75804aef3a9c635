#!/usr/bin/env python3
"""Generates tests/golden/bvh_*.npz on the GPU box from the UNMODIFIED reference BVH kernels
(oracle/_ref/libref_bvh.so: bvh/src/construct.cu [one-line patched copy, see oracle/build_ref.sh]
+ bvh/src/trace.cu) driven exactly like bvh/__init__.py / scene/gaussian_model.py:312-342:

    gpurun -- 'python tests/golden/make_golden_bvh.py gpurun_out/golden'
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import bvh_case, npy  # noqa: E402
from oracle import ref_gpu  # noqa: E402

CASES = {"cube": ("cube-v1", 600, 48, 24, 6.0), "shell": ("shell-v1", 500, 40, 32, 8.0)}


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, (recipe, P, n_src, N, boost) in CASES.items():
        c = bvh_case(recipe, P, n_src, N, boost)
        d = {k: v.cuda() for k, v in c.items()}
        nodes, aabbs, morton = ref_gpu.ref_bvh_create(d["means3D"], d["scales"], d["rotations"])
        contrib, opa = ref_gpu.ref_bvh_trace_opacity(nodes, aabbs, d["rays_o"], d["rays_d"], d["means3D"], d["inv_cov"],
                                                     d["opacity"], d["normals"])
        torch.cuda.synchronize()
        rec = {"in_" + k: npy(v) for k, v in c.items()}
        rec.update(nodes=npy(nodes), aabbs=npy(aabbs), morton=npy(morton), contribute=npy(contrib), visibility=npy(opa))
        path = os.path.join(outdir, f"bvh_{name}.npz")
        np.savez_compressed(path, **rec)
        print(f"wrote {path}: P={P}, rays={contrib.numel()}, blocked={(opa == 0).float().mean().item():.3f}, "
              f"mean contrib={contrib.float().mean().item():.2f}, {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
