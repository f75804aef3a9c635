"""Multi-GPU view parallelism: the model is replicated, rank r renders view r of
the step's batch, and ONE all-reduce over a flat per-Gaussian gradient buffer follows backward.

The reference has no multi-GPU code (single process, `cuda:0` pinned, utils/general_utils.py:167;
one view per iteration, train.py:115-127), so an N-view step is defined here as the mean of N
single-view losses: grad = (1/N) * sum_r grad_r.  The path shards over independent views with no
data-path collective other than this gradient sum.

`GradBucket` is pure torch (works on CPU tensors with the gloo backend — that is how the host
logic is tested without a GPU); on the GPU box it is backed by NCCL over NVLink and the backward
kernels write directly into its views (`_C_raster.rasterize_gaussians_backward(_out=...)`), so
no packing copy precedes the collective.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist

# per-Gaussian parameter gradients that are summed across views (means2D / colors / cov3D are
# per-view by-products: means2D feeds per-view densification statistics)
REDUCED = ("means3D", "features", "sh", "opacity", "scales", "rotations")


class GradBucket:
    def __init__(self, P, S, M, device, names=REDUCED):
        shapes = OrderedDict(means3D=(P, 3), features=(P, S), sh=(P, M, 3), opacity=(P, 1),
                             scales=(P, 3), rotations=(P, 4))
        self.names = tuple(n for n in names if n in shapes)
        self.shapes = OrderedDict((n, shapes[n]) for n in self.names)
        sizes = [int(torch.Size(s).numel()) for s in self.shapes.values()]
        # 128-float (512 B) aligned sections so every view starts on a cache-line boundary
        self.offsets, off = [], 0
        for n in sizes:
            self.offsets.append(off)
            off += (n + 127) // 128 * 128
        self.flat = torch.zeros(max(off, 1), dtype=torch.float32, device=device)
        self.views = OrderedDict()
        for (name, shape), o, n in zip(self.shapes.items(), self.offsets, sizes):
            self.views[name] = self.flat[o:o + n].view(shape)

    def bytes(self):
        return self.flat.numel() * 4

    def allreduce_mean(self, group=None, async_op=False):
        """One collective for the whole step.  Returns the work handle when async_op."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        if dist.get_backend(group) == "nccl":          # averaging is fused into the collective (no 2 x 256 MB pre-scale pass)
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        self.flat.mul_(1.0 / world)                    # gloo (CPU tests) has no AVG
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def view_for_rank(step, rank, world, num_views):
    """Views of a step are dealt round-robin: rank r of step s renders view (s*world + r) % V."""
    return (step * world + rank) % num_views


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env).
    Returns (rank, local_rank, world)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world
