"""Host-side mirror of the reference's BVH operator surface — `RayTracer` (bvh/__init__.py:28-71),
`sample_incident_rays` / `fibonacci_sphere_sampling` (scene/gaussian_model.py:20-28,
utils/graphics_utils.py:9-37, utils/sh_utils.py:36-68) and the visibility bake `update_visibility`
(scene/gaussian_model.py:312-342) — as thin marshalling over the B200 kernels.  No sampling or
tracing arithmetic lives in this file: directions come from `r3dg_sample_incident_dirs`, the bake
from `r3dg_bvh_bake_visibility` (directions generated inside the trace kernel, written once).

Differences from the reference that are not observable in the results:
  * the ~60 small PyTorch kernels of RayTracer.__init__ are one fused kernel (same fp32 rounding);
  * trace_visibility does not materialise `rays_o + 0.05 * rays_d` nor the expanded origins;
  * update_visibility is ONE launch over all Gaussians in Morton order instead of the reference's
    chunk loop (the chunks only bound the size of its [chunk,N,3] PyTorch temporaries); the
    `[P,N,3]` ray-origin / ray-direction inputs of the reference's trace call never exist.
"""
import torch

from . import _C_bvh as _C
from . import _lib


def _f32(t):
    return t.detach().float().contiguous()


def fibonacci_sphere_sampling(normals, sample_num, random_rotate=True):
    """utils/graphics_utils.py:9-37: (incident_dirs [...,N,3], incident_areas [...,N,1]).  With
    random_rotate the per-Gaussian phase is drawn here with torch.rand (as the reference does) and
    applied in the kernel."""
    lib = _lib.load()
    pre_shape = normals.shape[:-1]
    n = _f32(normals).reshape(-1, 3)
    P, N = n.shape[0], int(sample_num)
    if not n.is_cuda:
        raise RuntimeError("fibonacci_sphere_sampling runs on the GPU only (no CPU path)")
    dirs = torch.empty((P, N, 3), dtype=torch.float32, device=n.device)
    areas = torch.empty((P, N, 1), dtype=torch.float32, device=n.device)
    phase = torch.rand(P, device=n.device) if random_rotate else None
    if P > 0 and N > 0:
        with torch.cuda.device(n.device):
            _lib.check(lib.r3dg_sample_incident_dirs(P, N, n.data_ptr(), None if phase is None else phase.data_ptr(),
                                                     dirs.data_ptr(), areas.data_ptr(),
                                                     torch.cuda.current_stream(n.device).cuda_stream), "sample_incident_dirs")
    return dirs.reshape(*pre_shape, N, 3), areas.reshape(*pre_shape, N, 1)


def sample_incident_rays(normals, is_training=False, sample_num=24):
    """scene/gaussian_model.py:20-28."""
    return fibonacci_sphere_sampling(normals, sample_num, random_rotate=bool(is_training))


def inverse_covariance(scaling, rotation, scaling_modifier=1.0):
    """`GaussianModel.get_inverse_covariance` (scene/gaussian_model.py:257-260): the symmetric 6-vector of
    L L^T with L = R(q) diag(1 / (s / modifier))  (utils/general_utils.py:66-79,151-160).  Init-time host
    PyTorch in the reference too (a handful of [P,3,3] ops, once per bake)."""
    q = torch.nn.functional.normalize(rotation.float(), dim=-1)
    r, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).view(-1, 3, 3)
    L = R * ((1.0 / scaling.float()) * (1.0 / scaling_modifier)).unsqueeze(1)
    cov = L @ L.transpose(1, 2)
    return torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], dim=-1).contiguous()


class RayTracer:
    """bvh/__init__.py:28-71: builds the LBVH in __init__, `trace_visibility` bakes T in {0} U [0.9,1]."""

    def __init__(self, means3D, scales, rotations):
        lib = _lib.load()
        P = means3D.shape[0]
        dev = means3D.device
        nodes = torch.empty((2 * P - 1, 5), dtype=torch.int32, device=dev)
        aabbs = torch.empty((2 * P - 1, 6), dtype=torch.float32, device=dev)
        m, s, r = _f32(means3D), _f32(scales), _f32(rotations)
        stream = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.r3dg_bvh_leaf_aabbs(P, m.data_ptr(), s.data_ptr(), r.data_ptr(), nodes.data_ptr(),
                                               aabbs.data_ptr(), stream.cuda_stream), "RayTracer leaf boxes")
        self.tree, self.aabb, self.morton = _C.create_bvh(m, s, r, nodes, aabbs)

    @torch.no_grad()
    def trace_visibility(self, rays_o, rays_d, means3D, symm_inv, opacity, normals):
        # origins that are an expand() over the sample dimension are passed un-expanded
        group = 1
        if rays_o.dim() >= 2 and rays_o.shape == rays_d.shape and rays_o.stride(-2) == 0 and rays_d.is_contiguous():
            group = rays_o.shape[-2]
            rays_o = rays_o[..., 0, :]
        cotrib, opa = _C._trace(self.tree, self.aabb, rays_o, group, 0.05, rays_d, means3D, symm_inv, opacity, normals)
        return {"visibility": opa.unsqueeze(-1), "contribute": cotrib.unsqueeze(-1)}

    @torch.no_grad()
    def bake_visibility(self, means3D, symm_inv, opacity, normals, sample_num, first_slot=0, count=None, out=None,
                        want_contribute=False, write_dirs=True):
        """The whole of the reference's bake loop body for leaf slots [first_slot, first_slot + count) in one
        launch (r3dg_bvh_bake_visibility).  Returns dict(visibility [P,N,1], incident_dirs [P,N,3],
        incident_areas [P,N,1][, contribute]); rows of Gaussians outside the slot range keep the content of
        `out` (zero-initialised when not given) — that is what the sharded bake sums over the ranks."""
        lib = _lib.load()
        P, N = means3D.shape[0], int(sample_num)
        dev = means3D.device
        count = P - first_slot if count is None else count
        f = dict(dtype=torch.float32, device=dev)
        if out is None:
            partial = first_slot != 0 or count != P
            mk = torch.zeros if partial else torch.empty
            out = {"visibility": mk((P, N, 1), **f)}
            if write_dirs:
                out.update(incident_dirs=mk((P, N, 3), **f), incident_areas=mk((P, N, 1), **f))
            if want_contribute:
                out["contribute"] = mk((P, N, 1), dtype=torch.int32, device=dev)
        m, ci, op, nr = _f32(means3D), _f32(symm_inv), _f32(opacity), _f32(normals)
        tmp = torch.empty((lib.r3dg_bvh_trace_tmp_bytes(P),), dtype=torch.uint8, device=dev)
        contrib = out.get("contribute")
        if P > 0 and N > 0 and count > 0:
            with torch.cuda.device(dev):
                _lib.check(lib.r3dg_bvh_bake_visibility(
                    P, int(first_slot), int(count), N, self.tree.data_ptr(), self.aabb.data_ptr(), m.data_ptr(), ci.data_ptr(),
                    op.data_ptr(), nr.data_ptr(), 0.05, None if contrib is None else contrib.data_ptr(),
                    out["visibility"].data_ptr(), out["incident_dirs"].data_ptr() if "incident_dirs" in out else None,
                    out["incident_areas"].data_ptr() if "incident_areas" in out else None,
                    tmp.data_ptr(), tmp.numel(), torch.cuda.current_stream(dev).cuda_stream), "bake_visibility")
        return out


@torch.no_grad()
def update_visibility(xyz, scaling, rotation, inverse_covariance, opacity, normal, sample_num, shard_group=None):
    """scene/gaussian_model.py:312-342 as a free function of the activated Gaussian tensors.
    Returns (visibility [P,N,1], incident_dirs [P,N,3], incident_areas [P,N,1]).

    `shard_group` (a torch.distributed process group, or True for the default group) shards the one-off
    bake over the ranks: the model is replicated, so every rank builds the same LBVH,
    traces only its contiguous range of ceil(P / world) leaf slots into a zero-initialised full-size
    visibility tensor, and ONE sum all-reduce restores the complete tensor on every rank (each row is
    written by exactly one rank); directions / areas are deterministic and cheap, so every rank
    generates all of them locally (r3dg_sample_incident_dirs) instead of exchanging 16 B per ray.
    The reference is single-GPU; without `shard_group` this is its loop as one launch."""
    import torch.distributed as tdist
    P = xyz.shape[0]
    world, rank, group = 1, 0, None
    if shard_group is not None and shard_group is not False and tdist.is_available() and tdist.is_initialized():
        group = None if shard_group is True else shard_group
        world, rank = tdist.get_world_size(group), tdist.get_rank(group)
    raytracer = RayTracer(xyz, scaling, rotation)
    per_rank = -(-P // world)
    lo, hi = min(rank * per_rank, P), min((rank + 1) * per_rank, P)
    if world == 1:
        out = raytracer.bake_visibility(xyz, inverse_covariance, opacity, normal, sample_num)
        return out["visibility"], out["incident_dirs"], out["incident_areas"]
    out = raytracer.bake_visibility(xyz, inverse_covariance, opacity, normal, sample_num, first_slot=lo, count=hi - lo, write_dirs=False)
    tdist.all_reduce(out["visibility"], op=tdist.ReduceOp.SUM, group=group)
    dirs, areas = sample_incident_rays(normal, False, sample_num)
    return out["visibility"], dirs, areas
