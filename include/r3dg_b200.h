/*
 * r3dg_b200 — C ABI of the B200-native (sm_100a) relightable-Gaussian rasterizer hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types, an
 * explicit CUDA stream, caller-allocated memory, no hidden allocation, errors returned as
 * negative cudaError_t (never thrown).  Every entry point names the reference interface it
 * replaces; the Python host side (relightable3dgaussian_b200/_C_raster.py) rebuilds the exact
 * tuples of the reference's pybind module `r3dg_rasterization._C` on top of it.
 *
 * All pointers are DEVICE pointers unless stated otherwise; optional inputs are NULL when absent
 * (the reference signals absence with an empty tensor whose data_ptr is null,
 * r3dg-rasterization/cuda_rasterizer/forward.cu:206,242).
 */
#ifndef R3DG_B200_H_
#define R3DG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* r3dg_stream_t; /* == cudaStream_t */

#define R3DG_OK 0
#define R3DG_ERR_BAD_ARG (-10001)
#define R3DG_ERR_UNSUPPORTED (-10002)

/* ------------------------------------------------------------------------------------------
 * Work-buffer sizing.  Replaces CudaRasterizer::required<GeometryState/ImageState/BinningState>
 * (r3dg-rasterization/cuda_rasterizer/rasterizer_impl.h:67-73, rasterizer_impl.cu:155-195).
 * The three buffers are opaque blobs that round-trip from forward to backward exactly like the
 * reference's geomBuffer / imgBuffer / binningBuffer (rasterize_points.cu:83-90,136-140).
 * `capacity` is the number of (tile, Gaussian) instances the binning buffer can hold.
 * ------------------------------------------------------------------------------------------ */
size_t r3dg_raster_geom_bytes(int P, int S);
size_t r3dg_raster_img_bytes(int W, int H);
size_t r3dg_raster_binning_bytes(long long capacity);
/* Byte offset of the int32 n_contrib[H,W] plane inside the image buffer: the reference returns
 * n_contrib as a view of imgBuffer (rasterize_points.cu:136-139) and so does the host glue. */
size_t r3dg_raster_img_n_contrib_offset(int W, int H);

typedef struct r3dg_raster_fwd_args {
    int P, S, D, M, W, H;          /* Gaussians, feature channels, SH degree, SH coeffs, image */
    const float* background;       /* [3] */
    const float* means3D;          /* [P,3] */
    const float* shs;              /* [P,M,3] or NULL */
    const float* colors_precomp;   /* [P,3]  or NULL (exactly one of shs / colors_precomp) */
    const float* features;         /* [P,S]  or NULL when S == 0 */
    const float* opacities;        /* [P] */
    const float* scales;           /* [P,3]  or NULL */
    const float* rotations;        /* [P,4]  or NULL */
    const float* cov3D_precomp;    /* [P,6]  or NULL (exactly one of scales+rotations / cov3D) */
    const float* viewmatrix;       /* [16] world->view, stored transposed (scene/cameras.py:62) */
    const float* projmatrix;       /* [16] full projection, stored transposed */
    const float* campos;           /* [3] */
    float scale_modifier, tan_fovx, tan_fovy, cx, cy;
    int prefiltered, computer_pseudo_normal, debug;
    /* outputs (need not be initialised; every element is written) */
    float* out_color;              /* [3,H,W] */
    float* out_opacity;            /* [1,H,W] */
    float* out_depth;              /* [1,H,W] */
    float* out_feature;            /* [S,H,W] */
    float* out_normal;             /* [3,H,W] */
    float* out_surface_xyz;        /* [3,H,W] */
    float* out_weights;            /* [P] */
    int* radii;                    /* [P] */
    int* n_contrib;                /* [H,W] */
    /* opaque work buffers */
    void* geom;   size_t geom_bytes;
    void* img;    size_t img_bytes;
    void* binning; size_t binning_bytes;
    /* Pinned HOST int receiving num_rendered via an async copy on `stream` (may be NULL).
     * num_rendered > capacity means the binning buffer was too small: outputs are invalid and
     * the call must be repeated with a larger buffer (nothing out of bounds is ever written). */
    int* num_rendered_host;
    /* optional cudaEvent_t: recorded on `stream` right after the copy into num_rendered_host has been enqueued, which
     * happens as soon as the instance count exists on the device (after the binning offsets, ~1/4 into the forward) —
     * a host that waits on this event instead of the stream gets the count while the compositor is still running and
     * can already enqueue what follows the forward.  NULL: not used. */
    void* count_ready_event;
} r3dg_raster_fwd_args;

/* Replaces CudaRasterizer::Rasterizer::forward (rasterizer.h:34-65, rasterizer_impl.cu:199-380)
 * == `_C.rasterize_gaussians` (rasterize_points.cu:36-141).  Enqueues the whole forward on
 * `stream` without any host synchronisation. */
int r3dg_raster_forward(const r3dg_raster_fwd_args* args, r3dg_stream_t stream);

typedef struct r3dg_raster_bwd_args {
    int P, S, D, M, W, H;
    const float* background;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* features;
    const float* scales;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float scale_modifier, tan_fovx, tan_fovy;
    int backward_geometry, debug;
    /* cotangents, contiguous */
    const float* dL_dout_color;    /* [3,H,W] */
    const float* dL_dout_opacity;  /* [1,H,W] */
    const float* dL_dout_depth;    /* [1,H,W] */
    const float* dL_dout_feature;  /* [S,H,W] */
    /* gradients out (need not be initialised; every element is written) */
    float* dL_dmeans2D;            /* [P,3]  (depth gradient in .z) */
    float* dL_dcolors;             /* [P,3] */
    float* dL_dopacity;            /* [P] */
    float* dL_dmeans3D;            /* [P,3] */
    float* dL_dfeatures;           /* [P,S] */
    float* dL_dcov3D;              /* [P,6] */
    float* dL_dsh;                 /* [P,M,3]; may be NULL when dL_dsh_factor is given */
    float* dL_dscales;             /* [P,3] */
    float* dL_drotations;          /* [P,4] */
    float* dL_dsh_factor;          /* [P,3] or NULL: the view's rank-1 factor of dL_dsh, see r3dg_sh_grad_from_factors */
    /* work buffers written by the matching forward */
    void* geom;   size_t geom_bytes;
    void* img;    size_t img_bytes;
    void* binning; size_t binning_bytes;
} r3dg_raster_bwd_args;

/* Replaces CudaRasterizer::Rasterizer::backward (rasterizer.h:67-100, rasterizer_impl.cu:384-491)
 * == `_C.rasterize_gaussians_backward` (rasterize_points.cu:143-235). */
int r3dg_raster_backward(const r3dg_raster_bwd_args* args, r3dg_stream_t stream);

/* Multi-GPU exchange step (the reference is single-GPU).  For one view the SH
 * gradient is rank-1 per Gaussian: dL_dsh[g,k,:] = basis_k(normalize(mean_g - campos)) * f[g,:]
 * with f = dL_dRGB gated by the forward's clamp flags (backward.cu:20-139) and zero for culled
 * Gaussians — r3dg_raster_backward writes f to `dL_dsh_factor`.  Instead of all-reducing the dense
 * [P,M,3] tensor (192 B per Gaussian) the ranks all-gather f (12 B per Gaussian and view) and every
 * rank rebuilds the reduced gradient locally:
 *     dL_dsh[g,k,c] = scale * sum_v basis_k(normalize(means3D[g] - campos[v])) * factors[v][g][c]
 * for k < (D+1)^2 (zero above), v = 0..num_views-1 in order (deterministic, identical on every
 * rank).  campos: [num_views,3]; factors: [num_views,P,3]; dL_dsh: [P,M,3] (every element written). */
int r3dg_sh_grad_from_factors(int P, int D, int M, int num_views, const float* means3D,
                              const float* campos, const float* factors, float scale,
                              float* dL_dsh, r3dg_stream_t stream);

/* The same exchange as ONE kernel over NVLink peer memory (no NCCL call): the per-view factors and the dense
 * per-Gaussian gradient section live in symmetric buffers mapped into every rank (the caller allocates and
 * rendezvouses them, e.g. torch.distributed._symmetric_memory).  The launch
 *   (a) rebuilds dL_dsh as r3dg_sh_grad_from_factors does, loading view v's 12 B per Gaussian straight from rank v's
 *       buffer (coalesced P2P loads: the transfer overlaps the outer-product math), scale = 1/world;
 *   (b) replaces the section dense[rank][0..n_dense) on EVERY rank by its mean over the ranks: rank r reduces slice r
 *       with `multimem.ld_reduce` through the NVSwitch multicast mapping `dense_multicast` (in-switch reduction) and
 *       writes it back to all ranks with `multimem.st`; with dense_multicast == NULL it falls back to peer loads/stores.
 * The caller must order it between two cross-rank barriers (all ranks' backward kernels done before; all ranks'
 * exchange kernels done before anyone reads the dense section or overwrites a factor buffer).  n_dense: floats,
 * a multiple of 4, 16-byte aligned sections.  world <= 64. */
typedef struct r3dg_exchange_args {
    int P, D, M, world, rank;
    const float* means3D;          /* [P,3] (replicated) */
    const float* campos;           /* [world,3] camera centres of the ranks' views */
    const float* factors[64];      /* [v]: rank v's [P,3] factor buffer (peer pointer; [rank] is local) */
    float* dL_dsh;                 /* [P,M,3] local output */
    long long n_dense;             /* 0: no dense job */
    float* dense[64];              /* [v]: rank v's dense section (peer pointer) */
    float* dense_multicast;        /* multicast address of the section or NULL */
} r3dg_exchange_args;
int r3dg_exchange_p2p(const r3dg_exchange_args* args, r3dg_stream_t stream);

/* Replaces Rasterizer::markVisible (rasterizer_impl.cu:141-153) == `_C.mark_visible`
 * (rasterize_points.cu:237-256).  present: uint8/bool [P]. */
int r3dg_mark_visible(int P, const float* means3D, const float* viewmatrix,
                      const float* projmatrix, uint8_t* present, r3dg_stream_t stream);

/* Debug / parity introspection of the opaque buffers after a forward (SURVEY.md §8c asks for the
 * same visibility the reference's obtain() layout gives).  Copies one named intermediate into
 * `dst` (device pointer) as a dense array in the REFERENCE's element layout.
 * ids: 0 depths f32[P] | 1 clamped u8[3P] | 3 means2D f32[2P] | 5 conic_opacity f32[4P]
 *      6 rgb f32[3P] | 7 tiles_touched u32[P] | 8 point_offsets u32[P] | 9 point_list u32[R]
 *      10 point_list_keys u64[R] | 13 final_T f32[HW] | 14 n_contrib i32[HW] | 15 ranges u32[2T]
 *      (ours only) 16 bwd_work u32[2T]: entries composited by the busiest warp of each half-tile
 *      CTA | 17 bwd_order u32[2T]: the backward compositor's launch order (valid after a backward)
 * Returns bytes written or a negative error. */
long long r3dg_raster_debug_copy(int id, int P, int S, int W, int H, const void* geom,
                                 const void* img, const void* binning, size_t binning_bytes,
                                 void* dst, long long max_bytes, r3dg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BVH visibility: LBVH over the Gaussians' 3-sigma boxes + opacity ray trace.
 * Replaces the reference's `bvh_tracing._C` (bvh/src/bindings.cpp:8-13, bvh/include/bvh.h).
 * ------------------------------------------------------------------------------------------ */
size_t r3dg_bvh_build_tmp_bytes(int P);
size_t r3dg_bvh_trace_tmp_bytes(int P);

/* Fused replacement of the PyTorch prologue of RayTracer.__init__ (bvh/__init__.py:29-57):
 * fills nodes i32[2P-1,5] (all -1, count column 0 for internal / 1 for leaf nodes) and aabbs
 * f32[2P-1,6] (internal: +1e5/-1e5, leaf i at row P-1+i: box of the 8 corners mu +- 3 s_a a +- ...),
 * with the same op-by-op fp32 rounding as the PyTorch expression. */
int r3dg_bvh_leaf_aabbs(int P, const float* means3D, const float* scales, const float* rotations,
                        int32_t* nodes, float* aabbs, r3dg_stream_t stream);

/* Replaces create_bvh / construct_bvh (bvh/src/bvh.cu:8-27, bvh/src/construct.cu:147-266):
 * nodes / aabbs pre-filled as above are completed IN PLACE (leaf half reordered by Morton code,
 * internal nodes, parent links, subtree leaf counts, refitted boxes); morton u64[P] out.
 * tmp: r3dg_bvh_build_tmp_bytes(P) bytes of scratch. */
int r3dg_bvh_build(int P, int32_t* nodes, float* aabbs, uint64_t* morton, void* tmp, size_t tmp_bytes,
                   r3dg_stream_t stream);

/* Replaces trace_bvh_opacity (bvh/src/bvh.cu:88-116, bvh/src/trace.cu:196-287).  Ray r has
 * direction rays_d[r] and origin rays_o[r / rays_per_origin] + origin_offset * rays_d[r]
 * (rays_per_origin = 1, origin_offset = 0 reproduces the reference call; the host mirror of
 * RayTracer.trace_visibility passes the un-expanded origins and 0.05, bvh/__init__.py:63).
 * covs3D holds the INVERSE covariances [P,6].  Outputs: num_contributes i32[num_rays],
 * rendered_opacity f32[num_rays] in {0} U [0.9, 1]; every element is written.
 * tmp: r3dg_bvh_trace_tmp_bytes(P) bytes of scratch (re-packed tree). */
int r3dg_bvh_trace_opacity(int P, long long num_rays, const int32_t* nodes, const float* aabbs,
                           const float* rays_o, int rays_per_origin, float origin_offset,
                           const float* rays_d, const float* means3D, const float* covs3D,
                           const float* opacities, const float* normals, int32_t* num_contributes,
                           float* rendered_opacity, void* tmp, size_t tmp_bytes, r3dg_stream_t stream);

/* Replaces `sample_incident_rays(normals, False, N)` of the visibility bake
 * (scene/gaussian_model.py:20-28 -> utils/graphics_utils.py:9-37 fibonacci_sphere_sampling with
 * random_rotate=False -> utils/sh_utils.py:36-68 rotation_between_z; ~20 PyTorch kernels and
 * [P,N,3]-sized intermediates in the reference): dirs f32[P,N,3] unit vectors, areas f32[P,N,1]
 * = 2*pi (may be NULL).  phase f32[P] (may be NULL) = the caller's torch.rand(P) of
 * random_rotate=True (is_training): theta_i += phase[g]*2*pi.  Same op-by-op fp32 rounding as
 * the PyTorch expression. */
int r3dg_sample_incident_dirs(int P, int N, const float* normals, const float* phase, float* dirs,
                              float* areas, r3dg_stream_t stream);

/* The visibility bake of `GaussianModel.update_visibility` (scene/gaussian_model.py:312-342) for
 * the Gaussians of leaf slots [first_slot, first_slot + count) of a built tree (slot = position in
 * Morton order; 0, P bakes everything) in ONE launch: per Gaussian g and sample i the ray
 * direction is generated in the kernel exactly as r3dg_sample_incident_dirs does (never read from
 * HBM), the origin is mean_g + origin_offset * dir (0.05, bvh/__init__.py:63), and the ray is
 * traced like r3dg_bvh_trace_opacity.  Results are written to row g of the FULL-size outputs
 * visibility f32[P,N], num_contributes i32[P,N] (may be NULL), dirs f32[P,N,3] and areas
 * f32[P,N,1] (may be NULL: written once, here); rows of other slots are not touched.
 * tmp: r3dg_bvh_trace_tmp_bytes(P). */
int r3dg_bvh_bake_visibility(int P, int first_slot, int count, int N, const int32_t* nodes,
                             const float* aabbs, const float* means3D, const float* covs3D,
                             const float* opacities, const float* normals, float origin_offset,
                             int32_t* num_contributes, float* visibility, float* dirs, float* areas,
                             void* tmp, size_t tmp_bytes, r3dg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optional fused epilogue of the rasterizer call sites:
 *     feature / opacity.clamp_min(1e-5) * (num_contrib > 0)
 * (gaussian_renderer/neilf.py:135-137, render.py:106-108; 3 PyTorch kernels forward, ~6 backward).
 * feature / out / dL_* are [S,HW] planar, opacity [HW], n_contrib i32[HW].
 * ------------------------------------------------------------------------------------------ */
int r3dg_unpremultiply_forward(int S, long long HW, const float* feature, const float* opacity,
                               const int32_t* n_contrib, float* out, r3dg_stream_t stream);
int r3dg_unpremultiply_backward(int S, long long HW, const float* feature, const float* opacity,
                                const int32_t* n_contrib, const float* dL_dout, float* dL_dfeature,
                                float* dL_dopacity, r3dg_stream_t stream);

/* Optional fused feature pack in front of the rasterizer (gaussian_renderer/neilf.py:110-126,
 * render.py:88-93): features[P,S] = cat([depths, depths^2, src_0, src_1, ...], -1) with
 * depths = (cat([means3D, 1]) @ viewmatrix)[:, 2] — means3D / viewmatrix NULL: no depth channels.  S must equal
 * (2 if depth) + sum(width).  Backward: writes every source's gradient slice contiguous (ptr NULL: skipped) and, when
 * dL_dmeans3D is given, the gradient through the two depth channels. */
#define R3DG_PACK_MAX 16
typedef struct r3dg_pack_src {
    const float* ptr;              /* [P,width] contiguous (backward: destination of the gradient slice) */
    int width;
} r3dg_pack_src;
int r3dg_pack_features_forward(int P, int S, const float* means3D, const float* viewmatrix, int num,
                               const r3dg_pack_src* srcs, float* out, r3dg_stream_t stream);
int r3dg_pack_features_backward(int P, int S, const float* means3D, const float* viewmatrix,
                                const float* dL_dout, int num, const r3dg_pack_src* dsts,
                                float* dL_dmeans3D, r3dg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Densification surgery: stable compaction of the rows of MANY per-Gaussian tensors
 * by one keep-mask.  Replaces the ~40 `tensor[mask]` boolean-index launches of
 * `_prune_optimizer` / `prune_points` / `densify_and_prune` (scene/gaussian_model.py:682-729,890-929):
 *   r3dg_compact_scan  : exclusive prefix sum of keep[P] into tmp, number of kept rows -> *count_host
 *                        (pinned host int or NULL; copied asynchronously on `stream`);
 *   r3dg_compact_rows  : dst_k[offset(i)] = src_k[i] for every kept row i of every tensor k, one launch
 *                        per <= 32 tensors.  row_bytes: a multiple of 4; dst must hold count rows.
 * ------------------------------------------------------------------------------------------ */
#define R3DG_COMPACT_MAX 32
typedef struct r3dg_compact_tensor {
    const void* src;               /* [P, row_bytes] */
    void* dst;                     /* [count, row_bytes] */
    long long row_bytes;
} r3dg_compact_tensor;
size_t r3dg_compact_tmp_bytes(int P);
int r3dg_compact_scan(int P, const uint8_t* keep, void* tmp, size_t tmp_bytes, int* count_host,
                      r3dg_stream_t stream);
int r3dg_compact_rows(int P, int num_tensors, const r3dg_compact_tensor* tensors, const uint8_t* keep,
                      const void* tmp, r3dg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BRDF shading: fused `rendering_equation` (gaussian_renderer/neilf.py:339-371) + `GGX_specular`
 * (:374-406) + lat-long environment lookup (scene/direct_light_map.py:70-83) + SH incident light.
 * The reference has no extension boundary here (pure PyTorch; its render_equation.cu is dead,
 * unbuilt code with a different BRDF — SURVEY.md fact 2); the host mirror
 * relightable3dgaussian_b200/shading.py gives this entry the reference function's signature.
 * ------------------------------------------------------------------------------------------ */
typedef struct r3dg_shade_args {
    int P, N;                          /* Gaussians, incident samples per Gaussian */
    int sh_coeffs;                     /* incidents.shape[1]; 16 (degree 3) is supported */
    int env_h, env_w;                  /* environment texture size */
    const float* base_color;           /* [P,3] */
    const float* roughness;            /* [P]   */
    const float* normals;              /* [P,3] (no gradient: detached at neilf.py:94) */
    const float* viewdirs;             /* [P,3] */
    const float* incidents;            /* [P,16,3] SH coefficients of the local incident light */
    const float* env;                  /* [env_h,env_w,3] ACTIVATED environment map (softplus / HDR) */
    const float* env_transform;        /* [3,3] row-major or NULL (EnvLight.transform, envmap.py:39-42) */
    const float* visibility;           /* [P,N]   baked */
    const float* incident_dirs;        /* [P,N,3] baked */
    const float* incident_areas;       /* [P,N]   baked */
    /* forward outputs; the three [P,3] means and mean_visibility are optional (all or none),
     * the three per-sample [P,N,3] arrays are optional (all or none, eval only) */
    float* pbr;                        /* [P,3] */
    float* diffuse_light;              /* [P,3] */
    float* specular;                   /* [P,3] */
    float* mean_incident_lights;       /* [P,3] or NULL  (= incident_lights.mean(-2)) */
    float* mean_local_lights;          /* [P,3] or NULL */
    float* mean_global_lights;         /* [P,3] or NULL */
    float* mean_visibility;            /* [P]   or NULL */
    float* incident_lights;            /* [P,N,3] or NULL */
    float* local_incident_lights;      /* [P,N,3] or NULL */
    float* global_incident_lights;     /* [P,N,3] or NULL */
    /* backward: cotangents (dL_dpbr required, the others may be NULL) and gradients out */
    const float* dL_dpbr;              /* [P,3] */
    const float* dL_ddiffuse_light;    /* [P,3] or NULL */
    const float* dL_dspecular;         /* [P,3] or NULL */
    float* dL_dbase_color;             /* [P,3] */
    float* dL_droughness;              /* [P] */
    float* dL_dviewdirs;               /* [P,3] */
    float* dL_dincidents;              /* [P,16,3] */
    float* dL_denv;                    /* [env_h,env_w,3] (zeroed, then accumulated) */
} r3dg_shade_args;

int r3dg_render_equation_forward(const r3dg_shade_args* args, r3dg_stream_t stream);
int r3dg_render_equation_backward(const r3dg_shade_args* args, r3dg_stream_t stream);

/* Replaces `simple_knn._C.distCUDA2` (submodules/simple-knn/spatial.cu:15-26, simple_knn.cu:185-221):
 * mean_dist2[i] = mean of the 3 smallest squared distances from point i to the other points.
 * Init-time only in the reference (scene/gaussian_model.py:427), but imported at module import. */
size_t r3dg_knn_tmp_bytes(int P);
int r3dg_knn_dist2(int P, const float* points /* [P,3] */, float* mean_dist2 /* [P] */, void* tmp,
                   size_t tmp_bytes, r3dg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser step.  Replaces, for all parameter groups at once, the
 * per-group `torch.optim.Adam(l, lr=0.0, eps=1e-15).step()` of scene/gaussian_model.py:489,495-497
 * (torch 1.12.1 torch/optim/adam.py:_single_tensor_adam: no weight decay, no amsgrad, no
 * maximize — the only configuration the reference uses).  `tensors` is a HOST array; every
 * entry updates param / exp_avg / exp_avg_sq in place from grad.  `step` is the 1-based count of
 * this update for that tensor (torch increments state['step'] before using it).
 * ------------------------------------------------------------------------------------------ */
typedef struct r3dg_adam_tensor {
    float* param;              /* [n] */
    const float* grad;         /* [n] */
    float* exp_avg;            /* [n] */
    float* exp_avg_sq;         /* [n] */
    long long n;
    long long step;
    double lr, beta1, beta2, eps;
} r3dg_adam_tensor;
int r3dg_adam_step(int num_tensors, const r3dg_adam_tensor* tensors, r3dg_stream_t stream);

/* Measurement hooks used by bench.py (never needed by the reference's callers).
 * r3dg_launch_count: number of this library's kernels launched so far in the process.
 * r3dg_prof_begin/end: while active, every forward/backward records CUDA events on the launching
 * stream between its stages (0 project, 1 depth sort, 2 bin count, 3 bin offsets, 4 bin scatter,
 * 5 composite fwd, 6 surface/normal, 7 composite bwd, 8 projection bwd); r3dg_prof_end sums the
 * per-stage milliseconds over the recorded calls (caller synchronises first). */
unsigned long long r3dg_launch_count(void);
int r3dg_prof_begin(int max_calls);
int r3dg_prof_end(float* stage_ms /* [9] */, int* fwd_calls, int* bwd_calls);

/* Tuning knobs (development / measurement; results are identical for every setting, only the
 * kernel variant changes).  Sets knob `key` to `value`, stores the old value in *previous (may be
 * NULL); R3DG_ERR_UNSUPPORTED for an unknown key, R3DG_ERR_BAD_ARG for an out-of-range value.
 *   "shade_group"     lanes per Gaussian in the shading kernels: 8 (default), 16, 32
 *   "shade_fwd_variant" / "shade_bwd_variant"  0 per-Gaussian SH state in registers, 1 incident
 *                     coefficients in shared memory (forward default), 2 (backward, default) +
 *                     gradient accumulators in shared memory
 *   "shade_env_mode"  env-map gradient accumulation in r3dg_render_equation_backward:
 *                     2 warp-private tagged copies, 1 shared-memory atomics (default),
 *                     0 global atomics; larger textures fall back to the lower modes
 *   "composite_bulk"  1 = per-lane cp.async.bulk record staging in the forward compositor
 *                     (measured slower, profiles/r02_composite_bulk_staging.md), 0 (default)
 *   "composite_fwd_ctas" / "composite_bwd_ctas"  resident CTAs per SM of the two compositors,
 *                     enforced with dynamic shared memory; 0 (default) = whatever fits
 *                     (profiles/r02_warp_timing.md: every lower setting is slower)
 * Initial values can also be given by the environment (R3DG_SHADE_GROUP, R3DG_SHADE_ENV_MODE,
 * R3DG_SHADE_FWD_VARIANT, R3DG_SHADE_BWD_VARIANT, R3DG_COMPOSITE_BULK, R3DG_FWD_CTAS,
 * R3DG_BWD_CTAS). */
int r3dg_tune(const char* key, int value, int* previous);

/* Library identification: returns a static string such as "r3dg_b200 0.1 sm_100a". */
const char* r3dg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* R3DG_B200_H_ */
