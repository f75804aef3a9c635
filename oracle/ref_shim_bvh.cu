// TEST INFRASTRUCTURE — not product code.
// Thin C-ABI shim around the reference BVH (construct_bvh: /root/reference/bvh/src/construct.cu:147-266,
// trace_bvh_opacity_cuda: /root/reference/bvh/src/trace.cu:196-287), replacing the torch glue of
// bvh/src/bvh.cu:8-27,88-116 with raw device pointers.  Built by oracle/build_ref.sh into
// oracle/_ref/libref_bvh.so (construct.cu needs the one-line patched COPY described there).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "construct.cuh"
#include "trace.cuh"

extern "C" {

// nodes i32[2P-1,5] and aabbs f32[2P-1,6] are pre-filled by the caller exactly like
// bvh/__init__.py:32-57 and mutated in place; morton u64[P] out.
int ref_bvh_create(int P, const float* means3D, const float* scales, const float* rotations,
                   int32_t* nodes, float* aabbs, uint64_t* morton) {
    try {
        construct_bvh(P, means3D, scales, rotations, nodes, aabbs, morton);
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim_bvh] construct threw: %s\n", e.what());
        return -1;
    }
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

// num_contributes must be zero-filled and rendered_opacity one-filled by the caller (bvh.cu:101-102).
int ref_bvh_trace_opacity(int num_rays, int32_t* nodes, float* aabbs, float* rays_o, float* rays_d,
                          float* means3D, float* covs3D, float* opacities, float* normals,
                          int32_t* num_contributes, float* rendered_opacity) {
    try {
        trace_bvh_opacity_cuda(num_rays, nodes, aabbs, (float3*)rays_o, (float3*)rays_d, (float3*)means3D,
                               covs3D, opacities, (float3*)normals, num_contributes, rendered_opacity);
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim_bvh] trace threw: %s\n", e.what());
        return -1;
    }
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
