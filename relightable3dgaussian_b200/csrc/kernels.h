// Internal launcher declarations shared by the translation units of libr3dg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include "../../include/r3dg_b200.h"
#include "common.cuh"

namespace r3dg {

typedef void (*stage_mark_fn)(int, cudaStream_t);
int launch_projection(const r3dg_raster_fwd_args& a, const GeomLayout& gl, const BinLayout& bl,
                      cudaStream_t stream, stage_mark_fn mark);
int launch_tile_ranges(const void* geom_header, long long capacity, const uint64_t* keys_a,
                       const uint64_t* keys_b, void* ranges, uint32_t* tile_order, int num_tiles,
                       int num_sms, cudaStream_t stream);
int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                        cudaStream_t stream);
int launch_sort(void* geom_header, char* bin, const BinLayout& bl, int passes, int num_sms,
                cudaStream_t stream);
int launch_composite_forward(const r3dg_raster_fwd_args& a, const GeomLayout& gl,
                             const ImgLayout& il, const uint32_t* vals_a, const uint32_t* vals_b,
                             cudaStream_t stream, stage_mark_fn mark);
int launch_composite_backward(const r3dg_raster_bwd_args& a, const GeomLayout& gl,
                              const ImgLayout& il, const uint32_t* vals_a, const uint32_t* vals_b,
                              cudaStream_t stream);
int launch_projection_backward(const r3dg_raster_bwd_args& a, const GeomLayout& gl,
                               cudaStream_t stream);

size_t bvh_build_tmp_bytes(int P);
size_t bvh_packets_bytes(int P);
int launch_bvh_leaf_aabbs(int P, const float* means3D, const float* scales, const float* rotations,
                          int32_t* nodes, float* aabbs, cudaStream_t stream);
int launch_bvh_build(int P, int32_t* nodes, float* aabbs, uint64_t* morton, void* tmp, size_t tmp_bytes,
                     int num_sms, cudaStream_t stream);
int launch_bvh_trace(int P, long long num_rays, const int32_t* nodes, const float* aabbs,
                     const float* rays_o, int o_group, float o_offset, const float* rays_d,
                     const float* means3D, const float* covs3D, const float* opacities,
                     const float* normals, int32_t* num_contributes, float* rendered_opacity,
                     void* packets, size_t packets_bytes, int num_sms, cudaStream_t stream);

size_t knn_tmp_bytes(int P);
int launch_knn(int P, const float* points, float* out, void* tmp, size_t tmp_bytes, int num_sms, cudaStream_t stream);

}  // namespace r3dg
