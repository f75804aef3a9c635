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
                       const uint64_t* keys_b, void* ranges, int num_tiles, int num_sms,
                       cudaStream_t stream);
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

}  // namespace r3dg
