// TEST INFRASTRUCTURE — not product code.
// Thin C-ABI shim around the *unmodified* reference rasterizer
// (CudaRasterizer::Rasterizer::{forward,backward,markVisible},
//  /root/reference/r3dg-rasterization/cuda_rasterizer/rasterizer.h:20-100).
// The reference sources are compiled where they lie under /root/reference by
// oracle/build_ref.sh; only this shim is ours.  It replaces the torch glue of
// rasterize_points.cu:36-256 with raw device pointers so the reference kernels can
// be driven through ctypes and used (a) as the parity oracle on the GPU box and
// (b) as the timed `--impl reference` arm of bench.py.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/config.h"
#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"

namespace {
struct Buf {
    char* ptr = nullptr;
    size_t cap = 0;
    size_t size = 0;
    char* resize(size_t n) {
        if (n > cap) {
            if (ptr) cudaFree(ptr);
            size_t ncap = n + n / 4 + 1024;
            if (cudaMalloc(&ptr, ncap) != cudaSuccess) { ptr = nullptr; cap = 0; return nullptr; }
            cap = ncap;
        }
        size = n;
        return ptr;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = size = 0; }
};
struct Ctx {
    Buf geom, binning, img;
    int P = 0, R = 0, HW = 0;
};
}  // namespace

extern "C" {

void* ref_ctx_create() { return new Ctx(); }
void ref_ctx_destroy(void* c) {
    Ctx* ctx = (Ctx*)c;
    ctx->geom.release(); ctx->binning.release(); ctx->img.release();
    delete ctx;
}

int ref_raster_forward(void* c, int P, int S, int D, int M, const float* background, int W, int H,
                       const float* means3D, const float* shs, const float* colors_precomp,
                       const float* features, const float* opacities, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* campos,
                       float tan_fovx, float tan_fovy, float cx, float cy, int prefiltered,
                       int computer_pseudo_normal, float* out_color, float* out_opacity,
                       float* out_depth, float* out_feature, float* out_normal,
                       float* out_surface_xyz, float* out_weights, int* radii, int debug) {
    Ctx* ctx = (Ctx*)c;
    std::function<char*(size_t)> g = [ctx](size_t n) { return ctx->geom.resize(n); };
    std::function<char*(size_t)> b = [ctx](size_t n) { return ctx->binning.resize(n); };
    std::function<char*(size_t)> i = [ctx](size_t n) { return ctx->img.resize(n); };
    int R = -1;
    try {
        R = CudaRasterizer::Rasterizer::forward(
            g, b, i, P, S, D, M, background, W, H, means3D, shs, colors_precomp, features,
            opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
            campos, tan_fovx, tan_fovy, cx, cy, prefiltered != 0, computer_pseudo_normal != 0,
            out_color, out_opacity, out_depth, out_feature, out_normal, out_surface_xyz,
            out_weights, radii, debug != 0);
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] forward threw: %s\n", e.what());
        return -1;
    }
    ctx->P = P; ctx->R = R; ctx->HW = W * H;
    return R;
}

int ref_raster_backward(void* c, int P, int S, int D, int M, int R, const float* background, int W,
                        int H, const float* means3D, const float* shs, const float* features,
                        const float* colors_precomp, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos,
                        float tan_fovx, float tan_fovy, const int* radii, const float* dL_dpix,
                        const float* dL_dpix_o, const float* dL_dpix_d, const float* dL_dpix_f,
                        float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                        float* dL_dfeature, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                        float* dL_dscale, float* dL_drot, int backward_geometry, int debug) {
    Ctx* ctx = (Ctx*)c;
    try {
        CudaRasterizer::Rasterizer::backward(
            P, S, D, M, R, background, W, H, means3D, shs, features, colors_precomp, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx,
            tan_fovy, radii, ctx->geom.ptr, ctx->binning.ptr, ctx->img.ptr, dL_dpix, dL_dpix_o,
            dL_dpix_d, dL_dpix_f, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dfeature,
            dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, backward_geometry != 0, debug != 0);
    } catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] backward threw: %s\n", e.what());
        return -1;
    }
    return 0;
}

void ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
}

// Copy one named intermediate of the last forward out of the reference's opaque byte
// buffers (layouts: rasterizer_impl.cu:155-195) into caller memory (device pointer).
// ids: 0 depths f32[P] | 1 clamped u8[3P] | 2 radii i32[P] | 3 means2D f32[2P] | 4 cov3D f32[6P]
//      5 conic_opacity f32[4P] | 6 rgb f32[3P] | 7 tiles_touched u32[P] | 8 point_offsets u32[P]
//      9 point_list u32[R] | 10 point_list_keys u64[R] | 11 point_list_unsorted u32[R]
//      12 point_list_keys_unsorted u64[R] | 13 accum_alpha f32[HW] | 14 n_contrib u32[HW]
//      15 ranges uint2[HW] (first T entries meaningful)
long long ref_copy_out(void* c, int id, void* dst, long long max_bytes) {
    Ctx* ctx = (Ctx*)c;
    const void* src = nullptr; size_t n = 0;
    size_t P = ctx->P, R = ctx->R, HW = ctx->HW;
    if (id <= 8) {
        char* p = ctx->geom.ptr;
        auto s = CudaRasterizer::GeometryState::fromChunk(p, P);
        switch (id) {
            case 0: src = s.depths; n = 4 * P; break;
            case 1: src = s.clamped; n = 3 * P; break;
            case 2: src = s.internal_radii; n = 4 * P; break;
            case 3: src = s.means2D; n = 8 * P; break;
            case 4: src = s.cov3D; n = 24 * P; break;
            case 5: src = s.conic_opacity; n = 16 * P; break;
            case 6: src = s.rgb; n = 12 * P; break;
            case 7: src = s.tiles_touched; n = 4 * P; break;
            case 8: src = s.point_offsets; n = 4 * P; break;
        }
    } else if (id <= 12) {
        char* p = ctx->binning.ptr;
        auto s = CudaRasterizer::BinningState::fromChunk(p, R);
        switch (id) {
            case 9: src = s.point_list; n = 4 * R; break;
            case 10: src = s.point_list_keys; n = 8 * R; break;
            case 11: src = s.point_list_unsorted; n = 4 * R; break;
            case 12: src = s.point_list_keys_unsorted; n = 8 * R; break;
        }
    } else {
        char* p = ctx->img.ptr;
        auto s = CudaRasterizer::ImageState::fromChunk(p, HW);
        switch (id) {
            case 13: src = s.accum_alpha; n = 4 * HW; break;
            case 14: src = s.n_contrib; n = 4 * HW; break;
            case 15: src = s.ranges; n = 8 * HW; break;
        }
    }
    if (!src) return -1;
    if ((long long)n > max_bytes) n = (size_t)max_bytes;
    if (cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice) != cudaSuccess) return -2;
    return (long long)n;
}

}  // extern "C"
