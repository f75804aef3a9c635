// Stage 2 forward: per-tile front-to-back alpha compositing (reference K7, forward.cu:263-395)
// and the surface-xyz / pseudo-normal pass (K8+K9, forward.cu:398-491, fused into one kernel).
//
// B200 design notes
//  * one CTA per 16x16 tile, each warp owns a compact 8x4 pixel block so that early-outs and the
//    per-Gaussian weight reduction are warp-uniform;
//  * a batch of 256 instances is staged into shared memory as float4 SoA from ONE packed record
//    per Gaussian (16B vector loads, conflict-free stores, broadcast reads in the pixel loop);
//    the reference re-reads colour / feature / depth from global memory per (pixel, Gaussian);
//  * accumulators live in registers (template on the number of float4 channel groups) — the
//    reference's runtime-indexed F[33] spills to local memory;
//  * out_weights: one atomic per (warp, Gaussian) after a shuffle reduction instead of one per
//    (pixel, Gaussian);
//  * per-pixel arithmetic keeps the association of the reference binary, so n_contrib and the
//    images are bit-identical to it for identical lists.
#include "common.cuh"
#include "kernels.h"

namespace r3dg {

struct CompositeFwdParams {
    int W, H, gx, S, recf;
    const uint2* ranges;
    const uint32_t* point_list;
    const float* rec;
    const float* bg;
    float* final_T;
    int* n_contrib;
    float *out_color, *out_opacity, *out_depth, *out_feature, *out_weights;
};

template <int NG>
__global__ void __launch_bounds__(256) composite_fwd_kernel(const CompositeFwdParams p) {
    __shared__ float4 sA[256], sB[256];
    __shared__ float4 sC[NG][256];
    __shared__ int sId[256];
    __shared__ uint32_t sBits[8][8];          // [pixel block (= consumer warp)][32-entry group of the batch]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tx = tile % p.gx, ty = tile / p.gx;
    const float tile_x0 = (float)(tx * R3DG_TILE), tile_y0 = (float)(ty * R3DG_TILE);
    const int px = tx * R3DG_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = ty * R3DG_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = p.ranges[tile];
    const int toDo = (int)(range.y - range.x);
    const float4* __restrict__ rec4 = reinterpret_cast<const float4*>(p.rec);
    const int rec4n = p.recf >> 2;

    float T = 1.0f, Dp = 0.0f, Op = 0.0f;
    float C[4 * NG];
#pragma unroll
    for (int i = 0; i < 4 * NG; ++i) C[i] = 0.0f;
    uint32_t last_contributor = 0;
    bool done = !inside;

    for (int base = 0; base < toDo; base += 256) {
        if (__syncthreads_and(done)) break;
        const int n = min(256, toDo - base);
        unsigned tm = 0u;
        if (tid < n) {
            const uint32_t id = p.point_list[range.x + base + tid];
            const float4* r = rec4 + (size_t)id * rec4n;
            const float4 A = r[0], B = r[1];
            sId[tid] = (int)id;
            sA[tid] = A;
            sB[tid] = B;
#pragma unroll
            for (int g = 0; g < NG; ++g) sC[g][tid] = r[2 + g];
            tm = touch_mask(A, B, tile_x0, tile_y0);
        }
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t word = __ballot_sync(0xffffffffu, (tm >> w) & 1u);
            if (lane == 0) sBits[w][warp] = word;
        }
        __syncthreads();
        if (__all_sync(0xffffffffu, done)) continue;
        bool warp_done = false;
        for (int k = 0; k < 8 && !warp_done; ++k) {
          uint32_t word = sBits[warp][k];
          while (word) {
            const int j = k * 32 + __ffs(word) - 1;
            word &= word - 1;
            const float4 a = sA[j];
            const float4 b = sB[j];
            const float dx = sub_(a.x, pxf), dy = sub_(a.y, pyf);
            // power = -0.5f*(ca*dx*dx + cc*dy*dy) - cb*dx*dy  (forward.cu:344) as compiled
            const float q = fma_(dx, mul_(dx, a.z), mul_(dy, mul_(dy, b.x)));
            const float power = fma_(q, -0.5f, -mul_(dy, mul_(dx, a.w)));
            const float alpha = fminf(0.99f, mul_(b.y, expf(power)));
            const float test_T = mul_(T, sub_(1.0f, alpha));
            bool valid = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            if (valid && test_T < 0.0001f) { done = true; valid = false; }
            float w = 0.0f;
            if (valid) {
                w = mul_(T, alpha);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const float4 c = sC[g][j];
                    C[4 * g + 0] = fma_(w, c.x, C[4 * g + 0]);
                    C[4 * g + 1] = fma_(w, c.y, C[4 * g + 1]);
                    C[4 * g + 2] = fma_(w, c.z, C[4 * g + 2]);
                    C[4 * g + 3] = fma_(w, c.w, C[4 * g + 3]);
                }
                Dp = fma_(w, b.z, Dp);
                Op = add_(Op, w);
                T = test_T;
                last_contributor = (uint32_t)(base + j + 1);      // 1-based position in the tile list
            }
            if (__any_sync(0xffffffffu, valid)) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
                if (lane == 0) atomicAdd(&p.out_weights[sId[j]], w);
            }
            if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
          }
        }
    }
    if (inside) {
        const size_t HW = (size_t)p.H * p.W, pix = (size_t)p.W * py + px;
        p.final_T[pix] = T;
        p.n_contrib[pix] = (int)last_contributor;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) p.out_color[ch * HW + pix] = fma_(p.bg[ch], T, C[ch]);
#pragma unroll
        for (int ch = 3; ch < 4 * NG; ++ch)
            if (ch - 3 < p.S) p.out_feature[(size_t)(ch - 3) * HW + pix] = C[ch];
        p.out_depth[pix] = Dp;
        p.out_opacity[pix] = Op;
    }
}

// renderSurfaceXYZCUDA + renderPseudoNormalCUDA fused: neighbours' surface points are recomputed
// from (opacity, depth) with the identical formula, so no second pass over surface_xyz is needed.
__global__ void __launch_bounds__(256) surface_normal_kernel(int W, int H, const float* __restrict__ viewmatrix,
                                                             float focal_x, float focal_y, float cx, float cy,
                                                             const float* __restrict__ opacity,
                                                             const float* __restrict__ depth,
                                                             float* __restrict__ out_normal,
                                                             float* __restrict__ out_xyz) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W;
    auto point = [&](int xx, int yy, float* o) {
        const size_t q = (size_t)W * yy + xx;
        const float d = div_(depth[q], fmaxf(opacity[q], 0.0000001f));
        o[0] = mul_(div_(sub_((float)xx, cx), focal_x), d);
        o[1] = mul_(div_(sub_((float)yy, cy), focal_y), d);
        o[2] = d;
    };
    const size_t pix = (size_t)W * y + x;
    float c[3];
    point(x, y, c);
    out_xyz[pix] = c[0]; out_xyz[HW + pix] = c[1]; out_xyz[2 * HW + pix] = c[2];
    const int xm = x == 0 ? 0 : x - 1, xp = x == W - 1 ? W - 1 : x + 1;
    const int ym = y == 0 ? 0 : y - 1, yp = y == H - 1 ? H - 1 : y + 1;
    float p00[3], p01[3], p02[3], p10[3], p12[3], p20[3], p21[3], p22[3];
    point(xm, ym, p00); point(x, ym, p01); point(xp, ym, p02);
    point(xm, y, p10);                     point(xp, y, p12);
    point(xm, yp, p20); point(x, yp, p21); point(xp, yp, p22);
    float ga[3], gb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float h = mul_(p00[i], -0.125f);
        ga[i] = fma_(p22[i], 0.125f, fma_(p20[i], -0.125f, fma_(p12[i], 0.25f, fma_(p10[i], -0.25f, fma_(p02[i], 0.125f, h)))));
        gb[i] = fma_(p22[i], 0.125f, fma_(p21[i], 0.25f, fma_(p20[i], 0.125f, fma_(p02[i], -0.125f, fma_(p01[i], -0.25f, h)))));
    }
    const float n0 = fma_(ga[1], gb[2], -mul_(ga[2], gb[1]));
    const float n1 = fma_(ga[2], gb[0], -mul_(ga[0], gb[2]));
    const float n2 = fma_(ga[0], gb[1], -mul_(ga[1], gb[0]));
    const float norm = sqrt_(fma_(n2, n2, fma_(n0, n0, mul_(n1, n1))));
    float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f;
    if (!(norm <= 0.0f)) {
        const float N0 = div_(-n0, norm), N1 = div_(-n1, norm), N2 = div_(-n2, norm);
        const float* V = viewmatrix;
        o0 = dot3_(V[0], N0, V[1], N1, V[2], N2);
        o1 = dot3_(V[4], N0, V[5], N1, V[6], N2);
        o2 = dot3_(V[8], N0, V[9], N1, V[10], N2);
    }
    out_normal[pix] = o0; out_normal[HW + pix] = o1; out_normal[2 * HW + pix] = o2;
}

template <int NG>
static void launch_fwd_ng(const CompositeFwdParams& p, int tiles, cudaStream_t stream) {
    composite_fwd_kernel<NG><<<tiles, 256, 0, stream>>>(p);
}

int launch_composite_forward(const r3dg_raster_fwd_args& a, const GeomLayout& gl, const ImgLayout& il,
                             const uint32_t* point_list, cudaStream_t stream, stage_mark_fn mark) {
    char* geom = (char*)a.geom;
    char* img = (char*)a.img;
    CompositeFwdParams p;
    p.W = a.W; p.H = a.H; p.gx = (a.W + R3DG_TILE - 1) / R3DG_TILE; p.S = a.S; p.recf = gl.recf;
    const int gy = (a.H + R3DG_TILE - 1) / R3DG_TILE;
    p.ranges = (const uint2*)(img + il.ranges);
    p.point_list = point_list;
    p.rec = (const float*)(geom + gl.rec);
    p.bg = a.background;
    p.final_T = (float*)(img + il.final_T);
    p.n_contrib = (int*)(img + il.n_contrib);
    p.out_color = a.out_color; p.out_opacity = a.out_opacity; p.out_depth = a.out_depth;
    p.out_feature = a.out_feature; p.out_weights = a.out_weights;
    const int tiles = p.gx * gy;
    switch (num_groups(a.S)) {
        case 1: launch_fwd_ng<1>(p, tiles, stream); break;
        case 2: launch_fwd_ng<2>(p, tiles, stream); break;
        case 3: launch_fwd_ng<3>(p, tiles, stream); break;
        case 4: launch_fwd_ng<4>(p, tiles, stream); break;
        case 5: launch_fwd_ng<5>(p, tiles, stream); break;
        case 6: launch_fwd_ng<6>(p, tiles, stream); break;
        case 7: launch_fwd_ng<7>(p, tiles, stream); break;
        case 8: launch_fwd_ng<8>(p, tiles, stream); break;
        case 9: launch_fwd_ng<9>(p, tiles, stream); break;
        default: return R3DG_ERR_UNSUPPORTED;
    }
    const size_t HW = (size_t)a.H * a.W;
    mark(6, stream);
    if (a.computer_pseudo_normal) {
        const float focal_y = a.H / (2.0f * a.tan_fovy), focal_x = a.W / (2.0f * a.tan_fovx);
        dim3 grid((a.W + 31) / 32, (a.H + 7) / 8);
        surface_normal_kernel<<<grid, 256, 0, stream>>>(a.W, a.H, a.viewmatrix, focal_x, focal_y, a.cx, a.cy,
                                                        a.out_opacity, a.out_depth, a.out_normal, a.out_surface_xyz);
    } else {
        R3DG_CUDA_TRY(cudaMemsetAsync(a.out_normal, 0, 3 * HW * 4, stream));
        R3DG_CUDA_TRY(cudaMemsetAsync(a.out_surface_xyz, 0, 3 * HW * 4, stream));
    }
    if (a.n_contrib)   // the reference returns n_contrib as a view of imgBuffer (rasterize_points.cu:136-139)
        R3DG_CUDA_TRY(cudaMemcpyAsync(a.n_contrib, img + il.n_contrib, HW * 4, cudaMemcpyDeviceToDevice, stream));
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace r3dg
