// Stage 2 backward: per-tile back-to-front re-traversal (reference K10, backward.cu:401-614).
//
// B200 design notes
//  * the reference issues 10+S global atomicAdds per contributing (pixel, Gaussian) pair.  Here a
//    warp (8x4 pixels) first reduces its 32 per-pixel gradient rows with a transposed butterfly
//    ("reduce-scatter": ~V shuffles for V values instead of 5V), after which 16/32 lanes each
//    hold one finished component and add the whole per-Gaussian gradient row with ONE vectorised
//    atomic instruction into a packed [P][RECF] row (2-4 sectors);
//  * warps are autonomous (no CTA barrier; see composite.cu): each starts at ITS pixel block's
//    largest n_contrib instead of the end of the tile list — entries behind every pixel's last
//    contributor are never loaded — and skips entries its block cannot see via touch_block();
//  * per-pixel state is register resident (template on channel groups); the reference keeps three
//    float[24] arrays in local memory and 36 KB of shared memory per CTA;
//  * gradients differ from the reference only by fp32 summation order (the reference's own
//    atomics make it run-to-run nondeterministic), tolerance 1e-3 relative.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace r3dg {

struct CompositeBwdParams {
    int W, H, gx, S, recf, backward_geometry;
    const uint2* ranges;
    const uint32_t* point_list;  // per-tile depth-sorted Gaussian ids (binning.cu)
    const uint32_t* cmask32;     // per-instance contributor masks written by the forward compositor
    const GeomHeader* header;
    const uint32_t* tile_order;  // CTA -> tile, heaviest tiles first
    const float* rec;
    const float* bg;
    const float* final_T;
    const int* n_contrib;
    const float *dL_dpix, *dL_dpix_o, *dL_dpix_d, *dL_dpix_f;
    float* grad;     // [P][recf], zero-initialised
};

// Sum v[i] over the 32 lanes for every i in [0,N) (N power of two, 4 <= N <= 32) with a
// transposed butterfly; on return lane l holds the total of component (l >> (5 - log2 N)).
template <int N>
__device__ __forceinline__ float reduce_scatter(float (&v)[N], int lane) {
    int m = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const bool upper = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
        m >>= 1;
    }
    float r = v[0];
#pragma unroll
    for (int mm = 16 / N; mm > 0; mm >>= 1) r += __shfl_xor_sync(0xffffffffu, r, mm);
    return r;
}

// Per warp (one 8x4 pixel block): the forward pass left one contributor bit per (instance, block).  The warp streams
// its tile's contributor bytes BACK TO FRONT, 128 instances per step, compacts the positions whose bit is set into a
// circular queue (warp scan) and processes batches of 32 of them: every staged entry is one that some pixel of the
// block composited in the forward pass — no culling test, no record fetch and no exp() for anything else.
template <int NG, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) composite_bwd_kernel(const CompositeBwdParams p) {
    constexpr int NC = 4 * NG;              // padded channel count {r,g,b,f...}
    constexpr int V = 8 + NC;               // gradient row width
    constexpr int V0 = V <= 16 ? 16 : 32;   // first butterfly chunk
    constexpr int V1 = V > 32 ? 4 : 0;      // tail chunk (only V == 36)
    constexpr int RG = 2 + NG;
    __shared__ float4 sRec[NW][RG][32];
    __shared__ uint32_t sId[NW][32];
    __shared__ uint32_t sQ[NW][R3DG_QCAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PARTS = 8 / NW;
    const int tile = (int)p.tile_order[blockIdx.x / PARTS];
    const int wb = (blockIdx.x % PARTS) * NW + warp;
    const int tx = tile % p.gx, ty = tile / p.gx;
    const int bx0 = tx * R3DG_TILE + (wb & 1) * 8, by0 = ty * R3DG_TILE + (wb >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = p.ranges[tile];
    const uint32_t lo = range.x;
    const size_t HW = (size_t)p.H * p.W, pix = (size_t)p.W * py + px;
    const float4* __restrict__ rec4 = reinterpret_cast<const float4*>(p.rec);
    const int rec4n = p.recf >> 2;
    const uint32_t* __restrict__ plist = p.point_list + lo;
    uint32_t* q = sQ[warp];

    const float T_final = inside ? p.final_T[pix] : 0.0f;
    float T = T_final;
    const int last_contributor = inside ? p.n_contrib[pix] : 0;
    // nothing behind the warp's deepest last contributor matters to this warp
    int total = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total = max(total, __shfl_xor_sync(0xffffffffu, total, o));
    total = min(total, (int)(range.y - range.x));
    if (total == 0) return;

    float dpix[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float g = 0.0f;
        if (inside) {
            if (c < 3) g = p.dL_dpix[c * HW + pix];
            else if (c - 3 < p.S) g = p.dL_dpix_f[(size_t)(c - 3) * HW + pix];
        }
        dpix[c] = g;
    }
    const float dpix_d = inside ? p.dL_dpix_d[pix] : 0.0f;
    const float dpix_o = inside ? p.dL_dpix_o[pix] : 0.0f;
    const float bg_dot = p.bg[0] * dpix[0] + p.bg[1] * dpix[1] + p.bg[2] * dpix[2];
    const float ddelx_dx = 0.5f * p.W, ddely_dy = 0.5f * p.H;
    const bool geo = p.backward_geometry != 0;
    // The reference keeps, per channel, the colour accumulated BEHIND the current entry (accum_rec[ch], backward.cu:
    // 541-566) only to form  dL_dalpha = sum_ch (c_ch - accum_rec[ch]) * dL_dpix[ch].  With D = sum_ch c_ch * dL_dpix[ch]
    // (the entry's own channels, depth and opacity terms included) the same quantity is  D - A  where the ONE scalar
    // A = sum_ch accum_rec[ch] * dL_dpix[ch] obeys the same recurrence  A <- alpha * D + (1 - alpha) * A : one FMA per
    // channel instead of four, and no per-channel state (20 registers at S = 16).  Gradients are compared at 1e-3.
    float A = 0.0f;
    // feature channels enter dL_dalpha only when backward_geometry is set (backward.cu:563-566)
    float dsel[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) dsel[c] = (c < 3 || geo) ? dpix[c] : 0.0f;

    // ---- contributor-mask stream, descending: word w holds instances [4w, 4w+4); lane takes word w_next - lane -------
    const uint32_t hi = lo + (uint32_t)total;                 // instances [lo, hi) matter
    const long long w_low = (long long)(lo >> 2);             // lowest word of the tile
    long long w_next = (long long)((hi - 1u) >> 2);           // highest word still to scan
    uint32_t m_nxt = (w_next - lane >= w_low) ? p.cmask32[w_next - lane] : 0u;
    int qhead = 0, qcount = 0;
    auto scan_step = [&]() {
        const long long w = w_next - lane;
        uint32_t flags = 0u;
        if (w >= w_low) {
            flags = (m_nxt >> wb) & 0x01010101u;
            const uint32_t e0 = (uint32_t)w << 2;
            if (e0 < lo) flags &= 0xffffffffu << (8 * (lo - e0));
            if (e0 + 4 > hi) flags &= 0xffffffffu >> (8 * (e0 + 4 - hi));
        }
        w_next -= 32;
        m_nxt = (w_next - lane >= w_low) ? p.cmask32[w_next - lane] : 0u;
        uint32_t tot;
        uint32_t slot = (uint32_t)(qhead + qcount) + warp_excl_scan(__popc(flags), lane, tot);
        const uint32_t rel = ((uint32_t)w << 2) - lo;
#pragma unroll
        for (int k = 3; k >= 0; --k)                                              // back to front inside the word too
            if (flags & (1u << (8 * k))) { q[slot & (R3DG_QCAP - 1)] = rel + k; ++slot; }
        qcount += (int)tot;
    };
    while (qcount < 64 && w_next >= w_low) scan_step();
    __syncwarp();

    uint32_t id_cur = 0u;
    float4 r[RG];
#pragma unroll
    for (int g = 0; g < RG; ++g) r[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < qcount) {
        id_cur = plist[q[(qhead + lane) & (R3DG_QCAP - 1)]];
#pragma unroll
        for (int g = 0; g < RG; ++g) r[g] = rec4[(size_t)id_cur * rec4n + g];
    }

    while (qcount > 0) {
        const int n = min(32, qcount);
        __syncwarp();
        sId[warp][lane] = id_cur * (uint32_t)p.recf;                   // element offset of the Gaussian's gradient row
#pragma unroll
        for (int g = 0; g < RG; ++g) sRec[warp][g][lane] = r[g];
        const int h0 = qhead;
        qhead = (qhead + n) & (R3DG_QCAP - 1);
        qcount -= n;
        while (qcount < 64 && w_next >= w_low) scan_step();
        __syncwarp();
        if (lane < qcount) {
            id_cur = plist[q[(qhead + lane) & (R3DG_QCAP - 1)]];
#pragma unroll
            for (int g = 0; g < RG; ++g) r[g] = rec4[(size_t)id_cur * rec4n + g];
        }
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const int k = (int)q[(h0 + j) & (R3DG_QCAP - 1)];                    // 0-based list position == contributor
            const float4 a = sRec[warp][0][j];
            const float4 b = sRec[warp][1][j];
            const float dx = sub_(a.x, pxf), dy = sub_(a.y, pyf);
            const float qd = fma_(dx, mul_(dx, a.z), mul_(dy, mul_(dy, b.x)));
            const float power = fma_(qd, -0.5f, -mul_(dy, mul_(dx, a.w)));
            const float G = expf(power);
            const float alpha = fminf(0.99f, mul_(b.y, G));
            // the same three tests as the forward pass (identical arithmetic) + "not behind this pixel's last contributor"
            const bool valid = k < last_contributor && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            // branch-free: every lane runs the arithmetic, lanes that did not composite this entry contribute zeros and
            // keep their state (some lane always did — that is what the contributor bit says)
            const float gate = valid ? 1.0f : 0.0f;
            float inv;                                                   // 1 / (1 - alpha), 1 - alpha in [0.01, 1]: one MUFU
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.0f - alpha));
            const float Tn = T * inv;                                   // T before this entry (backward.cu:533)
            const float dchannel_dcolor = gate * alpha * Tn;
            float v0[V0];
            float v1[V1 > 0 ? V1 : 4];
#pragma unroll
            for (int i = 0; i < V0; ++i) v0[i] = 0.0f;
#pragma unroll
            for (int i = 0; i < (V1 > 0 ? V1 : 4); ++i) v1[i] = 0.0f;
            float D = fmaf(b.z, dpix_d, dpix_o);                        // depth * dL_ddepth + 1 * dL_dopacity
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const float4 c4 = sRec[warp][2 + g][j];
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * g + e;
                    D = fmaf(cc[e], dsel[c], D);
                    const float gv = dchannel_dcolor * dpix[c];
                    if (8 + c < V0) v0[(8 + c) < V0 ? (8 + c) : 0] = gv;
                    else v1[(8 + c - V0) >= 0 && (8 + c - V0) < 4 ? (8 + c - V0) : 0] = gv;
                }
            }
            const float dL_dalpha = gate * fmaf(D - A, Tn, -(T_final * inv) * bg_dot);
            if (valid) { A = fmaf(alpha, D - A, A); T = Tn; }           // A <- alpha D + (1 - alpha) A
            const float dL_dG = b.y * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * a.z - gdy * a.w;
            const float dG_ddely = -gdy * b.x - gdx * a.w;
            v0[0] = dL_dG * dG_ddelx * ddelx_dx;
            v0[1] = dL_dG * dG_ddely * ddely_dy;
            v0[2] = dpix_d * dchannel_dcolor;
            v0[3] = G * dL_dalpha;
            v0[4] = -0.5f * gdx * dx * dL_dG;
            v0[5] = -0.5f * gdx * dy * dL_dG;
            v0[6] = -0.5f * gdy * dy * dL_dG;
            float* grow = p.grad + sId[warp][j];
            const float r0 = reduce_scatter<V0>(v0, lane);
            {
                constexpr int SH = V0 == 32 ? 0 : 1;            // lanes per component - 1 (log2)
                const int comp = lane >> SH;
                if ((lane & ((1 << SH) - 1)) == 0 && comp < V && comp != 7) atomicAdd(grow + comp, r0);
            }
            if constexpr (V1 > 0) {
                const float r1 = reduce_scatter<4>(v1, lane);
                const int comp = lane >> 3;
                if ((lane & 7) == 0 && V0 + comp < V) atomicAdd(grow + V0 + comp, r1);
            }
        }
    }
}

#ifndef R3DG_BWD_OCC2
#define R3DG_BWD_OCC2 6
#endif
#ifndef R3DG_BWD_OCC5
#define R3DG_BWD_OCC5 4
#endif
template <int NG> struct BwdOcc { static constexpr int v = NG <= 2 ? R3DG_BWD_OCC2 : (NG <= 3 ? 5 : (NG <= 5 ? R3DG_BWD_OCC5 : 3)); };
template <int NG>
static void launch_bwd_ng(const CompositeBwdParams& p, int tiles, cudaStream_t stream) {
    composite_bwd_kernel<NG, 4, BwdOcc<NG>::v><<<tiles * 2, 128, 0, stream>>>(p);      // two 4-warp CTAs per tile
}

int launch_composite_backward(const r3dg_raster_bwd_args& a, const GeomLayout& gl, const ImgLayout& il,
                              char* bin, const BinLayout& bl, cudaStream_t stream) {
    char* geom = (char*)a.geom;
    char* img = (char*)a.img;
    CompositeBwdParams p;
    p.W = a.W; p.H = a.H; p.gx = (a.W + R3DG_TILE - 1) / R3DG_TILE; p.S = a.S; p.recf = gl.recf;
    p.backward_geometry = a.backward_geometry;
    const int gy = (a.H + R3DG_TILE - 1) / R3DG_TILE;
    p.ranges = (const uint2*)(img + il.ranges);
    p.point_list = (const uint32_t*)(bin + bl.point_list); p.cmask32 = (const uint32_t*)(bin + bl.cmask);
    p.header = (const GeomHeader*)(geom + gl.header);
    p.tile_order = (const uint32_t*)(img + il.tile_order);
    p.rec = (const float*)(geom + gl.rec);
    p.bg = a.background;
    p.final_T = (const float*)(img + il.final_T);
    p.dL_dpix = a.dL_dout_color; p.dL_dpix_o = a.dL_dout_opacity; p.dL_dpix_d = a.dL_dout_depth;
    p.dL_dpix_f = a.dL_dout_feature;
    p.grad = (float*)(geom + gl.grad);
    p.n_contrib = (const int*)(img + il.n_contrib);
    R3DG_CUDA_TRY(cudaMemsetAsync(p.grad, 0, (size_t)a.P * gl.recf * 4, stream));
    const int tiles = p.gx * gy;
    switch (num_groups(a.S)) {
        case 1: launch_bwd_ng<1>(p, tiles, stream); break;
        case 2: launch_bwd_ng<2>(p, tiles, stream); break;
        case 3: launch_bwd_ng<3>(p, tiles, stream); break;
        case 4: launch_bwd_ng<4>(p, tiles, stream); break;
        case 5: launch_bwd_ng<5>(p, tiles, stream); break;
        case 6: launch_bwd_ng<6>(p, tiles, stream); break;
        case 7: launch_bwd_ng<7>(p, tiles, stream); break;
        default: return R3DG_ERR_UNSUPPORTED;   // S > 24: reference backward limit (backward.cu:449)
    }
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace r3dg
