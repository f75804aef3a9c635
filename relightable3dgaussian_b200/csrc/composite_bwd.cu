// Stage 2 backward: per-tile back-to-front re-traversal (reference K10, backward.cu:401-614).
//
// B200 design notes
//  * the reference issues 10+S global atomicAdds per contributing (pixel, Gaussian) pair.  Here a
//    warp (8x4 pixels) first reduces its 32 per-pixel gradient rows with transposed butterflies
//    ("reduce-scatter": ~V shuffles for V values instead of 5V), after which every lane holds one
//    finished component and the warp adds the whole per-Gaussian gradient row with ONE reduction
//    instruction into a packed [P][RECF] row (2-4 sectors).  The colour / feature components are
//    all  (alpha T) * dL_dpix[c] : each lane keeps its dL_dpix in a lane-specific PERMUTED order,
//    so that what a lane sends and what it keeps at every butterfly level sit in fixed registers —
//    no select instructions (reduce_scatter_permuted); only the 7 geometry components pay them;
//  * warps are autonomous (no CTA barrier; see composite.cu): each walks ONLY the entries whose
//    contributor bit the forward pass set for its pixel block, back to front from the block's
//    deepest last contributor — no culling test, no record fetch, no exp() for anything else;
//  * CTAs are launched in the order of the work the forward pass measured (cta_order_kernel), and
//    the walk is laid out in rounds of U entries for instruction-level parallelism: the kernel is
//    bound by per-warp latency and its heaviest warps (profiles/r02_warp_timing.md);
//  * per-pixel state is register resident (template on channel groups); the reference keeps three
//    float[24] arrays in local memory and 36 KB of shared memory per CTA;
//  * gradients differ from the reference only by fp32 summation order (the reference's own
//    atomics make it run-to-run nondeterministic), tolerance 1e-3 relative.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "kernels.h"

#ifndef R3DG_BWD_CTAS          // default resident CTAs per SM of the backward compositor (0 = whatever fits)
#define R3DG_BWD_CTAS 0
#endif
#ifndef R3DG_BWD_ILP           // entries per round of the backward walk (composite_bwd_kernel), S <= 5
#define R3DG_BWD_ILP 2
#endif
#ifndef R3DG_BWD_ILP_WIDE      // the same for S > 5 (twice the accumulators: registers)
#define R3DG_BWD_ILP_WIDE 1
#endif

namespace r3dg {

template <int NG> struct BwdIlp { static constexpr int v = NG <= 2 ? R3DG_BWD_ILP : R3DG_BWD_ILP_WIDE; };

struct CompositeBwdParams {
    int W, H, gx, S, recf, backward_geometry;
    const uint2* ranges;
    const uint32_t* point_list;  // per-tile depth-sorted Gaussian ids (binning.cu)
    const uint32_t* cmask32;     // per-instance contributor masks written by the forward compositor
    const GeomHeader* header;
    const uint32_t* cta_order;   // CTA -> half tile (2 * tile + part), most backward work first (cta_order_kernel)
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

// U independent reduce-scatters advanced level by level together: U shuffle chains in flight instead of one (the last
// levels of a single butterfly are one dependent SHFL -> FADD after the other).
template <int U, int N>
__device__ __forceinline__ void reduce_scatter_multi(float (&v)[U][N], int lane, float (&r)[U]) {
    int m = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const bool upper = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float send = upper ? v[u][i] : v[u][i + n / 2];
                const float keep = upper ? v[u][i + n / 2] : v[u][i];
                v[u][i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
            }
        }
        m >>= 1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = v[u][0];
#pragma unroll
    for (int mm = 16 / N; mm > 0; mm >>= 1) {
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] += __shfl_xor_sync(0xffffffffu, r[u], mm);
    }
}

// The same reduce-scatter without selects, for W values whose slot order the caller chose per lane: slot p of lane l
// must hold component  p ^ (~(l >> (5 - log2 W)) & (W - 1)).  At level n every lane sends slots [0, n/2) and keeps
// [n/2, n): with that order the kept slot i + n/2 and the partner's sent slot i are the same component.  On return
// lane l holds the total of component  l >> (5 - log2 W)  (as reduce_scatter).
template <int W>
__device__ __forceinline__ float reduce_scatter_permuted(float (&x)[W]) {
    int m = 16;
#pragma unroll
    for (int n = W; n > 1; n >>= 1) {
#pragma unroll
        for (int i = 0; i < n / 2; ++i) x[i] = x[i + n / 2] + __shfl_xor_sync(0xffffffffu, x[i], m);
        m >>= 1;
    }
    float r = x[0];
#pragma unroll
    for (int mm = 16 / W; mm > 0; mm >>= 1) r += __shfl_xor_sync(0xffffffffu, r, mm);
    return r;
}

#ifdef R3DG_WARP_TIMING
__device__ WarpTiming g_wt_bwd[R3DG_WT_MAX];
extern "C" int r3dg_debug_wt_bwd(void* dst, size_t bytes) { return (int)cudaMemcpyFromSymbol(dst, g_wt_bwd, bytes); }
#endif

template <int U, int W>
__device__ __forceinline__ void reduce_scatter_permuted_multi(float (&x)[U][W], float (&r)[U]) {
    int m = 16;
#pragma unroll
    for (int n = W; n > 1; n >>= 1) {
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
#pragma unroll
            for (int u = 0; u < U; ++u) x[u][i] = x[u][i + n / 2] + __shfl_xor_sync(0xffffffffu, x[u][i], m);
        }
        m >>= 1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = x[u][0];
#pragma unroll
    for (int mm = 16 / W; mm > 0; mm >>= 1) {
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] += __shfl_xor_sync(0xffffffffu, r[u], mm);
    }
}

// Per warp (one 8x4 pixel block): the forward pass left one contributor bit per (instance, block).  The warp streams
// its tile's contributor bytes BACK TO FRONT, 128 instances per step, compacts the positions whose bit is set into a
// circular queue (warp scan) and processes batches of 32 of them: every staged entry is one that some pixel of the
// block composited in the forward pass — no culling test, no record fetch and no exp() for anything else.
template <int NG, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) composite_bwd_kernel(const CompositeBwdParams p) {
    constexpr int NC = 4 * NG;              // padded channel count {r,g,b,f...}; gradient row = 8 geometry slots + NC
    // the NC channel components are reduced in power-of-two chunks of 16 / 8 / 4 (select-free butterflies)
    constexpr int C16 = NC >= 16 ? 16 : 0, C8 = (NC - C16) >= 8 ? 8 : 0, C4 = NC - C16 - C8;
    constexpr int B8 = C16, B4 = C16 + C8;
    static_assert(C4 == 0 || C4 == 4, "channel groups come in fours");
    constexpr int RG = 2 + NG;
    constexpr int U = BwdIlp<NG>::v;                 // entries per round of the walk below
    __shared__ float4 sRec[NW][RG][32];
    __shared__ uint32_t sId[NW][32];
    __shared__ uint32_t sK[NW][32];
    __shared__ uint32_t sQ[NW][R3DG_QCAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PARTS = 8 / NW;
    static_assert(PARTS == 2, "bwd_work / cta_order are per half tile");
    const int cta = (int)p.cta_order[blockIdx.x];
    const int tile = cta / PARTS;
    const int wb = (cta % PARTS) * NW + warp;
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

#ifdef R3DG_WARP_TIMING
    const unsigned long long wt_t0 = wt_now();
    unsigned wt_iters = 0;
    auto wt_done = [&]() {
        if (lane == 0 && blockIdx.x * NW + warp < R3DG_WT_MAX) g_wt_bwd[blockIdx.x * NW + warp] = WarpTiming{wt_t0, wt_now(), wt_iters, wt_smid()};
    };
#endif
    const float T_final = inside ? p.final_T[pix] : 0.0f;
    float T = T_final;
    const int last_contributor = inside ? p.n_contrib[pix] : 0;
    // nothing behind the warp's deepest last contributor matters to this warp
    int total = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total = max(total, __shfl_xor_sync(0xffffffffu, total, o));
    total = min(total, (int)(range.y - range.x));
#ifdef R3DG_WARP_TIMING
    if (total == 0) wt_done();
#endif
    if (total == 0) return;

    auto load_dpix = [&](int c) -> float {                   // cotangent of channel c of {r,g,b,f0..} at this lane's pixel
        if (!inside) return 0.0f;
        if (c < 3) return p.dL_dpix[c * HW + pix];
        return c - 3 < p.S ? p.dL_dpix_f[(size_t)(c - 3) * HW + pix] : 0.0f;
    };
    float dpix[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) dpix[c] = load_dpix(c);
    // the same cotangents in this lane's butterfly slot order (reduce_scatter_permuted), one array per chunk
    float cst16[C16 ? C16 : 1], cst8[C8 ? C8 : 1], cst4[C4 ? C4 : 1];
    if constexpr (C16 > 0) {
        const int K = ~(lane >> 1) & 15;
#pragma unroll
        for (int i = 0; i < 16; ++i) cst16[i] = load_dpix(i ^ K);
    }
    if constexpr (C8 > 0) {
        const int K = ~(lane >> 2) & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) cst8[i] = load_dpix(B8 + (i ^ K));
    }
    if constexpr (C4 > 0) {
        const int K = ~(lane >> 3) & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) cst4[i] = load_dpix(B4 + (i ^ K));
    }
    // which finished component this lane adds to the Gaussian's gradient row: geometry slot lane >> 2 on lanes 0 mod 4,
    // the 16-chunk on odd lanes, the 8-chunk on lanes 2 mod 4, the 4-chunk on a free lane class (second instruction
    // only when all three chunks exist, NG == 7)
    constexpr bool C4_SECOND = C16 > 0 && C8 > 0 && C4 > 0;
    int my_sel = -1, my_off = 0;
    if ((lane & 3) == 0) { if ((lane >> 2) != 7) { my_sel = 0; my_off = lane >> 2; } }
    else if (C16 > 0 && (lane & 1)) { my_sel = 1; my_off = 8 + (lane >> 1); }
    else if (C8 > 0 && (lane & 3) == 2) { my_sel = 2; my_off = 8 + B8 + (lane >> 2); }
    if (C4 > 0 && !C4_SECOND && (lane & 7) == (C16 > 0 ? 2 : 1)) { my_sel = 3; my_off = 8 + B4 + (lane >> 3); }
    if (my_sel > 0 && my_off - 8 >= 3 + p.S) my_sel = -1;    // zero padding channel
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

    uint32_t id_cur = 0u, k_cur = 0u;
    float4 r[RG];
#pragma unroll
    for (int g = 0; g < RG; ++g) r[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < qcount) {
        k_cur = q[(qhead + lane) & (R3DG_QCAP - 1)];
        id_cur = plist[k_cur];
#pragma unroll
        for (int g = 0; g < RG; ++g) r[g] = rec4[(size_t)id_cur * rec4n + g];
    }

    while (qcount > 0) {
        const int n = min(32, qcount);
        __syncwarp();
        sId[warp][lane] = id_cur * (uint32_t)p.recf;                   // element offset of the Gaussian's gradient row
        sK[warp][lane] = k_cur;                                        // 0-based list position == contributor index
#pragma unroll
        for (int g = 0; g < RG; ++g) sRec[warp][g][lane] = r[g];
        qhead = (qhead + n) & (R3DG_QCAP - 1);
        qcount -= n;
        while (qcount < 64 && w_next >= w_low) scan_step();
        __syncwarp();
        if (lane < qcount) {
            k_cur = q[(qhead + lane) & (R3DG_QCAP - 1)];
            id_cur = plist[k_cur];
#pragma unroll
            for (int g = 0; g < RG; ++g) r[g] = rec4[(size_t)id_cur * rec4n + g];
        }
        // Rounds of U entries.  The kernel is latency-bound, not issue-bound (fewer resident CTAs make it slower at every
        // setting, profiles/r02_warp_timing.md), and the heaviest pixel blocks walk ~500 contributors one after the other,
        // so the round is laid out for instruction-level parallelism inside the warp: (1) everything that depends only on
        // the entry — record fetch, exp(), alpha, 1/(1-alpha), the entry's own cotangent dot D — for all U entries; (2) the
        // two short recurrences T and A in list order; (3) gradient rows and butterflies of the U entries, advanced level
        // by level together.  A round with fewer than U entries left re-evaluates the last one with its gate closed.
#pragma unroll 1
        for (int j0 = 0; j0 < n; j0 += U) {
            int ju[U];
            bool valid[U];
            float dx[U], dy[U], G[U], alpha[U], inv[U], D[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {                                   // (1)
                ju[u] = min(j0 + u, n - 1);
                const int k = (int)sK[warp][ju[u]];
#ifdef R3DG_WARP_TIMING
                wt_iters += j0 + u < n ? 1u : 0u;
#endif
                const float4 a = sRec[warp][0][ju[u]];
                const float4 b = sRec[warp][1][ju[u]];
                dx[u] = sub_(a.x, pxf); dy[u] = sub_(a.y, pyf);
                const float qd = fma_(dx[u], mul_(dx[u], a.z), mul_(dy[u], mul_(dy[u], b.x)));
                const float power = fma_(qd, -0.5f, -mul_(dy[u], mul_(dx[u], a.w)));
                G[u] = expf(power);
                alpha[u] = fminf(0.99f, mul_(b.y, G[u]));
                // the same three tests as the forward pass (identical arithmetic) + "not behind this pixel's last contributor"
                valid[u] = j0 + u < n && k < last_contributor && !(power > 0.0f) && !(alpha[u] < 1.0f / 255.0f);
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv[u]) : "f"(1.0f - alpha[u]));   // 1 - alpha in [0.01, 1]: one MUFU
                float d = fmaf(b.z, dpix_d, dpix_o);                        // depth * dL_ddepth + 1 * dL_dopacity
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const float4 c4 = sRec[warp][2 + g][ju[u]];
                    d = fmaf(c4.x, dsel[4 * g + 0], d);
                    d = fmaf(c4.y, dsel[4 * g + 1], d);
                    d = fmaf(c4.z, dsel[4 * g + 2], d);
                    d = fmaf(c4.w, dsel[4 * g + 3], d);
                }
                D[u] = d;
            }
            float dcc[U], dL_dalpha[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {                                   // (2) branch-free: lanes that did not composite
                const float gate = valid[u] ? 1.0f : 0.0f;                  // the entry contribute zeros and keep their state
                const float Tn = T * inv[u];                                // T before this entry (backward.cu:533)
                dcc[u] = gate * alpha[u] * Tn;
                dL_dalpha[u] = gate * fmaf(D[u] - A, Tn, -(T_final * inv[u]) * bg_dot);
                if (valid[u]) { A = fmaf(alpha[u], D[u] - A, A); T = Tn; } // A <- alpha D + (1 - alpha) A
            }
            float v0[U][8], mine[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {                                   // (3)
                const float4 a = sRec[warp][0][ju[u]];
                const float4 b = sRec[warp][1][ju[u]];
                const float dL_dG = b.y * dL_dalpha[u];
                const float gdx = G[u] * dx[u], gdy = G[u] * dy[u];
                const float dG_ddelx = -gdx * a.z - gdy * a.w;
                const float dG_ddely = -gdy * b.x - gdx * a.w;
                v0[u][0] = dL_dG * dG_ddelx * ddelx_dx;
                v0[u][1] = dL_dG * dG_ddely * ddely_dy;
                v0[u][2] = dpix_d * dcc[u];
                v0[u][3] = G[u] * dL_dalpha[u];
                v0[u][4] = -0.5f * gdx * dx[u] * dL_dG;
                v0[u][5] = -0.5f * gdx * dy[u] * dL_dG;
                v0[u][6] = -0.5f * gdy * dy[u] * dL_dG;
                v0[u][7] = 0.0f;
            }
            reduce_scatter_multi<U, 8>(v0, lane, mine);                     // geometry slots: lanes 0 mod 4
            if constexpr (C16 > 0) {
                float x[U][16], t[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[u][i] = dcc[u] * cst16[i];
                reduce_scatter_permuted_multi<U, 16>(x, t);
#pragma unroll
                for (int u = 0; u < U; ++u) if (my_sel == 1) mine[u] = t[u];
            }
            if constexpr (C8 > 0) {
                float x[U][8], t[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[u][i] = dcc[u] * cst8[i];
                reduce_scatter_permuted_multi<U, 8>(x, t);
#pragma unroll
                for (int u = 0; u < U; ++u) if (my_sel == 2) mine[u] = t[u];
            }
            float t4[U];
            if constexpr (C4 > 0) {
                float x[U][4];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[u][i] = dcc[u] * cst4[i];
                reduce_scatter_permuted_multi<U, 4>(x, t4);
                if constexpr (!C4_SECOND) {
#pragma unroll
                    for (int u = 0; u < U; ++u) if (my_sel == 3) mine[u] = t4[u];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (j0 + u < n) {                                           // (warp-uniform)
                    float* grow = p.grad + sId[warp][ju[u]];
                    if constexpr (C4_SECOND) {
                        if ((lane & 7) == 0 && B4 + (lane >> 3) < 3 + p.S) atomicAdd(grow + 8 + B4 + (lane >> 3), t4[u]);
                    }
                    if (my_sel >= 0) atomicAdd(grow + my_off, mine[u]);
                }
            }
        }
    }
#ifdef R3DG_WARP_TIMING
    wt_done();
#endif
}

#ifndef R3DG_BWD_OCC2
#define R3DG_BWD_OCC2 5
#endif
#ifndef R3DG_BWD_OCC5
#define R3DG_BWD_OCC5 4
#endif
template <int NG> struct BwdOcc { static constexpr int v = NG <= 2 ? R3DG_BWD_OCC2 : (NG <= 3 ? 5 : (NG <= 5 ? R3DG_BWD_OCC5 : 3)); };
int g_bwd_ctas = -1;            // r3dg_tune("composite_bwd_ctas"): resident CTAs per SM (0 = whatever fits)
static void composite_bwd_env() {
    if (g_bwd_ctas < 0) { const char* e = getenv("R3DG_BWD_CTAS"); g_bwd_ctas = e ? atoi(e) : R3DG_BWD_CTAS; if (g_bwd_ctas < 0) g_bwd_ctas = 0; }
}
int composite_bwd_tune(const char* key, int value, int* previous) {
    composite_bwd_env();
    if (strcmp(key, "composite_bwd_ctas") != 0) return R3DG_ERR_UNSUPPORTED;
    if (previous) *previous = g_bwd_ctas;
    if (value < 0 || value > 32) return R3DG_ERR_BAD_ARG;
    g_bwd_ctas = value;
    return 0;
}
template <int NG>
static void launch_bwd_ng(const CompositeBwdParams& p, int tiles, cudaStream_t stream) {
    composite_bwd_env();
    auto kernel = composite_bwd_kernel<NG, 4, BwdOcc<NG>::v>;
    static int pad_for = -1, pad_dev = -1;
    static size_t pad = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (pad_for != g_bwd_ctas || pad_dev != dev) { pad = residency_pad(kernel, g_bwd_ctas); pad_for = g_bwd_ctas; pad_dev = dev; }
    kernel<<<tiles * 2, 128, pad, stream>>>(p);                          // two 4-warp CTAs per tile
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
    p.cta_order = (const uint32_t*)(img + il.bwd_order);
    p.rec = (const float*)(geom + gl.rec);
    p.bg = a.background;
    p.final_T = (const float*)(img + il.final_T);
    p.dL_dpix = a.dL_dout_color; p.dL_dpix_o = a.dL_dout_opacity; p.dL_dpix_d = a.dL_dout_depth;
    p.dL_dpix_f = a.dL_dout_feature;
    p.grad = (float*)(geom + gl.grad);
    p.n_contrib = (const int*)(img + il.n_contrib);
    R3DG_CUDA_TRY(cudaMemsetAsync(p.grad, 0, (size_t)a.P * gl.recf * 4, stream));
    const int tiles = p.gx * gy;
    int rc = launch_cta_order((const uint32_t*)(img + il.bwd_work), (uint32_t*)(img + il.bwd_order), 2 * tiles, stream);
    if (rc != 0) return rc;
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
