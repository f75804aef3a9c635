// Stage 3a of the hot path: the BRDF shading step `rendering_equation` (reference
// gaussian_renderer/neilf.py:339-371) with `GGX_specular` (:374-406), the lat-long environment
// lookup (scene/direct_light_map.py:70-83, scene/envmap.py:35-53) and `eval_sh` (utils/sh_utils.py
// :71-128) — ~40 elementwise PyTorch kernels over [P,N,3] tensors in the reference, whose autograd
// graph keeps >= 15 such tensors alive — as ONE forward and ONE backward kernel.
//
// B200 design notes
//  * a group of 8 lanes per Gaussian (4 Gaussians per warp), lanes stride the N incident samples:
//    the baked [P,N,*] tensors (direction 12 B + visibility 4 B + area 4 B per sample — the only
//    per-sample HBM traffic) are read once per direction, rows of neighbouring Gaussians are
//    contiguous so the warp's loads stay coalesced; nothing of size [P,N,*] is written unless the
//    caller asks for the eval-only per-sample lights.  (One warp per Gaussian left 1-2 samples per
//    lane at the training sample counts N = 32 / 64 and paid 290 shuffles per Gaussian.)
//  * per-Gaussian operands (material, normal, view direction, 48 SH coefficients) live in
//    registers; the backward reduces its 48 SH-coefficient gradients with a transposed butterfly
//    (42 shuffles per 4 Gaussians) and the group writes the 192 B row contiguously;
//  * the (small) environment texture lives in shared memory of a persistent CTA; its gradient —
//    grid_sample's scatter-add of P*N*4 taps — is accumulated in warp-private shared-memory copies
//    without atomics (see "environment-map gradient" below), summed and flushed once per CTA;
//  * the backward recomputes the forward per sample instead of saving [P,N,*] activations.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "kernels.h"

namespace r3dg {

#define SHADE_THREADS 128
#define SHADE_WARPS (SHADE_THREADS / 32)
#define PI_F 3.14159274101257324f     // float32(np.pi)

struct ShadeArgs {
    int P, N, He, We, env_in_smem;
    const float *base_color, *roughness, *normals, *viewdirs, *incidents, *env, *transform;
    const float *visibility, *dirs, *areas;
    // forward outputs
    float *pbr, *diffuse, *specular, *mean_lights, *mean_local, *mean_global, *mean_vis;
    float *s_lights, *s_local, *s_global;        // optional per-sample [P,N,3]
    // backward
    const float *g_pbr, *g_diffuse, *g_specular;
    float *d_base, *d_rough, *d_view, *d_incidents, *d_env;
};

__device__ __forceinline__ void sh_basis3(float x, float y, float z, float* w) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    w[0] = 0.28209479177387814f;
    w[1] = -0.4886025119029199f * y; w[2] = 0.4886025119029199f * z; w[3] = -0.4886025119029199f * x;
    w[4] = 1.0925484305920792f * xy; w[5] = -1.0925484305920792f * yz; w[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    w[7] = -1.0925484305920792f * xz; w[8] = 0.5462742152960396f * (xx - yy);
    w[9] = -0.5900435899266435f * y * (3.f * xx - yy); w[10] = 2.890611442640554f * xy * z;
    w[11] = -0.4570457994644658f * y * (4.f * zz - xx - yy); w[12] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
    w[13] = -0.4570457994644658f * x * (4.f * zz - xx - yy); w[14] = 1.445305721320277f * z * (xx - yy);
    w[15] = -0.5900435899266435f * x * (xx - 3.f * yy);
}

struct EnvTap { int idx[4]; float w[4]; };      // texel base offsets (float index of channel 0, or -1) + weights

// grid_sample(bilinear, zeros padding, align_corners=True) at the lat-long coordinates of dir d
__device__ __forceinline__ EnvTap env_taps(float dx, float dy, float dz, int He, int We) {
    const float phi = acosf(dz) - 1e-6f;
    const float theta = atan2f(dy, dx);
    const float qy = (phi / PI_F) * 2.0f - 1.0f;
    const float qx = -theta / PI_F;
    const float ix = ((qx + 1.0f) / 2.0f) * (float)(We - 1);
    const float iy = ((qy + 1.0f) / 2.0f) * (float)(He - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float ax = ix - fx, ay = iy - fy;        // == ix - ix_nw ; (ix_se - ix) == 1 - ax
    EnvTap t;
    const int xs[2] = {x0, x0 + 1}, ys[2] = {y0, y0 + 1};
    const float wx[2] = {(fx + 1.0f) - ix, ax}, wy[2] = {(fy + 1.0f) - iy, ay};
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const bool in = xs[i] >= 0 && xs[i] < We && ys[j] >= 0 && ys[j] < He;
            t.idx[2 * j + i] = in ? (ys[j] * We + xs[i]) * 3 : -1;
            t.w[2 * j + i] = wx[i] * wy[j];
        }
    return t;
}

struct GaussianOperands {
    float base[3], rough, n[3], Ns[3], V[3], vnorm, nx_raw[3];
    float a2, k;
    float inc[48];
};

template <bool INC_REGS>
__device__ __forceinline__ void load_operands(const ShadeArgs& a, int g, GaussianOperands& o) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { o.base[c] = a.base_color[3 * (size_t)g + c]; o.n[c] = a.normals[3 * (size_t)g + c]; }
    o.rough = a.roughness[g];
    const float v0 = a.viewdirs[3 * (size_t)g], v1 = a.viewdirs[3 * (size_t)g + 1], v2 = a.viewdirs[3 * (size_t)g + 2];
    o.vnorm = fmaxf(sqrtf(v0 * v0 + v1 * v1 + v2 * v2), 1e-12f);           // F.normalize eps
    o.V[0] = v0 / o.vnorm; o.V[1] = v1 / o.vnorm; o.V[2] = v2 / o.vnorm;
    const float nn = fmaxf(sqrtf(o.n[0] * o.n[0] + o.n[1] * o.n[1] + o.n[2] * o.n[2]), 1e-12f);
    const float N0 = o.n[0] / nn, N1 = o.n[1] / nn, N2 = o.n[2] / nn;
    const float nov = o.V[0] * N0 + o.V[1] * N1 + o.V[2] * N2;
    const float sg = nov > 0.f ? 1.f : (nov < 0.f ? -1.f : 0.f);             // torch.sign
    o.Ns[0] = N0 * sg; o.Ns[1] = N1 * sg; o.Ns[2] = N2 * sg;
    const float al = o.rough * o.rough;
    o.a2 = al * al;
    o.k = (al + 2.0f * o.rough + 1.0f) / 8.0f;
    if (INC_REGS) {
        const float4* ip = reinterpret_cast<const float4*>(a.incidents + 48 * (size_t)g);
#pragma unroll
        for (int q = 0; q < 12; ++q) { const float4 v = ip[q]; o.inc[4 * q] = v.x; o.inc[4 * q + 1] = v.y; o.inc[4 * q + 2] = v.z; o.inc[4 * q + 3] = v.w; }
    }
}

// Kernel variants (r3dg_tune "shade_fwd_variant" / "shade_bwd_variant"): where the per-Gaussian SH
// state lives.  0: registers.  1: the 48 incident-light coefficients of each group in shared memory
// (broadcast LDS.128 reads, -48 registers).  2 (backward): additionally the 48 per-lane gradient
// accumulators in shared memory ([48][threads], conflict-free), loaded back once per Gaussian for
// the group reduction.  Fewer registers = more resident warps for a kernel that is bound by the
// latency of its dependent arithmetic (acos / atan2 / divisions), not by HBM.
#define SHADE_INC_STRIDE 52       // floats per group slot: 16-byte aligned, bank-staggered (52 mod 32 = 20)

template <int G>
__device__ __forceinline__ void stage_incidents(const ShadeArgs& a, int g, float* slot, int sub) {
    const float4* src = reinterpret_cast<const float4*>(a.incidents + 48 * (size_t)g);
    float4* dst = reinterpret_cast<float4*>(slot);
    for (int q = sub; q < 12; q += G) dst[q] = src[q];
    __syncwarp();
}

// 1/x with one MUFU (<= 1 ulp), for the BACKWARD kernel only: its recomputed forward quantities and its gradient
// arithmetic tolerate a last-bit difference (tests: 1e-4 relative), while the ~19 IEEE divisions per sample cost ~230 of
// its ~1500 instructions (FCHK + slow-path branch each; profiles/r02_ncu_shade_bwd.md).  The forward kernel keeps the
// IEEE divisions of the PyTorch expression.
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define INV_PI_F 0.318309873342514038f    // float32(1 / float32(np.pi))

struct SampleEval {       // forward quantities of one (Gaussian, direction) pair
    float local_raw[3], glob[3], ndi, fs;
    float w[16];
    EnvTap taps;
    // GGX intermediates for the backward
    float H[3], hn, NoL, NoV, NoH, VoH, NoL_r, NoV_r, NoH_r, VoH_r, p2, frac0, nom0, nom1, nom2, nom_r, nom;
};

#ifndef R3DG_SHADE_FWD_FAST        // experiment knob: one-MUFU reciprocals in the FORWARD kernel too (default: IEEE divisions)
#define R3DG_SHADE_FWD_FAST false
#endif
template <bool SMEM_ENV, bool INC_SMEM, bool FAST = R3DG_SHADE_FWD_FAST>
__device__ __forceinline__ void eval_sample(const ShadeArgs& a, const GaussianOperands& o, const float* __restrict__ env,
                                            const float* __restrict__ sinc, float dx, float dy, float dz, float vis, SampleEval& e) {
    // ---- environment + local SH light --------------------------------------------------------
    float ex = dx, ey = dy, ez = dz;
    if (a.transform) {                                     // dirs @ transform.T  (scene/envmap.py:39-42)
        const float* T = a.transform;
        ex = dx * T[0] + dy * T[1] + dz * T[2]; ey = dx * T[3] + dy * T[4] + dz * T[5]; ez = dx * T[6] + dy * T[7] + dz * T[8];
    }
    e.taps = env_taps(ex, ey, ez, a.He, a.We);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) if (e.taps.idx[t] >= 0) s += (SMEM_ENV ? env[e.taps.idx[t] + c] : __ldg(env + e.taps.idx[t] + c)) * e.taps.w[t];
        e.glob[c] = s * vis;
    }
    sh_basis3(dx, dy, dz, e.w);
    if (INC_SMEM) {
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            const float4 v = reinterpret_cast<const float4*>(sinc)[q];
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[(4 * q + j) % 3] += e.w[(4 * q + j) / 3] * vv[j];   // per channel still k = 0..15 in order
            if (q % 3 == 2) asm volatile("" ::: "memory");     // keep at most 3 LDS.128 in flight: hoisting all 12 costs the 48 registers back
        }
        e.local_raw[0] = acc[0]; e.local_raw[1] = acc[1]; e.local_raw[2] = acc[2];
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) s += e.w[k] * o.inc[3 * k + c];      // eval_sh's left-to-right order
            e.local_raw[c] = s;
        }
    }
    e.ndi = fmaxf(o.n[0] * dx + o.n[1] * dy + o.n[2] * dz, 0.f);
    // ---- GGX_specular (neilf.py:374-406) ----------------------------------------------------------
    const float ln = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
    float L0, L1, L2;
    if (FAST) { const float il = fast_rcp(ln); L0 = dx * il; L1 = dy * il; L2 = dz * il; }
    else { L0 = dx / ln; L1 = dy / ln; L2 = dz / ln; }
    const float h0 = (L0 + o.V[0]) / 2.0f, h1 = (L1 + o.V[1]) / 2.0f, h2 = (L2 + o.V[2]) / 2.0f;
    e.hn = fmaxf(sqrtf(h0 * h0 + h1 * h1 + h2 * h2), 1e-12f);
    if (FAST) { const float ih = fast_rcp(e.hn); e.H[0] = h0 * ih; e.H[1] = h1 * ih; e.H[2] = h2 * ih; }
    else { e.H[0] = h0 / e.hn; e.H[1] = h1 / e.hn; e.H[2] = h2 / e.hn; }
    e.NoL_r = o.Ns[0] * L0 + o.Ns[1] * L1 + o.Ns[2] * L2;
    e.NoV_r = o.Ns[0] * o.V[0] + o.Ns[1] * o.V[1] + o.Ns[2] * o.V[2];
    e.NoH_r = o.Ns[0] * e.H[0] + o.Ns[1] * e.H[1] + o.Ns[2] * e.H[2];
    e.VoH_r = o.V[0] * e.H[0] + o.V[1] * e.H[1] + o.V[2] * e.H[2];
    e.NoL = fminf(fmaxf(e.NoL_r, 1e-6f), 1.f); e.NoV = fminf(fmaxf(e.NoV_r, 1e-6f), 1.f);
    e.NoH = fminf(fmaxf(e.NoH_r, 1e-6f), 1.f); e.VoH = fminf(fmaxf(e.VoH_r, 1e-6f), 1.f);
    const float FMi = (-5.55473f * e.VoH - 6.98316f) * e.VoH;
    e.p2 = exp2f(FMi);
    e.frac0 = 0.04f + (1.0f - 0.04f) * e.p2;
    e.nom0 = e.NoH * e.NoH * (o.a2 - 1.0f) + 1.0f;
    e.nom1 = e.NoV * (1.0f - o.k) + o.k;
    e.nom2 = e.NoL * (1.0f - o.k) + o.k;
    e.nom_r = 4.0f * PI_F * e.nom0 * e.nom0 * e.nom1 * e.nom2;
    e.nom = fminf(fmaxf(e.nom_r, 1e-6f), 4.0f * PI_F);
    e.fs = FAST ? e.frac0 * o.a2 * fast_rcp(e.nom) : e.frac0 * o.a2 / e.nom;
}

// ---- sub-warp groups ---------------------------------------------------------------------------
// A group of G lanes (G = 8, 16 or 32) owns one Gaussian, so a warp shades 32/G Gaussians at once:
// at the training sample counts (N = 32 / 64, script/run_*.sh) a whole warp per Gaussian leaves each
// lane 1-2 samples and the per-Gaussian work (operand loads, 58 values x 5 shuffle levels in the
// backward) dominates.  Groups read consecutive rows of the baked tensors, so loads stay coalesced.
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Transposed butterfly ("reduce-scatter") over the G lanes of a group: V values per lane in, and
// lane `sub` returns with v[0 .. V/G) = the group totals of components [sub*V/G, (sub+1)*V/G).
// V*(1 - 1/G) shuffles instead of V*log2(G).
template <int G, int V>
__device__ __forceinline__ void group_reduce_scatter(float (&v)[V], int sub) {
    static_assert(V % G == 0, "V must be a multiple of the group width");
    int n = V;
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1) {
        const bool upper = (sub & m) != 0;
        const int h = n / 2;
#pragma unroll
        for (int i = 0; i < V / 2; ++i) {
            if (i < h) {
                const float send = upper ? v[i] : v[i + h];
                const float keep = upper ? v[i + h] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
            }
        }
        n = h;
    }
}

template <bool SMEM_ENV, int G, int VAR>
__global__ void __launch_bounds__(SHADE_THREADS, VAR == 1 ? 6 : 1) shade_fwd_kernel(const ShadeArgs a) {
    extern __shared__ __align__(16) float s_fwd[];     // [incident slots (VAR 1)][env texture (SMEM_ENV)]
    constexpr int GPW = 32 / G;                    // Gaussians per warp
    constexpr int INC_FLOATS = VAR >= 1 ? SHADE_WARPS * GPW * SHADE_INC_STRIDE : 0;
    float* s_env = s_fwd + INC_FLOATS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & (G - 1), grp = lane / G;
    float* sinc = s_fwd + (warp * GPW + grp) * SHADE_INC_STRIDE;
    if (SMEM_ENV) {
        for (int i = threadIdx.x; i < a.He * a.We * 3; i += SHADE_THREADS) s_env[i] = a.env[i];
        __syncthreads();
    }
    const float* env = SMEM_ENV ? s_env : a.env;
    const float invN = 1.0f / (float)a.N;
    const int units = (a.P + GPW - 1) / GPW;
    for (int u = blockIdx.x * SHADE_WARPS + warp; u < units; u += gridDim.x * SHADE_WARPS) {
        const int g_raw = u * GPW + grp;
        const bool valid = g_raw < a.P;
        const int g = valid ? g_raw : a.P - 1;      // idle groups of the last warp shade a copy, write nothing
        GaussianOperands o;
        load_operands<VAR == 0>(a, g, o);
        if (VAR >= 1) stage_incidents<G>(a, g, sinc, sub);
        float pbr[3] = {0, 0, 0}, spec[3] = {0, 0, 0}, diff[3] = {0, 0, 0}, ml[3] = {0, 0, 0}, mloc[3] = {0, 0, 0}, mg[3] = {0, 0, 0}, mv = 0;
        const size_t row = (size_t)g * a.N;
        for (int i = sub; i < a.N; i += G) {
            const float dx = a.dirs[3 * (row + i)], dy = a.dirs[3 * (row + i) + 1], dz = a.dirs[3 * (row + i) + 2];
            const float vis = a.visibility[row + i], area = a.areas[row + i];
            SampleEval e;
            eval_sample<SMEM_ENV, (VAR >= 1)>(a, o, env, sinc, dx, dy, dz, vis, e);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float loc = fmaxf(e.local_raw[c], 0.f);
                const float Lc = loc + e.glob[c];
                const float T = Lc * area * e.ndi;
                const float fd = o.base[c] / PI_F;
                pbr[c] += (fd + e.fs) * T; spec[c] += e.fs * T; diff[c] += T;
                ml[c] += Lc; mloc[c] += loc; mg[c] += e.glob[c];
                if (a.s_lights && valid) { a.s_lights[3 * (row + i) + c] = Lc; a.s_local[3 * (row + i) + c] = loc; a.s_global[3 * (row + i) + c] = e.glob[c]; }
            }
            mv += vis;
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            pbr[c] = group_sum<G>(pbr[c]); spec[c] = group_sum<G>(spec[c]); diff[c] = group_sum<G>(diff[c]);
            if (a.mean_lights) { ml[c] = group_sum<G>(ml[c]); mloc[c] = group_sum<G>(mloc[c]); mg[c] = group_sum<G>(mg[c]); }
        }
        if (a.mean_vis) mv = group_sum<G>(mv);
        if (sub == 0 && valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                a.pbr[3 * (size_t)g + c] = pbr[c] * invN; a.specular[3 * (size_t)g + c] = spec[c] * invN; a.diffuse[3 * (size_t)g + c] = diff[c] * invN;
                if (a.mean_lights) { a.mean_lights[3 * (size_t)g + c] = ml[c] * invN; a.mean_local[3 * (size_t)g + c] = mloc[c] * invN; a.mean_global[3 * (size_t)g + c] = mg[c] * invN; }
            }
            if (a.mean_vis) a.mean_vis[g] = mv * invN;
        }
    }
}

// ---- environment-map gradient (grid_sample backward: a scatter-add of P*N*4 taps) ----------------
// There is no native fp32 shared-memory atomic add on sm_100: `atomicAdd(float*)` on shared memory
// compiles to an LDS / FADD / ATOMS.CAST.SPIN loop that occupies the LSU for ~64 cycles per warp
// instruction, and 12 of them per sample bounded the whole backward kernel.  ENV_TAG mode gives
// every warp a PRIVATE float4 {r,g,b,-} copy of the texture gradient and resolves the (rare)
// intra-warp collisions with a one-byte tag per texel: every pending lane stores its lane id, the
// id that survives owns the texel this round and does a plain 128-bit read-modify-write, the losers
// retry.  The copies are summed and flushed once per CTA.
enum { ENV_GLOBAL = 0, ENV_CAS = 1, ENV_TAG = 2 };

__device__ __forceinline__ void warp_private_add(float4* wg, volatile unsigned char* tag, int texel, float x, float y, float z, int lane) {
    bool pending = texel >= 0;
    while (true) {                                   // warp-uniform trip count
        if (pending) tag[texel] = (unsigned char)lane;
        __syncwarp();
        if (pending && tag[texel] == (unsigned char)lane) {
            float4 acc = wg[texel];
            acc.x += x; acc.y += y; acc.z += z;
            wg[texel] = acc;
            pending = false;
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, pending)) break;
    }
}

// resident CTAs the register allocation is pinned to: variant 1 -> 3 (<= 168 registers); variant 2 ->
// 3 with the warp-private env copies (67 KB of shared memory per CTA), 5 otherwise (<= 102 registers)
template <int ENV_MODE, int G, int VAR>
__global__ void __launch_bounds__(SHADE_THREADS, VAR == 0 ? 1 : VAR == 1 ? 3 : (ENV_MODE == 2 ? 3 : 5)) shade_bwd_kernel(const ShadeArgs a) {
    extern __shared__ __align__(16) float s_bwd[];    // [incident slots (VAR >= 1)][gradient accumulators (VAR 2)][env ...]
    constexpr bool SMEM_ENV = ENV_MODE != ENV_GLOBAL;
    constexpr int GPW = 32 / G;
    constexpr int INC_FLOATS = VAR >= 1 ? SHADE_WARPS * GPW * SHADE_INC_STRIDE : 0;
    constexpr int DINC_FLOATS = VAR == 2 ? 48 * SHADE_THREADS : 0;
    float* s_mem = s_bwd + INC_FLOATS + DINC_FLOATS;
    float* sinc = s_bwd + ((threadIdx.x >> 5) * GPW + (threadIdx.x & 31) / G) * SHADE_INC_STRIDE;
    constexpr int VP = (48 + G - 1) / G * G;         // SH-gradient row padded to a multiple of G
    constexpr int RQ = VP / G;                       // finished components per lane
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & (G - 1), grp = lane / G;
    const int nt = a.He * a.We, ne = nt * 3;
    // ENV_CAS: [env ne][grad ne]; ENV_TAG: [env ne (padded to 4)][per warp: float4 grad[nt]][per warp: tag[nt] padded to 16]
    float* s_env = s_mem;
    float* s_genv = s_mem + ne;
    const int ne4 = (ne + 3) & ~3, ntp = (nt + 15) & ~15;
    float4* wgrad_all = reinterpret_cast<float4*>(s_mem + ne4);
    float4* wgrad = wgrad_all + (size_t)warp * nt;
    volatile unsigned char* wtag = reinterpret_cast<unsigned char*>(wgrad_all + (size_t)SHADE_WARPS * nt) + (size_t)warp * ntp;
    if (SMEM_ENV) {
        for (int i = threadIdx.x; i < ne; i += SHADE_THREADS) s_env[i] = a.env[i];
        if (ENV_MODE == ENV_CAS) for (int i = threadIdx.x; i < ne; i += SHADE_THREADS) s_genv[i] = 0.f;
        if (ENV_MODE == ENV_TAG) for (int i = threadIdx.x; i < SHADE_WARPS * nt; i += SHADE_THREADS) wgrad_all[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }
    const float* env = SMEM_ENV ? s_env : a.env;
    float* genv = ENV_MODE == ENV_CAS ? s_genv : a.d_env;
    const float invN = 1.0f / (float)a.N;
    const int units = (a.P + GPW - 1) / GPW;
    for (int u = blockIdx.x * SHADE_WARPS + warp; u < units; u += gridDim.x * SHADE_WARPS) {
        const int g_raw = u * GPW + grp;
        const bool valid = g_raw < a.P;
        const int g = valid ? g_raw : a.P - 1;
        GaussianOperands o;
        load_operands<VAR == 0>(a, g, o);
        if (VAR >= 1) stage_incidents<G>(a, g, sinc, sub);
        float gp[3], gs[3], gd[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            gp[c] = a.g_pbr[3 * (size_t)g + c] * invN;
            gs[c] = a.g_specular ? a.g_specular[3 * (size_t)g + c] * invN : 0.f;
            gd[c] = a.g_diffuse ? a.g_diffuse[3 * (size_t)g + c] * invN : 0.f;
        }
        float dbase[3] = {0, 0, 0}, drough = 0.f, dV[3] = {0, 0, 0};
        float dinc[VP];
#pragma unroll
        for (int q = 0; q < VP; ++q) dinc[q] = 0.f;
        if (VAR == 2) {
            float4* acc4 = reinterpret_cast<float4*>(s_bwd + INC_FLOATS) + threadIdx.x;
#pragma unroll
            for (int q = 0; q < 12; ++q) acc4[q * SHADE_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const size_t row = (size_t)g * a.N;
        for (int i0 = 0; i0 < a.N; i0 += G) {          // warp-uniform trip count (collectives inside in ENV_TAG mode)
            const bool act = valid && (i0 + sub) < a.N;
            const int i = (i0 + sub) < a.N ? (i0 + sub) : a.N - 1;
            const float dx = a.dirs[3 * (row + i)], dy = a.dirs[3 * (row + i) + 1], dz = a.dirs[3 * (row + i) + 2];
            const float vis = a.visibility[row + i], area = a.areas[row + i];
            SampleEval e;
            eval_sample<SMEM_ENV, (VAR >= 1), true>(a, o, env, sinc, dx, dy, dz, vis, e);
            float dfs = 0.f;
            float dG[3], dLg[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float loc = fmaxf(e.local_raw[c], 0.f);
                const float T = (loc + e.glob[c]) * area * e.ndi;
                const float fd = o.base[c] * INV_PI_F;
                const float dT = act ? gp[c] * (fd + e.fs) + gs[c] * e.fs + gd[c] : 0.f;
                if (act) { dbase[c] += gp[c] * T * INV_PI_F; dfs += (gp[c] + gs[c]) * T; }
                const float dL = dT * area * e.ndi;
                dLg[c] = (act && e.local_raw[c] >= 0.f) ? dL : 0.f;  // clamp_min(0) passes the gradient at equality
                if (VAR != 2 && dLg[c] != 0.f) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) dinc[3 * k + c] = fmaf(dL, e.w[k], dinc[3 * k + c]);
                }
                dG[c] = act ? dL * vis : 0.f;                        // -> env texels (grid_sample backward)
            }
            if (VAR == 2) {
                // dinc[3k + c] += dL[c] * w[k]: the lane's 48 accumulators as 12 float4 in shared memory ([12][threads]
                // 16-byte slots: conflict-free 128-bit accesses) — 12 LDS.128 + 48 FFMA + 12 STS.128 per sample
                float4* acc4 = reinterpret_cast<float4*>(s_bwd + INC_FLOATS) + threadIdx.x;
#pragma unroll
                for (int q = 0; q < 12; ++q) {
                    float4 v = acc4[q * SHADE_THREADS];
                    v.x = fmaf(dLg[(4 * q) % 3], e.w[(4 * q) / 3], v.x);
                    v.y = fmaf(dLg[(4 * q + 1) % 3], e.w[(4 * q + 1) / 3], v.y);
                    v.z = fmaf(dLg[(4 * q + 2) % 3], e.w[(4 * q + 2) / 3], v.z);
                    v.w = fmaf(dLg[(4 * q + 3) % 3], e.w[(4 * q + 3) / 3], v.w);
                    acc4[q * SHADE_THREADS] = v;
                }
            }
            if (ENV_MODE == ENV_TAG) {
                const bool any = dG[0] != 0.f || dG[1] != 0.f || dG[2] != 0.f;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int texel = (any && e.taps.idx[t] >= 0) ? e.taps.idx[t] / 3 : -1;
                    warp_private_add(wgrad, wtag, texel, dG[0] * e.taps.w[t], dG[1] * e.taps.w[t], dG[2] * e.taps.w[t], lane);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (dG[c] != 0.f) {
#pragma unroll
                        for (int t = 0; t < 4; ++t) if (e.taps.idx[t] >= 0) atomicAdd(genv + e.taps.idx[t] + c, dG[c] * e.taps.w[t]);
                    }
            }
            if (act) {
                // ---- GGX backward --------------------------------------------------------------
                const float inom = fast_rcp(e.nom);
                const float dfrac = dfs * inom;
                const float dnom = -dfs * (e.frac0 * o.a2) * inom * inom;
                const float dnom_r = (e.nom_r >= 1e-6f && e.nom_r <= 4.0f * PI_F) ? dnom : 0.f;
                const float dnom0 = dnom_r * 8.0f * PI_F * e.nom0 * e.nom1 * e.nom2;
                const float dnom1 = dnom_r * 4.0f * PI_F * e.nom0 * e.nom0 * e.nom2;
                const float dnom2 = dnom_r * 4.0f * PI_F * e.nom0 * e.nom0 * e.nom1;
                const float dfrac0 = dfrac * o.a2;
                const float da2 = dfrac * e.frac0 + dnom0 * e.NoH * e.NoH;
                const float dk = dnom1 * (1.0f - e.NoV) + dnom2 * (1.0f - e.NoL);
                const float r = o.rough;
                drough += da2 * 4.0f * r * r * r + dk * (2.0f * r + 2.0f) / 8.0f;
                const float dNoH = (e.NoH_r >= 1e-6f && e.NoH_r <= 1.f) ? dnom0 * 2.0f * e.NoH * (o.a2 - 1.0f) : 0.f;
                const float dNoV = (e.NoV_r >= 1e-6f && e.NoV_r <= 1.f) ? dnom1 * (1.0f - o.k) : 0.f;
                const float dVoH = (e.VoH_r >= 1e-6f && e.VoH_r <= 1.f)
                                       ? dfrac0 * (1.0f - 0.04f) * e.p2 * 0.693147180559945f * (2.0f * -5.55473f * e.VoH - 6.98316f) : 0.f;
                float dH[3], hdot = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) { dH[c] = dNoH * o.Ns[c] + dVoH * o.V[c]; hdot += e.H[c] * dH[c]; }
                const float ihn = fast_rcp(e.hn);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float dHraw = (dH[c] - e.H[c] * hdot) * ihn;
                    dV[c] += dNoV * o.Ns[c] + dVoH * e.H[c] + 0.5f * dHraw;
                }
            }
        }
        __syncwarp();
        // ---- group reduction and per-Gaussian outputs ---------------------------------------------
#pragma unroll
        for (int c = 0; c < 3; ++c) { dbase[c] = group_sum<G>(dbase[c]); dV[c] = group_sum<G>(dV[c]); }
        drough = group_sum<G>(drough);
        if (VAR == 2) {
            const float4* acc4 = reinterpret_cast<const float4*>(s_bwd + INC_FLOATS) + threadIdx.x;
#pragma unroll
            for (int q = 0; q < 12; ++q) { const float4 v = acc4[q * SHADE_THREADS]; dinc[4 * q] = v.x; dinc[4 * q + 1] = v.y; dinc[4 * q + 2] = v.z; dinc[4 * q + 3] = v.w; }
        }
        group_reduce_scatter<G, VP>(dinc, sub);        // lane `sub` now owns components [sub*RQ, sub*RQ + RQ)
        if (valid) {
            if (sub == 0) {
                const float vd = o.V[0] * dV[0] + o.V[1] * dV[1] + o.V[2] * dV[2];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    a.d_base[3 * (size_t)g + c] = dbase[c];
                    a.d_view[3 * (size_t)g + c] = (dV[c] - o.V[c] * vd) / o.vnorm;        // through V = normalize(viewdirs)
                }
                a.d_rough[g] = drough;
            }
            float* dp = a.d_incidents + 48 * (size_t)g + sub * RQ;    // the group writes the 192 B row contiguously
            if (RQ % 2 == 0) {
#pragma unroll
                for (int q = 0; q < RQ; q += 2)
                    if (sub * RQ + q < 48) *reinterpret_cast<float2*>(dp + q) = make_float2(dinc[q], dinc[q + 1]);
            } else {
#pragma unroll
                for (int q = 0; q < RQ; ++q)
                    if (sub * RQ + q < 48) dp[q] = dinc[q];
            }
        }
    }
    if (ENV_MODE == ENV_CAS) {
        __syncthreads();
        for (int i = threadIdx.x; i < ne; i += SHADE_THREADS) { const float v = s_genv[i]; if (v != 0.f) atomicAdd(a.d_env + i, v); }
    }
    if (ENV_MODE == ENV_TAG) {
        __syncthreads();
        for (int t = threadIdx.x; t < nt; t += SHADE_THREADS) {
            float4 s = wgrad_all[t];
#pragma unroll
            for (int w = 1; w < SHADE_WARPS; ++w) { const float4 v = wgrad_all[(size_t)w * nt + t]; s.x += v.x; s.y += v.y; s.z += v.z; }
            if (s.x != 0.f) atomicAdd(a.d_env + 3 * t, s.x);
            if (s.y != 0.f) atomicAdd(a.d_env + 3 * t + 1, s.y);
            if (s.z != 0.f) atomicAdd(a.d_env + 3 * t + 2, s.z);
        }
    }
}

// ---- launch ---------------------------------------------------------------------------------------
// tuning knobs (A/B measurements and tests of the non-default paths): initial value from the
// environment, changed at run time through r3dg_tune (include/r3dg_b200.h)
struct ShadeKnob { const char* key; const char* env; int lo, hi, dflt, value; };
static ShadeKnob g_knobs[] = {
    {"shade_group", "R3DG_SHADE_GROUP", 8, 32, 8, -1},            // lanes per Gaussian: 8, 16 or 32
    // defaults = the fastest measured combination (profiles/r01_stage3_shading_variants.jsonl, 300k x 64 backward:
    // variant 0 + tag 1.92 ms, variant 1 + tag 1.42, variant 2 + tag 1.51 (3 CTAs/SM by shared memory),
    // variant 2 + shared atomics 1.20 (5 CTAs/SM); forward variant 0 0.57 ms, variant 1 0.47)
    {"shade_env_mode", "R3DG_SHADE_ENV_MODE", 0, 2, ENV_CAS, -1}, // highest env-gradient mode allowed
    {"shade_fwd_variant", "R3DG_SHADE_FWD_VARIANT", 0, 1, 1, -1}, // 1: incident SH coefficients in shared memory
    {"shade_bwd_variant", "R3DG_SHADE_BWD_VARIANT", 0, 2, 2, -1}, // 1: as forward; 2: + gradient accumulators in shared memory
};
static bool knob_ok(const ShadeKnob& k, int v) {
    if (v < k.lo || v > k.hi) return false;
    return strcmp(k.key, "shade_group") != 0 || v == 8 || v == 16 || v == 32;
}
static int knob(int i) {
    ShadeKnob& k = g_knobs[i];
    if (k.value < 0) {
        const char* e = getenv(k.env);
        const int v = e ? atoi(e) : k.dflt;
        k.value = knob_ok(k, v) ? v : k.dflt;
    }
    return k.value;
}
int shade_tune(const char* key, int value, int* previous) {
    for (int i = 0; i < (int)(sizeof(g_knobs) / sizeof(g_knobs[0])); ++i) {
        if (strcmp(key, g_knobs[i].key)) continue;
        *previous = knob(i);
        if (!knob_ok(g_knobs[i], value)) return R3DG_ERR_BAD_ARG;
        g_knobs[i].value = value;
        return 0;
    }
    return R3DG_ERR_UNSUPPORTED;
}

template <typename K>
static int shade_launch(K kernel, const ShadeArgs& a, int units, int num_sms, size_t smem, cudaStream_t stream) {
    if (smem > 48 * 1024) R3DG_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SHADE_THREADS, smem) != cudaSuccess || per_sm <= 0) per_sm = 2;
    const int want = (units + SHADE_WARPS - 1) / SHADE_WARPS;
    const int cap = num_sms * per_sm;                 // persistent CTAs: one env staging / flush per resident CTA
    const int grid = want < cap ? (want > 0 ? want : 1) : cap;
    kernel<<<grid, SHADE_THREADS, smem, stream>>>(a);
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

static size_t inc_slot_bytes(int G, int var) { return var >= 1 ? (size_t)SHADE_WARPS * (32 / G) * SHADE_INC_STRIDE * sizeof(float) : 0; }

template <bool SMEM>
static int launch_fwd(const ShadeArgs& a, int num_sms, size_t env_smem, cudaStream_t stream) {
    const int G = knob(0);
    const int var = G == 8 ? knob(2) : 0;              // the shared-memory variants are built for the default group width
    const size_t smem = env_smem + inc_slot_bytes(G, var);
    const int units = (a.P + 32 / G - 1) / (32 / G);
    if (G == 8) return var ? shade_launch(shade_fwd_kernel<SMEM, 8, 1>, a, units, num_sms, smem, stream)
                           : shade_launch(shade_fwd_kernel<SMEM, 8, 0>, a, units, num_sms, smem, stream);
    if (G == 16) return shade_launch(shade_fwd_kernel<SMEM, 16, 0>, a, units, num_sms, smem, stream);
    return shade_launch(shade_fwd_kernel<SMEM, 32, 0>, a, units, num_sms, smem, stream);
}

int launch_shade_forward(ShadeArgs a, int num_sms, cudaStream_t stream) {
    if (a.P <= 0) return 0;
    const size_t env_bytes = (size_t)a.He * a.We * 3 * sizeof(float);
    const bool smem = env_bytes <= 40 * 1024;
    a.env_in_smem = smem;
    return smem ? launch_fwd<true>(a, num_sms, env_bytes, stream) : launch_fwd<false>(a, num_sms, 0, stream);
}

template <int MODE>
static int launch_bwd(const ShadeArgs& a, int num_sms, size_t env_smem, cudaStream_t stream) {
    const int G = knob(0);
    const int var = G == 8 ? knob(3) : 0;
    const size_t smem = env_smem + inc_slot_bytes(G, var) + (var == 2 ? (size_t)48 * SHADE_THREADS * sizeof(float) : 0);
    const int units = (a.P + 32 / G - 1) / (32 / G);
    if (G == 8) return var == 2 ? shade_launch(shade_bwd_kernel<MODE, 8, 2>, a, units, num_sms, smem, stream)
                     : var == 1 ? shade_launch(shade_bwd_kernel<MODE, 8, 1>, a, units, num_sms, smem, stream)
                                : shade_launch(shade_bwd_kernel<MODE, 8, 0>, a, units, num_sms, smem, stream);
    if (G == 16) return shade_launch(shade_bwd_kernel<MODE, 16, 0>, a, units, num_sms, smem, stream);
    return shade_launch(shade_bwd_kernel<MODE, 32, 0>, a, units, num_sms, smem, stream);
}

int launch_shade_backward(ShadeArgs a, int num_sms, cudaStream_t stream) {
    const size_t nt = (size_t)a.He * a.We, env_bytes = nt * 3 * sizeof(float);
    R3DG_CUDA_TRY(cudaMemsetAsync(a.d_env, 0, env_bytes, stream));
    if (a.P <= 0) return 0;
    // shared-memory cost of the env-gradient modes (kept within the 48 KB that need no opt-in on their own)
    const size_t tag_bytes = ((nt * 3 + 3) & ~(size_t)3) * sizeof(float) + SHADE_WARPS * (nt * sizeof(float4) + ((nt + 15) & ~(size_t)15));
    const size_t cas_bytes = 2 * env_bytes;
    const int cap = knob(1);
    if (cap >= ENV_TAG && tag_bytes <= 48 * 1024) { a.env_in_smem = 1; return launch_bwd<ENV_TAG>(a, num_sms, tag_bytes, stream); }
    if (cap >= ENV_CAS && cas_bytes <= 40 * 1024) { a.env_in_smem = 1; return launch_bwd<ENV_CAS>(a, num_sms, cas_bytes, stream); }
    a.env_in_smem = 0;
    return launch_bwd<ENV_GLOBAL>(a, num_sms, 0, stream);
}

}  // namespace r3dg

// ---- C ABI --------------------------------------------------------------------------------------
using namespace r3dg;
extern "C" {

static int shade_num_sms() {
    static int n = 0;
    if (n == 0) { int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148; }
    return n;
}

static int fill_args(const r3dg_shade_args* s, ShadeArgs& a) {
    if (!s || s->P < 0 || s->N <= 0 || s->env_h <= 0 || s->env_w <= 0) return R3DG_ERR_BAD_ARG;
    if (s->sh_coeffs != 16) return R3DG_ERR_UNSUPPORTED;          // incidents are degree-3 SH ([P,16,3]) in the reference
    a.P = s->P; a.N = s->N; a.He = s->env_h; a.We = s->env_w; a.env_in_smem = 0;
    a.base_color = s->base_color; a.roughness = s->roughness; a.normals = s->normals; a.viewdirs = s->viewdirs;
    a.incidents = s->incidents; a.env = s->env; a.transform = s->env_transform;
    a.visibility = s->visibility; a.dirs = s->incident_dirs; a.areas = s->incident_areas;
    a.pbr = s->pbr; a.diffuse = s->diffuse_light; a.specular = s->specular;
    a.mean_lights = s->mean_incident_lights; a.mean_local = s->mean_local_lights; a.mean_global = s->mean_global_lights;
    a.mean_vis = s->mean_visibility;
    a.s_lights = s->incident_lights; a.s_local = s->local_incident_lights; a.s_global = s->global_incident_lights;
    a.g_pbr = s->dL_dpbr; a.g_diffuse = s->dL_ddiffuse_light; a.g_specular = s->dL_dspecular;
    a.d_base = s->dL_dbase_color; a.d_rough = s->dL_droughness; a.d_view = s->dL_dviewdirs; a.d_incidents = s->dL_dincidents;
    a.d_env = s->dL_denv;
    return 0;
}

int r3dg_render_equation_forward(const r3dg_shade_args* s, r3dg_stream_t stream) {
    ShadeArgs a;
    int rc = fill_args(s, a);
    if (rc) return rc;
    if (a.P > 0 && (!a.pbr || !a.diffuse || !a.specular)) return R3DG_ERR_BAD_ARG;
    if ((a.mean_lights != nullptr) != (a.mean_local != nullptr) || (a.mean_lights != nullptr) != (a.mean_global != nullptr)) return R3DG_ERR_BAD_ARG;
    if ((a.s_lights != nullptr) != (a.s_local != nullptr) || (a.s_lights != nullptr) != (a.s_global != nullptr)) return R3DG_ERR_BAD_ARG;
    return launch_shade_forward(a, shade_num_sms(), (cudaStream_t)stream);
}

int r3dg_render_equation_backward(const r3dg_shade_args* s, r3dg_stream_t stream) {
    ShadeArgs a;
    int rc = fill_args(s, a);
    if (rc) return rc;
    if (!a.g_pbr || !a.d_base || !a.d_rough || !a.d_view || !a.d_incidents || !a.d_env) return R3DG_ERR_BAD_ARG;
    return launch_shade_backward(a, shade_num_sms(), (cudaStream_t)stream);
}

}  // extern "C"
