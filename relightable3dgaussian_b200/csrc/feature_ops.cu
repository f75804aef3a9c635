// §8(f)2: the host-side epilogue the reference wraps around the rasterizer
// (gaussian_renderer/neilf.py:135-137, render.py:106-108):
//     rendered_feature = rendered_feature / rendered_opacity.clamp_min(1e-5) * (num_contrib > 0)
// In PyTorch this is 3 elementwise kernels forward over [S,H,W] (plus ~6 in autograd's backward, which
// also keeps two [S,H,W] intermediates alive).  Here: ONE pass forward, ONE pass backward, both
// HBM-bound (forward reads S+2 planes and writes S; backward reads 2S+2 and writes S+1), optional —
// the reference's unmodified render functions keep using PyTorch; `rasterizer.unpremultiply` opts in.
#include "common.cuh"
#include "kernels.h"

namespace r3dg {

__global__ void __launch_bounds__(256) unpremultiply_fwd_kernel(int S, long long HW, const float* __restrict__ feature,
                                                                const float* __restrict__ opacity,
                                                                const int32_t* __restrict__ n_contrib, float* __restrict__ out) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
        const float den = fmaxf(opacity[p], 1e-5f);
        const float m = n_contrib[p] > 0 ? 1.0f : 0.0f;
        for (int c = 0; c < S; ++c) out[c * HW + p] = mul_(div_(feature[c * HW + p], den), m);   // torch op order: (f / den) * mask
    }
}

// d_feature = g * m / den;  d_opacity = -(sum_c g_c * m * f_c) / den^2 where opacity >= 1e-5 (clamp_min passes its gradient there)
__global__ void __launch_bounds__(256) unpremultiply_bwd_kernel(int S, long long HW, const float* __restrict__ feature,
                                                                const float* __restrict__ opacity,
                                                                const int32_t* __restrict__ n_contrib, const float* __restrict__ g,
                                                                float* __restrict__ d_feature, float* __restrict__ d_opacity) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
        const float o = opacity[p];
        const float den = fmaxf(o, 1e-5f);
        const float m = n_contrib[p] > 0 ? 1.0f : 0.0f;
        const float inv = 1.0f / den;
        float acc = 0.0f;
        for (int c = 0; c < S; ++c) {
            const float gm = g[c * HW + p] * m;
            d_feature[c * HW + p] = gm * inv;
            acc += gm * feature[c * HW + p];
        }
        d_opacity[p] = o >= 1e-5f ? -acc * inv * inv : 0.0f;
    }
}

static int grid_for(long long n, int num_sms) {
    const long long want = (n + 255) / 256;
    return (int)(want < (long long)num_sms * 16 ? want : (long long)num_sms * 16);
}

int launch_unpremultiply_forward(int S, long long HW, const float* feature, const float* opacity, const int32_t* n_contrib,
                                 float* out, int num_sms, cudaStream_t stream) {
    if (S <= 0 || HW <= 0) return 0;
    unpremultiply_fwd_kernel<<<grid_for(HW, num_sms), 256, 0, stream>>>(S, HW, feature, opacity, n_contrib, out);
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_unpremultiply_backward(int S, long long HW, const float* feature, const float* opacity, const int32_t* n_contrib,
                                  const float* g, float* d_feature, float* d_opacity, int num_sms, cudaStream_t stream) {
    if (HW <= 0) return 0;
    unpremultiply_bwd_kernel<<<grid_for(HW, num_sms), 256, 0, stream>>>(S, HW, feature, opacity, n_contrib, g, d_feature, d_opacity);
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace r3dg
