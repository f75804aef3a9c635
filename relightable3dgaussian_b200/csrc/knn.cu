// `simple_knn._C.distCUDA2` (reference submodules/simple-knn/simple_knn.cu:185-221, spatial.cu:15-26):
// mean squared distance of every point to its 3 nearest neighbours.  Init-time only (called once
// from GaussianModel.create_from_pcd, scene/gaussian_model.py:427) but the symbol is imported at
// module import of scene/gaussian_model.py:13, so a drop-in needs it (SURVEY.md §8f "next #1").
//
// Same algorithm class as the reference (Morton order, boxes of 1024 consecutive points, exact
// pruning by box distance), so the result is the exact 3-NN mean; built from this library's own
// radix sort, with the points gathered into Morton order once so the box scans are coalesced.
#include <cfloat>
#include "common.cuh"
#include "kernels.h"

namespace r3dg {

#define KNN_BOX 1024

struct KnnTmp {
    size_t header, bounds, sorted_pts, boxes, bin, total;
    SortLayout bl;
    __host__ KnnTmp(int P) : bl(P < 1 ? 1 : P) {
        size_t off = 0;
        header = off;     off = align_up(off + sizeof(GeomHeader), 256);
        bounds = off;     off = align_up(off + 6 * 4, 256);
        sorted_pts = off; off = align_up(off + (size_t)P * 16, 256);
        boxes = off;      off = align_up(off + ((size_t)P / KNN_BOX + 1) * 32, 256);
        bin = off;        off = align_up(off + bl.total, 256);
        total = off;
    }
};

__device__ __forceinline__ int f2ord_k(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f_k(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

__global__ void knn_init_kernel(GeomHeader* h, int* bounds, int P) {
    if (threadIdx.x == 0) {
        h->num_rendered = (uint32_t)P; h->depth_or = 0x3fffffffu; h->depth_nor = 0x3fffffffu;   // sort all 30 Morton bits
        for (int k = 0; k < 3; ++k) { bounds[k] = f2ord_k(0.f); bounds[3 + k] = f2ord_k(0.f); }   // reduction init {0,0,0}, simple_knn.cu:191
    }
}

__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float* __restrict__ pts, int* bounds) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float v = pts[3 * (size_t)i + k]; lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[k], f2ord_k(lo[k])); atomicMax(&bounds[3 + k], f2ord_k(hi[k])); }
    }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x) {      // prepMorton, simple_knn.cu:44-52
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256) knn_morton_kernel(int P, const float* __restrict__ pts, const int* __restrict__ bounds,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float lo = ord2f_k(bounds[k]), hi = ord2f_k(bounds[3 + k]);
        const float t = ((pts[3 * (size_t)i + k] - lo) / (hi - lo)) * 1023.0f;
        c[k] = spread10((uint32_t)fminf(fmaxf(t, 0.f), 1023.f));
    }
    keys[i] = c[0] | (c[1] << 1) | (c[2] << 2);
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) knn_gather_kernel(int P, const GeomHeader* __restrict__ h, const uint32_t* __restrict__ va,
                                                         const uint32_t* __restrict__ vb, const float* __restrict__ pts,
                                                         float4* __restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t idx = (h->sort_exec & 1u) ? vb[i] : va[i];
    sorted[i] = make_float4(pts[3 * (size_t)idx], pts[3 * (size_t)idx + 1], pts[3 * (size_t)idx + 2], __uint_as_float(idx));
}

__global__ void __launch_bounds__(KNN_BOX) knn_boxes_kernel(int P, const float4* __restrict__ sorted, float* __restrict__ boxes) {
    __shared__ float red[6][KNN_BOX / 32];
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) { const float4 p = sorted[i]; lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
        if ((threadIdx.x & 31) == 0) { red[k][threadIdx.x >> 5] = lo[k]; red[3 + k][threadIdx.x >> 5] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[threadIdx.x][0];
        for (int w = 1; w < KNN_BOX / 32; ++w) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][w]) : fmaxf(v, red[threadIdx.x][w]);
        boxes[8 * (size_t)blockIdx.x + threadIdx.x] = v;
    }
}

__device__ __forceinline__ void update3(float dx, float dy, float dz, float* best) {      // updateKBest<3>, simple_knn.cu:131-145
    float dist = dx * dx + dy * dy + dz * dz;
#pragma unroll
    for (int j = 0; j < 3; ++j) if (best[j] > dist) { const float t = best[j]; best[j] = dist; dist = t; }
}

__global__ void __launch_bounds__(256) knn_search_kernel(int P, const float4* __restrict__ sorted, const float* __restrict__ boxes,
                                                         float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float4 me = sorted[idx];
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int i = max(0, idx - 3); i <= min(P - 1, idx + 3); ++i) {
        if (i == idx) continue;
        const float4 q = sorted[i];
        update3(q.x - me.x, q.y - me.y, q.z - me.z, best);
    }
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;
    const int nboxes = (P + KNN_BOX - 1) / KNN_BOX;
    for (int b = 0; b < nboxes; ++b) {
        const float* bx = boxes + 8 * (size_t)b;
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {      // distBoxPoint, simple_knn.cu:118-128
            const float p = k == 0 ? me.x : (k == 1 ? me.y : me.z);
            if (p < bx[k] || p > bx[3 + k]) { const float t = fminf(fabsf(p - bx[k]), fabsf(p - bx[3 + k])); d += t * t; }
        }
        if (d > reject || d > best[2]) continue;
        const int e = min(P, (b + 1) * KNN_BOX);
        for (int i = b * KNN_BOX; i < e; ++i) {
            if (i == idx) continue;
            const float4 q = sorted[i];
            update3(q.x - me.x, q.y - me.y, q.z - me.z, best);
        }
    }
    out[__float_as_uint(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

size_t knn_tmp_bytes(int P) { return KnnTmp(P).total; }

int launch_knn(int P, const float* points, float* out, void* tmp_, size_t tmp_bytes, int num_sms, cudaStream_t stream) {
    if (P <= 0) return 0;
    const KnnTmp t(P);
    if (tmp_bytes < t.total) return R3DG_ERR_BAD_ARG;
    char* tmp = (char*)tmp_;
    GeomHeader* h = (GeomHeader*)(tmp + t.header);
    int* bounds = (int*)(tmp + t.bounds);
    float4* sorted = (float4*)(tmp + t.sorted_pts);
    float* boxes = (float*)(tmp + t.boxes);
    char* bin = tmp + t.bin;
    R3DG_CUDA_TRY(cudaMemsetAsync(h, 0, sizeof(GeomHeader), stream));
    knn_init_kernel<<<1, 32, 0, stream>>>(h, bounds, P);
    knn_bounds_kernel<<<num_sms * 4, 256, 0, stream>>>(P, points, bounds);
    knn_morton_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, points, bounds, (uint32_t*)(bin + t.bl.keys_a), (uint32_t*)(bin + t.bl.vals_a));
    int rc = launch_sort(h, bin, t.bl, P, num_sms, stream);
    if (rc != 0) return rc;
    knn_gather_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, h, (const uint32_t*)(bin + t.bl.vals_a), (const uint32_t*)(bin + t.bl.vals_b), points, sorted);
    knn_boxes_kernel<<<(P + KNN_BOX - 1) / KNN_BOX, KNN_BOX, 0, stream>>>(P, sorted, boxes);
    knn_search_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, sorted, boxes, out);
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace r3dg
