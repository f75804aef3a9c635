// Depth-ordered tile binning: builds the per-tile, depth-sorted Gaussian lists
// (binningState.point_list + imgState.ranges) that the reference obtains from
// duplicateWithKeys -> cub::DeviceRadixSort::SortPairs over 64-bit (tile | depth) keys ->
// identifyTileRanges (rasterizer_impl.cu:70-138, 296-333).
//
// The reference sorts all R ~ 8 P (tile, Gaussian) instances by (tile, depth), stable, instances
// emitted in Gaussian-index order.  The same total order is obtained much cheaper:
//   1. stable-sort the P Gaussians by their depth bits (radix_sort.cu; ties keep index order);
//   2. append them, in that order, to the lists of the tiles their rectangle covers.
// Step 2 is a counting sort by tile id that must preserve the sequence order, done with two
// passes over the depth-sorted sequence cut into chunks (one warp per chunk, a private per-tile
// table in shared memory):
//   bin_pass<false> : M[chunk][tile] = number of instances the chunk adds to the tile;
//   bin_scan        : exclusive scan of every column of M (per tile over the chunks) + totals;
//   bin_tile_start  : exclusive scan of the totals -> ranges[tile], R;
//   bin_pass<true>  : every chunk re-walks its Gaussians and writes each instance to
//                     start[tile] + M[chunk][tile] + (rank inside the chunk, in sequence order).
// Equal (tile, depth) keys end up in Gaussian-index order, exactly as the reference's stable sort
// leaves them, so point_list and ranges are bit-identical.
#include "common.cuh"
#include "kernels.h"

namespace r3dg {

// slot -> (Gaussian of the batch, tile) for a batch of 32 Gaussians whose inclusive instance
// counts are staged in shared memory (same scheme as the cooperative key emission it replaces).
struct BatchStage { uint32_t incl[32], xy[32], w[32], g[32]; };

template <bool SCATTER>
__global__ void __launch_bounds__(256) bin_pass_kernel(int P, int T, int gx, int CH, int chunks,
                                                       const GeomHeader* __restrict__ header,
                                                       const uint32_t* __restrict__ vals_a,
                                                       const uint32_t* __restrict__ vals_b,
                                                       const uint2* __restrict__ rects,
                                                       uint32_t* __restrict__ M,
                                                       const uint2* __restrict__ ranges,
                                                       uint32_t* __restrict__ point_list, long long capacity) {
    extern __shared__ __align__(16) uint32_t bin_smem[];
    const int nwb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * nwb + warp;
    if (c >= chunks) return;                              // no CTA-wide barriers below
    uint32_t* tbl = bin_smem + (size_t)warp * T;
    BatchStage& st = reinterpret_cast<BatchStage*>(bin_smem + (size_t)nwb * T)[warp];
    const uint32_t* __restrict__ order = (header->sort_exec & 1u) ? vals_b : vals_a;
    uint32_t* row = M + (size_t)c * T;
    for (int t = lane; t < T; t += 32) tbl[t] = SCATTER ? ranges[t].x + row[t] : 0u;
    __syncwarp();
    const int j0 = c * CH, j1 = min(P, j0 + CH);
    // two-deep software pipeline over the dependent gathers order[j] -> rects[g]
    uint32_t g_cur = 0, g_nxt = 0;
    uint2 r_cur = make_uint2(0u, 0u);
    if (j0 + lane < j1) { g_cur = order[j0 + lane]; r_cur = rects[g_cur]; }
    if (j0 + 32 + lane < j1) g_nxt = order[j0 + 32 + lane];
    for (int jb = j0; jb < j1; jb += 32) {
        const uint32_t g = g_cur;
        const uint2 r = (jb + lane < j1) ? r_cur : make_uint2(0u, 0u);
        // prefetch: rect of the next batch, index of the one after
        g_cur = g_nxt;
        if (jb + 32 + lane < j1) r_cur = rects[g_cur];
        if (jb + 64 + lane < j1) g_nxt = order[jb + 64 + lane];
        const uint32_t w = r.y & 0xffffu, h = r.y >> 16;
        const uint32_t cnt = w * h;
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        st.incl[lane] = incl; st.xy[lane] = r.x; st.w[lane] = w ? w : 1u; st.g[lane] = g;
        __syncwarp();
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t s = base + lane;
            const bool active = s < total;
            int lo = 0, hi = 31;                          // smallest q with incl[q] > s
#pragma unroll
            for (int it = 0; it < 5; ++it) { const int mid = (lo + hi) >> 1; if (st.incl[mid] > s) hi = mid; else lo = mid + 1; }
            const int q = lo;
            const uint32_t excl = q == 0 ? 0u : st.incl[q - 1];
            const uint32_t tt = active ? s - excl : 0u, ww = st.w[q], xy = st.xy[q];
            const uint32_t ty = (xy >> 16) + tt / ww, tx = (xy & 0xffffu) + tt % ww;
            const uint32_t tile = ty * (uint32_t)gx + tx;
            if (!SCATTER) {
                if (active) atomicAdd(&tbl[tile], 1u);
            } else {
                // lanes are in sequence order: equal tiles are ranked by lane
                const uint32_t key = active ? tile : 0xffffffffu - (uint32_t)lane;
                const uint32_t peers = __match_any_sync(0xffffffffu, key);
                const int leader = __ffs(peers) - 1;
                const uint32_t below = __popc(peers & ((1u << lane) - 1u));
                uint32_t old = 0;
                if (active && lane == leader) { old = tbl[tile]; tbl[tile] = old + __popc(peers); }
                old = __shfl_sync(0xffffffffu, old, leader);
                if (active) {
                    const long long pos = (long long)old + below;
                    if (pos < capacity) point_list[pos] = st.g[q];
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
    if (!SCATTER) {
        __syncwarp();
        for (int t = lane; t < T; t += 32) row[t] = tbl[t];
    }
}

// Column scan of M: block = 32 tiles x 32 chunk groups.
__global__ void __launch_bounds__(1024) bin_scan_kernel(int T, int chunks, uint32_t* __restrict__ M,
                                                        uint32_t* __restrict__ tile_total) {
    __shared__ uint32_t s[32][33];
    const int tx = threadIdx.x, gy = threadIdx.y;
    const int t = blockIdx.x * 32 + tx;
    const int G = (chunks + 31) / 32;
    const int c0 = min(chunks, gy * G), c1 = min(chunks, c0 + G);
    uint32_t sum = 0;
    if (t < T)
        for (int c = c0; c < c1; ++c) sum += M[(size_t)c * T + t];
    s[gy][tx] = sum;
    __syncthreads();
    if (gy == 0) {
        uint32_t acc = 0;
#pragma unroll
        for (int g = 0; g < 32; ++g) { const uint32_t v = s[g][tx]; s[g][tx] = acc; acc += v; }
        if (t < T) tile_total[t] = acc;
    }
    __syncthreads();
    uint32_t run = s[gy][tx];
    if (t < T)
        for (int c = c0; c < c1; ++c) { const uint32_t v = M[(size_t)c * T + t]; M[(size_t)c * T + t] = run; run += v; }
}

// Exclusive scan of the tile totals -> ranges (empty tiles stay (0,0) like the reference's
// zero-initialised imgState.ranges), R -> header.  One CTA.
__global__ void __launch_bounds__(1024) bin_tile_start_kernel(int T, const uint32_t* __restrict__ tile_total,
                                                              uint2* __restrict__ ranges, GeomHeader* header,
                                                              long long capacity) {
    __shared__ uint32_t s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (T + 1023) / 1024;
    const int t0 = min(T, tid * per), t1 = min(T, t0 + per);
    uint32_t sum = 0;
    for (int t = t0; t < t1; ++t) sum += tile_total[t];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t wv = s_warp[lane];
        uint32_t wi = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
        s_warp[lane] = wi - wv;
        if (lane == 31) header->num_rendered = wi;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + inc - sum;
    for (int t = t0; t < t1; ++t) {
        const uint32_t n = tile_total[t];
        // on overflow (R > capacity: the host grows the buffer and re-runs) keep every range inside the buffer
        ranges[t] = n ? make_uint2((uint32_t)min((long long)run, capacity), (uint32_t)min((long long)run + n, capacity))
                      : make_uint2(0u, 0u);
        run += n;
    }
}

// Debug only (r3dg_raster_debug_copy id 10): the reference's sorted 64-bit keys, rebuilt from the
// lists: (tile << 32) | depth bits of the listed Gaussian.
__global__ void rebuild_keys_kernel(int T, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                    const float* __restrict__ rec, int recf, long long limit, uint64_t* __restrict__ keys) {
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        const uint2 r = ranges[t];
        for (long long i = (long long)r.x + threadIdx.x; i < (long long)r.y && i < limit; i += blockDim.x)
            keys[i] = ((uint64_t)(uint32_t)t << 32) | (uint64_t)__float_as_uint(rec[(size_t)point_list[i] * recf + 6]);
    }
}

int launch_rebuild_keys(int T, const void* ranges, const uint32_t* point_list, const float* rec, int recf,
                        long long limit, uint64_t* keys, cudaStream_t stream) {
    rebuild_keys_kernel<<<1024, 128, 0, stream>>>(T, (const uint2*)ranges, point_list, rec, recf, limit, keys);
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

// Warps per CTA for the binning passes: as many private per-tile tables as fit.
static int bin_warps_per_block(int T) {
    const size_t per_warp = (size_t)T * 4 + sizeof(BatchStage);
    const size_t budget = 100 * 1024;                     // two CTAs per SM when it fits
    int nwb = (int)(budget / per_warp);
    if (nwb >= 8) return 8;
    if (nwb >= 1) return nwb;
    return per_warp <= 227 * 1024 - 1024 ? 1 : 0;
}

int launch_binning(int P, int W, int H, char* geom, const GeomLayout& gl, char* img, const ImgLayout& il,
                   char* bin, const BinLayout& bl, int num_sms, cudaStream_t stream, stage_mark_fn mark) {
    const int gx = (W + R3DG_TILE - 1) / R3DG_TILE, gy = (H + R3DG_TILE - 1) / R3DG_TILE;
    const int T = gx * gy;
    if (gx > 65535 || gy > 65535) return R3DG_ERR_UNSUPPORTED;
    const int nwb = bin_warps_per_block(T);
    if (nwb == 0) return R3DG_ERR_UNSUPPORTED;             // > ~58k tiles: per-warp table exceeds shared memory
    GeomHeader* header = (GeomHeader*)(geom + gl.header);
    const SortLayout sl(P);
    char* sbuf = geom + gl.sort;
    int rc = launch_sort(header, sbuf, sl, P, num_sms, stream);
    if (rc != 0) return rc;
    mark(2, stream);
    const int CH = bin_chunk_len(P, (size_t)T);
    const int chunks = (P + CH - 1) / CH;
    const size_t smem = (size_t)nwb * ((size_t)T * 4 + sizeof(BatchStage));
    static size_t attr_smem[2] = {0, 0};
    if (smem > attr_smem[0]) {
        R3DG_CUDA_TRY(cudaFuncSetAttribute(bin_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        R3DG_CUDA_TRY(cudaFuncSetAttribute(bin_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem[0] = smem;
    }
    const uint32_t* va = (const uint32_t*)(sbuf + sl.vals_a);
    const uint32_t* vb = (const uint32_t*)(sbuf + sl.vals_b);
    const uint2* rects = (const uint2*)(geom + gl.rects);
    uint32_t* M = (uint32_t*)(img + il.bin_matrix);
    uint2* ranges = (uint2*)(img + il.ranges);
    uint32_t* tile_total = (uint32_t*)(img + il.tile_total);
    uint32_t* point_list = (uint32_t*)(bin + bl.point_list);
    const int grid = (chunks + nwb - 1) / nwb;
    bin_pass_kernel<false><<<grid, nwb * 32, smem, stream>>>(P, T, gx, CH, chunks, header, va, vb, rects, M, ranges, point_list, bl.capacity);
    mark(3, stream);
    bin_scan_kernel<<<(T + 31) / 32, dim3(32, 32), 0, stream>>>(T, chunks, M, tile_total);
    bin_tile_start_kernel<<<1, 1024, 0, stream>>>(T, tile_total, ranges, header, bl.capacity);
    mark(4, stream);
    bin_pass_kernel<true><<<grid, nwb * 32, smem, stream>>>(P, T, gx, CH, chunks, header, va, vb, rects, M, ranges, point_list, bl.capacity);
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace r3dg
