// Stable LSD radix sort of (u64 key, u32 value) pairs — replaces cub::DeviceRadixSort::SortPairs
// at reference rasterizer_impl.cu:313-318 (K5).  Hand-written one-sweep design for sm_100a:
//
//   1. sort_histogram_kernel : one read of the keys builds the digit histograms of ALL passes
//      (and clears the look-back descriptors this frame will use);
//   2. sort_scan_kernel      : exclusive scan of each 256-bin histogram -> global digit bases;
//   3. sort_onesweep_kernel  : per pass ONE read + ONE write of the pairs.  A persistent grid
//      pulls 3072-key tiles from an atomic ticket; per tile: warp-synchronous match-any ranking
//      (stable), chained decoupled look-back per digit for the tile's global offset, reorder
//      through shared memory so that global writes are digit-contiguous runs.
//
// The number of keys R is read from the geometry header on the device: no host sync, grids are
// sized from the SM count.  Stability + emission order (Gaussian index ascending) reproduces
// the reference's tie order for equal (tile, depth) keys bit-exactly.
#include "common.cuh"
#include "kernels.h"

namespace r3dg {

#define FLAG_AGG 0x40000000u
#define FLAG_INC 0x80000000u
#define VAL_MASK 0x3fffffffu

__global__ void __launch_bounds__(256) sort_histogram_kernel(const GeomHeader* __restrict__ header,
                                                             long long capacity, int passes,
                                                             const uint64_t* __restrict__ keys,
                                                             uint32_t* __restrict__ hist,
                                                             uint32_t* __restrict__ lookback,
                                                             long long max_tiles) {
    __shared__ uint32_t sh[R3DG_SORT_MAX_PASSES * 256];
    const long long R = min((long long)header->num_rendered, capacity);
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < R; i += stride) {
        const uint64_t k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * 256 + (uint32_t)((k >> (8 * p)) & 0xff)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
    // clear the look-back descriptors of the tiles this frame uses
    const long long ntiles = (R + R3DG_SORT_TILE - 1) / R3DG_SORT_TILE;
    for (int p = 0; p < passes; ++p) {
        uint32_t* lb = lookback + (size_t)p * max_tiles * 256;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ntiles * 256; i += stride) lb[i] = 0;
    }
}

__global__ void __launch_bounds__(256) sort_scan_kernel(int passes, uint32_t* __restrict__ hist) {
    // one block; thread d owns bin d of every pass; exclusive scan through shared memory
    __shared__ uint32_t s[256];
    for (int p = 0; p < passes; ++p) {
        const uint32_t v = hist[p * 256 + threadIdx.x];
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            uint32_t t = threadIdx.x >= d ? s[threadIdx.x - d] : 0u;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        hist[p * 256 + threadIdx.x] = s[threadIdx.x] - v;
        __syncthreads();
    }
}

struct SortSmem {
    uint64_t keys[R3DG_SORT_TILE];
    uint32_t vals[R3DG_SORT_TILE];
    uint32_t warp_hist[R3DG_SORT_THREADS / 32][256];
    uint32_t local_off[256];     // exclusive prefix of digit counts inside the tile
    uint32_t global_off[256];    // global start of this tile's run of digit d
    uint32_t tile;
};

__global__ void __launch_bounds__(R3DG_SORT_THREADS) sort_onesweep_kernel(
    GeomHeader* header, long long capacity, int slot, int passes, const uint64_t* __restrict__ keys_in,
    const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
    const uint32_t* __restrict__ digit_base, volatile uint32_t* lookback) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem& sm = *reinterpret_cast<SortSmem*>(smem_raw);
    const long long R = min((long long)header->num_rendered, capacity);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // slot = index among the EXECUTED passes; the digit it sorts on is decided on the device
    const uint32_t depth_diff = header->depth_or & header->depth_nor;
    const int pass = sort_digit_of_slot(depth_diff, passes, slot);
    if (pass < 0) return;                              // fewer digits needed than slots launched
    if (blockIdx.x == 0 && tid == 0 && slot == 0) header->sort_exec = (uint32_t)sort_num_exec(depth_diff, passes);
    const int shift = 8 * pass;
    const long long ntiles = (R + R3DG_SORT_TILE - 1) / R3DG_SORT_TILE;
    constexpr int NW = R3DG_SORT_THREADS / 32;
    constexpr int PER_WARP = R3DG_SORT_TILE / NW;        // 384 consecutive keys per warp

    while (true) {
        __syncthreads();
        if (tid == 0) sm.tile = atomicAdd(&header->sort_ticket[slot], 1u);
        for (int i = tid; i < NW * 256; i += R3DG_SORT_THREADS) (&sm.warp_hist[0][0])[i] = 0;
        __syncthreads();
        const long long tile = sm.tile;
        if (tile >= ntiles) break;
        const long long tile_base = tile * R3DG_SORT_TILE;
        const int count = (int)min((long long)R3DG_SORT_TILE, R - tile_base);

        // ---- load + stable ranking within the warp's 384-key slice ---------------------------
        uint64_t k[R3DG_SORT_ITEMS];
        uint32_t v[R3DG_SORT_ITEMS];
        uint32_t rank[R3DG_SORT_ITEMS];
#pragma unroll
        for (int i = 0; i < R3DG_SORT_ITEMS; ++i) {
            const int local = warp * PER_WARP + i * 32 + lane;
            const bool valid = local < count;
            k[i] = valid ? keys_in[tile_base + local] : ~0ull;
            v[i] = valid ? vals_in[tile_base + local] : 0u;
        }
#pragma unroll
        for (int i = 0; i < R3DG_SORT_ITEMS; ++i) {
            const uint32_t d = (uint32_t)(k[i] >> shift) & 0xffu;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            const uint32_t below = __popc(peers & ((1u << lane) - 1u));
            uint32_t old = 0;
            if (lane == leader) { old = sm.warp_hist[warp][d]; sm.warp_hist[warp][d] = old + __popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[i] = old + below;
            __syncwarp();
        }
        __syncthreads();

        // ---- per digit: warp prefixes, tile count, decoupled look-back ------------------------
        {
            const int d = tid;                            // R3DG_SORT_THREADS == 256 digits
            uint32_t acc = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { const uint32_t t = sm.warp_hist[w][d]; sm.warp_hist[w][d] = acc; acc += t; }
            // padding keys (~0) inflate digit 255 of the last tile: remove them from the count
            uint32_t cnt = acc;
            if (d == 255) cnt -= (uint32_t)(R3DG_SORT_TILE - count);
            uint32_t* lb = const_cast<uint32_t*>(lookback) + (size_t)tile * 256;
            uint32_t prefix = 0;
            if (tile == 0) {
                st_relaxed_gpu(lb + d, cnt | FLAG_INC);
            } else {
                st_relaxed_gpu(lb + d, cnt | FLAG_AGG);
                // Decoupled look-back.  The nearest predecessor is usually still ranking: spin on it
                // with single loads; everything further back was published long ago, so those
                // descriptors are fetched LB_BATCH at a time instead of as a chain of dependent
                // L2 round trips.
                constexpr int LB_BATCH = 8;
                const uint32_t* col = const_cast<const uint32_t*>(lookback) + d;
                long long t = tile - 1;
                bool done = false;
                while (!done) {
                    uint32_t s0;
                    do { s0 = ld_relaxed_gpu(col + (size_t)t * 256); } while ((s0 & (FLAG_AGG | FLAG_INC)) == 0u);
                    prefix += s0 & VAL_MASK;
                    if (s0 & FLAG_INC) break;
                    --t;
                    uint32_t sv[LB_BATCH];
#pragma unroll
                    for (int i = 0; i < LB_BATCH; ++i) sv[i] = (t - i >= 0) ? ld_relaxed_gpu(col + (size_t)(t - i) * 256) : FLAG_INC;
                    int used = 0;
#pragma unroll
                    for (int i = 0; i < LB_BATCH; ++i) {
                        if (!done && used == i && (sv[i] & (FLAG_AGG | FLAG_INC)) != 0u) {
                            prefix += sv[i] & VAL_MASK;
                            ++used;
                            if (sv[i] & FLAG_INC) done = true;
                        }
                    }
                    t -= used;
                }
                st_relaxed_gpu(lb + d, (prefix + cnt) | FLAG_INC);
            }
            sm.global_off[d] = digit_base[pass * 256 + d] + prefix;
            // block-wide exclusive scan of acc over the 256 digits -> local_off
            uint32_t inc = acc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t2 = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t2; }
            __shared__ uint32_t s_w[NW];
            if (lane == 31) s_w[warp] = inc;
            __syncthreads();
            uint32_t woff = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) if (w < warp) woff += s_w[w];
            sm.local_off[d] = woff + inc - acc;
        }
        __syncthreads();

        // ---- reorder through shared memory, then digit-contiguous global writes ----------------
#pragma unroll
        for (int i = 0; i < R3DG_SORT_ITEMS; ++i) {
            const uint32_t d = (uint32_t)(k[i] >> shift) & 0xffu;
            const uint32_t pos = sm.local_off[d] + sm.warp_hist[warp][d] + rank[i];
            sm.keys[pos] = k[i];
            sm.vals[pos] = v[i];
        }
        __syncthreads();
        for (int i = tid; i < count; i += R3DG_SORT_THREADS) {
            const uint64_t kk = sm.keys[i];
            const uint32_t d = (uint32_t)(kk >> shift) & 0xffu;
            const uint32_t dst = sm.global_off[d] + ((uint32_t)i - sm.local_off[d]);
            keys_out[dst] = kk;
            vals_out[dst] = sm.vals[i];
        }
    }
}

int launch_sort(void* geom_header, char* bin, const BinLayout& bl, int passes, int num_sms,
                cudaStream_t stream) {
    if (passes > R3DG_SORT_MAX_PASSES) return R3DG_ERR_UNSUPPORTED;
    GeomHeader* header = (GeomHeader*)geom_header;
    uint32_t* hist = (uint32_t*)(bin + bl.hist);
    uint32_t* lookback = (uint32_t*)(bin + bl.lookback);
    R3DG_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)R3DG_SORT_MAX_PASSES * 256 * 4, stream));
    sort_histogram_kernel<<<num_sms * 4, 256, 0, stream>>>(header, bl.capacity, passes,
                                                           (const uint64_t*)(bin + bl.keys_a), hist,
                                                           lookback, bl.max_tiles);
    sort_scan_kernel<<<1, 256, 0, stream>>>(passes, hist);
    static bool attr_set = false;
    const size_t smem = sizeof(SortSmem);
    if (!attr_set) {
        R3DG_CUDA_TRY(cudaFuncSetAttribute(sort_onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    uint64_t* ka = (uint64_t*)(bin + bl.keys_a); uint64_t* kb = (uint64_t*)(bin + bl.keys_b);
    uint32_t* va = (uint32_t*)(bin + bl.vals_a); uint32_t* vb = (uint32_t*)(bin + bl.vals_b);
    for (int k = 0; k < passes; ++k) {
        const bool even = (k & 1) == 0;
        sort_onesweep_kernel<<<num_sms * 3, R3DG_SORT_THREADS, smem, stream>>>(
            header, bl.capacity, k, passes, even ? ka : kb, even ? va : vb, even ? kb : ka, even ? vb : va, hist,
            (volatile uint32_t*)(lookback + (size_t)k * bl.max_tiles * 256));
    }
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace r3dg
