// Stable LSD radix sort of (u32 key, u32 value) pairs.  In the rasterizer it orders the Gaussians
// by view depth once (P pairs) instead of the reference's cub::DeviceRadixSort::SortPairs over
// all R ~ 8 P (tile, depth) instances (rasterizer_impl.cu:313-318); the per-tile lists are then
// produced by the order-preserving binning in binning.cu.  The BVH and kNN builders sort their
// 30-bit Morton codes with it (bvh.cu, knn.cu).  Hand-written one-sweep design for sm_100a:
//
//   1. sort_histogram_kernel : one read of the keys builds the digit histograms of ALL passes;
//   2. sort_scan_kernel      : exclusive scan of each histogram -> global digit bases;
//   3. sort_onesweep_kernel  : per pass ONE read + ONE write of the pairs.  A persistent grid
//      pulls 2048-pair tiles from an atomic ticket.  Per tile: a shared-memory histogram gives the
//      tile's digit counts, which are published for the successors BEFORE the expensive ranking
//      (so nobody waits on it); warp-synchronous match-any ranking (stable); decoupled look-back
//      per digit with escalating batches of independent loads; reorder through shared memory so
//      that the global writes are digit-contiguous runs.
//
// Which key bits need sorting is decided on the device (sort_plan in common.cuh): no host sync.
#include "common.cuh"
#include "kernels.h"

namespace r3dg {

#define FLAG_AGG 0x40000000u
#define FLAG_INC 0x80000000u
#define VAL_MASK 0x3fffffffu

__global__ void __launch_bounds__(256) sort_histogram_kernel(GeomHeader* __restrict__ header, long long n,
                                                             const uint32_t* __restrict__ keys,
                                                             uint32_t* __restrict__ hist,
                                                             uint32_t* __restrict__ lookback0) {
    __shared__ uint32_t sh[R3DG_SORT_MAX_PASSES * 256];
    int passes, w;
    sort_plan(header->depth_or & header->depth_nor, passes, w);
    if (blockIdx.x == 0 && threadIdx.x == 0) header->sort_exec = (uint32_t)passes;
    const uint32_t mask = (1u << w) - 1u;
    for (int i = threadIdx.x; i < R3DG_SORT_MAX_PASSES * 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * 256 + ((k >> (p * w)) & mask)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
    // clear descriptor plane 0 (pass 0); every pass clears the other plane for its successor
    const long long ntiles = (n + R3DG_SORT_TILE - 1) / R3DG_SORT_TILE;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ntiles * 256; i += stride) lookback0[i] = 0;
}

__global__ void __launch_bounds__(256) sort_scan_kernel(uint32_t* __restrict__ hist) {
    // one block; thread d owns bin d of every pass; exclusive scan through shared memory
    __shared__ uint32_t s[256];
    for (int p = 0; p < R3DG_SORT_MAX_PASSES; ++p) {
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

// One look-back step: B descriptors of consecutive predecessors are fetched with independent
// loads and consumed in order while they are ready.  Returns true when an inclusive prefix ended
// the walk.
template <int B>
__device__ __forceinline__ bool lookback_step(const uint32_t* col, long long& t, uint32_t& prefix) {
    uint32_t sv[B];
#pragma unroll
    for (int i = 0; i < B; ++i) sv[i] = (t - i >= 0) ? ld_relaxed_gpu(col + (size_t)(t - i) * 256) : FLAG_INC;
    int used = 0;
    bool done = false;
#pragma unroll
    for (int i = 0; i < B; ++i) {
        if (!done && used == i && (sv[i] & (FLAG_AGG | FLAG_INC)) != 0u) {
            prefix += sv[i] & VAL_MASK;
            ++used;
            if (sv[i] & FLAG_INC) done = true;
        }
    }
    t -= used;
    return done;
}

struct SortSmem {
    uint32_t keys[R3DG_SORT_TILE];
    uint32_t vals[R3DG_SORT_TILE];
    uint32_t warp_hist[R3DG_SORT_THREADS / 32][256];
    uint32_t block_hist[256];    // digit counts of the tile (valid keys only)
    uint32_t local_off[256];     // exclusive prefix of the digit counts inside the tile
    uint32_t global_off[256];    // global start of this tile's run of digit d
    uint32_t s_w[R3DG_SORT_THREADS / 32];
    uint32_t tile;
};

__global__ void __launch_bounds__(R3DG_SORT_THREADS) sort_onesweep_kernel(
    GeomHeader* header, long long n, int slot, const uint32_t* __restrict__ keys_in,
    const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
    const uint32_t* __restrict__ digit_base, uint32_t* lb_cur, uint32_t* lb_next) {
    __shared__ SortSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int passes, w;
    sort_plan(header->depth_or & header->depth_nor, passes, w);
    if (slot >= passes) return;                          // fewer digits needed than slots launched
    const int shift = slot * w;
    const uint32_t mask = (1u << w) - 1u;
    const int bins = 1 << w;
    const long long ntiles = (n + R3DG_SORT_TILE - 1) / R3DG_SORT_TILE;
    constexpr int NW = R3DG_SORT_THREADS / 32;
    constexpr int PER_WARP = R3DG_SORT_TILE / NW;        // 256 consecutive pairs per warp

    while (true) {
        __syncthreads();
        if (tid == 0) sm.tile = atomicAdd(&header->sort_ticket[slot], 1u);
        for (int i = tid; i < NW * 256; i += R3DG_SORT_THREADS) (&sm.warp_hist[0][0])[i] = 0;
        sm.block_hist[tid] = 0;
        __syncthreads();
        const long long tile = sm.tile;
        if (tile >= ntiles) break;
        const long long tile_base = tile * R3DG_SORT_TILE;
        const int count = (int)min((long long)R3DG_SORT_TILE, n - tile_base);

        // ---- load; tile digit counts through shared atomics ----------------------------------
        uint32_t k[R3DG_SORT_ITEMS], v[R3DG_SORT_ITEMS], dg[R3DG_SORT_ITEMS], rank[R3DG_SORT_ITEMS];
#pragma unroll
        for (int i = 0; i < R3DG_SORT_ITEMS; ++i) {
            const int local = warp * PER_WARP + i * 32 + lane;
            const bool valid = local < count;
            k[i] = valid ? keys_in[tile_base + local] : 0u;
            v[i] = valid ? vals_in[tile_base + local] : 0u;
        }
#pragma unroll
        for (int i = 0; i < R3DG_SORT_ITEMS; ++i) {
            const bool valid = warp * PER_WARP + i * 32 + lane < count;
            // padding slots take the last digit: they trail every valid key of that digit (they are
            // the tail of the tile), so valid ranks are unaffected, and they are never stored
            dg[i] = valid ? ((k[i] >> shift) & mask) : mask;
            if (valid) atomicAdd(&sm.block_hist[dg[i]], 1u);
        }
        __syncthreads();

        // ---- publish the aggregate at once; tile-local exclusive offsets ----------------------
        const int d = tid;                                // R3DG_SORT_THREADS == 256 >= bins
        const uint32_t cnt = sm.block_hist[d];
        uint32_t* lb = lb_cur + (size_t)tile * 256;
        if (d < bins) st_relaxed_gpu(lb + d, cnt | (tile == 0 ? FLAG_INC : FLAG_AGG));
        {
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t2 = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t2; }
            if (lane == 31) sm.s_w[warp] = inc;
            __syncthreads();
            uint32_t woff = 0;
#pragma unroll
            for (int ww = 0; ww < NW; ++ww) if (ww < warp) woff += sm.s_w[ww];
            sm.local_off[d] = woff + inc - cnt;
        }

        // ---- stable ranking within the warp's 256-pair slice ----------------------------------
#pragma unroll
        for (int i = 0; i < R3DG_SORT_ITEMS; ++i) {
            const uint32_t peers = __match_any_sync(0xffffffffu, dg[i]);
            const int leader = __ffs(peers) - 1;
            const uint32_t below = __popc(peers & ((1u << lane) - 1u));
            uint32_t old = 0;
            if (lane == leader) { old = sm.warp_hist[warp][dg[i]]; sm.warp_hist[warp][dg[i]] = old + __popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[i] = old + below;
            __syncwarp();
        }
        __syncthreads();

        // ---- per digit: warp prefixes; decoupled look-back ------------------------------------
        {
            uint32_t acc = 0;
#pragma unroll
            for (int ww = 0; ww < NW; ++ww) { const uint32_t t = sm.warp_hist[ww][d]; sm.warp_hist[ww][d] = acc; acc += t; }
            uint32_t prefix = 0;
            if (tile > 0 && d < bins) {
                const uint32_t* col = lb_cur + d;
                long long t = tile - 1;
                // predecessors published their aggregates before ranking, so the walk rarely waits:
                // 1 load, then 8, then 32 independent loads per step
                if (!lookback_step<1>(col, t, prefix))
                    if (!lookback_step<8>(col, t, prefix))
                        while (!lookback_step<32>(col, t, prefix)) {}
                st_relaxed_gpu(lb + d, (prefix + cnt) | FLAG_INC);
            }
            sm.global_off[d] = digit_base[slot * 256 + d] + prefix;
            lb_next[(size_t)tile * 256 + d] = 0;          // descriptor plane of the next pass
        }
        __syncthreads();

        // ---- reorder through shared memory, then digit-contiguous global writes ----------------
#pragma unroll
        for (int i = 0; i < R3DG_SORT_ITEMS; ++i) {
            const bool valid = warp * PER_WARP + i * 32 + lane < count;
            if (valid) {
                const uint32_t pos = sm.local_off[dg[i]] + sm.warp_hist[warp][dg[i]] + rank[i];
                sm.keys[pos] = k[i];
                sm.vals[pos] = v[i];
            }
        }
        __syncthreads();
        for (int i = tid; i < count; i += R3DG_SORT_THREADS) {
            const uint32_t kk = sm.keys[i];
            const uint32_t dd = (kk >> shift) & mask;
            const uint32_t dst = sm.global_off[dd] + ((uint32_t)i - sm.local_off[dd]);
            keys_out[dst] = kk;
            vals_out[dst] = sm.vals[i];
        }
    }
}

// header: depth_or / depth_nor hold the OR of the keys and of their complements, sort_ticket[]
// is zero.  On return (stream order) the sorted pairs are in buffer (header->sort_exec & 1).
int launch_sort(void* geom_header, char* buf, const SortLayout& sl, long long n, int num_sms,
                cudaStream_t stream) {
    if (n > sl.n || n >= (1ll << 30)) return R3DG_ERR_BAD_ARG;
    GeomHeader* header = (GeomHeader*)geom_header;
    uint32_t* hist = (uint32_t*)(buf + sl.hist);
    uint32_t* lb0 = (uint32_t*)(buf + sl.lookback);
    uint32_t* lb1 = lb0 + (size_t)sl.tiles * 256;
    R3DG_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)R3DG_SORT_MAX_PASSES * 256 * 4, stream));
    const long long ntiles = (n + R3DG_SORT_TILE - 1) / R3DG_SORT_TILE;
    const int hb = (int)std::min<long long>((long long)num_sms * 4, std::max<long long>(1, (n + 2047) / 2048));
    sort_histogram_kernel<<<hb, 256, 0, stream>>>(header, n, (const uint32_t*)(buf + sl.keys_a), hist, lb0);
    sort_scan_kernel<<<1, 256, 0, stream>>>(hist);
    uint32_t* ka = (uint32_t*)(buf + sl.keys_a); uint32_t* kb = (uint32_t*)(buf + sl.keys_b);
    uint32_t* va = (uint32_t*)(buf + sl.vals_a); uint32_t* vb = (uint32_t*)(buf + sl.vals_b);
    const int grid = (int)std::min<long long>((long long)num_sms * 4, std::max<long long>(1, ntiles));
    for (int k = 0; k < R3DG_SORT_MAX_PASSES; ++k) {
        const bool even = (k & 1) == 0;
        sort_onesweep_kernel<<<grid, R3DG_SORT_THREADS, 0, stream>>>(
            header, n, k, even ? ka : kb, even ? va : vb, even ? kb : ka, even ? vb : va, hist,
            even ? lb0 : lb1, even ? lb1 : lb0);
    }
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace r3dg
