// Stable LSD radix sort of (u32 key, u32 value) pairs.  In the rasterizer it orders the Gaussians
// by view depth once (P pairs) instead of the reference's cub::DeviceRadixSort::SortPairs over
// all R ~ 8 P (tile, depth) instances (rasterizer_impl.cu:313-318); the per-tile lists are then
// produced by the order-preserving binning in binning.cu.  The BVH and kNN builders sort their
// 30-bit Morton codes with it (bvh.cu, knn.cu).  Hand-written one-sweep design for sm_100a:
//
//   1. sort_histogram_kernel : one read of the keys builds the digit histograms of ALL passes;
//   2. sort_onesweep_kernel  : per pass ONE read + ONE write of the pairs.  A persistent grid
//      pulls 2048-pair tiles from an atomic ticket.  Per tile: a shared-memory histogram gives the
//      tile's digit counts, which are published for the successors BEFORE the expensive ranking
//      (so nobody waits on it); warp-synchronous match-any ranking (stable); two-level decoupled
//      look-back per digit; reorder through shared memory so that the global writes are
//      digit-contiguous runs.
//      Look-back: the sorts here are small (P ~ 1M pairs = ~500 tiles, all resident at once), so a
//      plain chained look-back degenerates into every tile walking hundreds of predecessors.
//      Tiles are grouped by 32: a tile sums the early-published aggregates of the <= 31 earlier
//      tiles of its group in ONE batch of independent loads; the last tile of a group publishes
//      the group aggregate, and group prefixes are chained by the usual aggregate/inclusive
//      look-back one level up.  Depth of the dependency chain: 2-3 memory round trips.
//
// Which key bits need sorting is decided on the device (sort_plan in common.cuh): no host sync.
#include "common.cuh"
#include "kernels.h"

namespace r3dg {

#define FLAG_AGG 0x40000000u
#define FLAG_INC 0x80000000u
#define VAL_MASK 0x3fffffffu
#define SORT_GROUP 32            // tiles per look-back group

__global__ void __launch_bounds__(256) sort_histogram_kernel(GeomHeader* __restrict__ header, long long n,
                                                             const uint32_t* __restrict__ keys,
                                                             uint32_t* __restrict__ hist,
                                                             uint32_t* __restrict__ lookback0, long long max_tiles) {
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
    const long long ngroups = (ntiles + SORT_GROUP - 1) / SORT_GROUP;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ntiles * 256; i += stride) lookback0[i] = 0;
    uint32_t* grp0 = lookback0 + (size_t)max_tiles * 256;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ngroups * 256; i += stride) grp0[i] = 0;
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

__global__ void __launch_bounds__(R3DG_SORT_THREADS, 4) sort_onesweep_kernel(
    GeomHeader* header, long long n, int slot, const uint32_t* __restrict__ keys_in,
    const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
    const uint32_t* __restrict__ hist, uint32_t* lb_cur, uint32_t* lb_next, long long max_tiles) {
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
    uint32_t* grp_cur = lb_cur + (size_t)max_tiles * 256;    // group-level descriptors of this pass
    uint32_t* grp_next = lb_next + (size_t)max_tiles * 256;

    // global start of digit tid: exclusive scan of this pass's histogram (thread d <-> digit d)
    uint32_t digit_base;
    {
        const uint32_t hv = hist[slot * 256 + tid];
        uint32_t inc = hv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t2 = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t2; }
        if (lane == 31) sm.s_w[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) if (ww < warp) woff += sm.s_w[ww];
        digit_base = woff + inc - hv;
    }

    // Tile order.  When every tile has its own CTA (ntiles <= gridDim.x: sorts up to ~1.2M pairs, i.e.
    // the usual case here) tile = blockIdx.x and the CTA retires after it: CTAs are dispatched in
    // index order, so whatever a tile waits for is running or done.  Handing those tiles out
    // through an atomic ticket would serialise ~500 same-address L2 atomics (~10 us) in front of
    // a ~5 us tile.  Larger sorts keep the persistent grid + ticket (start order = tile order).
    const bool one_shot = ntiles <= (long long)gridDim.x;
    bool first_tile = true;
    while (true) {
        __syncthreads();
        if (one_shot) { if (!first_tile) break; if (tid == 0) sm.tile = blockIdx.x; }
        else if (tid == 0) sm.tile = atomicAdd(&header->sort_ticket[slot], 1u);
        first_tile = false;
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
        if (d < bins) st_relaxed_gpu(lb_cur + (size_t)tile * 256 + d, cnt | FLAG_AGG);
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
            // peers with the same digit, from w ballots: a match.any over ~32 distinct values costs
            // ~400 cycles on sm_100, eight ballots ~140 (profiles/r01_ubench_warp_primitives.jsonl)
            uint32_t peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (b < w) {
                    const bool bit = (dg[i] >> b) & 1u;
                    const uint32_t m = __ballot_sync(0xffffffffu, bit);
                    peers &= bit ? m : ~m;
                }
            }
            const int leader = __ffs(peers) - 1;
            const uint32_t below = __popc(peers & ((1u << lane) - 1u));
            uint32_t old = 0;
            if (lane == leader) { old = sm.warp_hist[warp][dg[i]]; sm.warp_hist[warp][dg[i]] = old + __popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[i] = old + below;
            __syncwarp();
        }
        __syncthreads();

        // ---- per digit: warp prefixes; two-level look-back -------------------------------------
        {
            uint32_t acc = 0;
#pragma unroll
            for (int ww = 0; ww < NW; ++ww) { const uint32_t t = sm.warp_hist[ww][d]; sm.warp_hist[ww][d] = acc; acc += t; }
            uint32_t prefix = 0;
            const long long grp = tile / SORT_GROUP;
            if (d < bins) {
                // (1) the earlier tiles of this group: their aggregates were published before their
                //     ranking, one batch of independent loads normally finds all of them
                const int m = (int)(tile - grp * SORT_GROUP);
                const uint32_t* col = lb_cur + (size_t)grp * SORT_GROUP * 256 + d;
                uint32_t pending = m ? (0xffffffffu >> (32 - m)) : 0u;
                uint32_t within = 0;
                while (pending) {
#pragma unroll
                    for (int i = 0; i < SORT_GROUP - 1; ++i) {
                        if (pending & (1u << i)) {
                            const uint32_t v = ld_relaxed_gpu(col + (size_t)i * 256);
                            if (v & FLAG_AGG) { within += v & VAL_MASK; pending &= ~(1u << i); }
                        }
                    }
                }
                // (2) the earlier groups: chained aggregate / inclusive look-back one level up
                const bool closes_group = m == SORT_GROUP - 1;
                if (closes_group) st_relaxed_gpu(grp_cur + (size_t)grp * 256 + d, (within + cnt) | (grp == 0 ? FLAG_INC : FLAG_AGG));
                uint32_t before = 0;
                if (grp > 0) {
                    const uint32_t* gcol = grp_cur + d;
                    long long t = grp - 1;
                    if (!lookback_step<1>(gcol, t, before))
                        if (!lookback_step<8>(gcol, t, before))
                            while (!lookback_step<32>(gcol, t, before)) {}
                    if (closes_group) st_relaxed_gpu(grp_cur + (size_t)grp * 256 + d, (before + within + cnt) | FLAG_INC);
                }
                prefix = before + within;
                lb_next[(size_t)tile * 256 + d] = 0;       // descriptor planes of the next pass
                if (closes_group) grp_next[(size_t)grp * 256 + d] = 0;
            }
            sm.global_off[d] = digit_base + prefix;
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
    uint32_t* lb1 = lb0 + (size_t)sl.plane_words;
    R3DG_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)R3DG_SORT_MAX_PASSES * 256 * 4, stream));
    const long long ntiles = (n + R3DG_SORT_TILE - 1) / R3DG_SORT_TILE;
    const int hb = (int)std::min<long long>((long long)num_sms * 4, std::max<long long>(1, (n + 2047) / 2048));
    sort_histogram_kernel<<<hb, 256, 0, stream>>>(header, n, (const uint32_t*)(buf + sl.keys_a), hist, lb0, sl.tiles);
    uint32_t* ka = (uint32_t*)(buf + sl.keys_a); uint32_t* kb = (uint32_t*)(buf + sl.keys_b);
    uint32_t* va = (uint32_t*)(buf + sl.vals_a); uint32_t* vb = (uint32_t*)(buf + sl.vals_b);
    const int grid = (int)std::min<long long>((long long)num_sms * 4, std::max<long long>(1, ntiles));   // resident: 4 CTAs / SM
    for (int k = 0; k < R3DG_SORT_MAX_PASSES; ++k) {
        const bool even = (k & 1) == 0;
        sort_onesweep_kernel<<<grid, R3DG_SORT_THREADS, 0, stream>>>(
            header, n, k, even ? ka : kb, even ? va : vb, even ? kb : ka, even ? vb : va, hist,
            even ? lb0 : lb1, even ? lb1 : lb0, sl.tiles);
    }
    R3DG_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace r3dg
