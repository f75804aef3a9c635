/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP) of the reference's BVH visibility path:
 *   - leaf AABBs of the 3-sigma oriented boxes   bvh/__init__.py:29-57 (PyTorch ops, fp32, no FMA)
 *   - LBVH build                                  bvh/src/construct.cu:7-266
 *   - opacity ray trace                           bvh/src/trace.cu:196-287, bvh/include/utility.cuh:35-111
 * Used only by tests/, smoke() and bench.py as the checker.  Parity status: PINNED against the
 * unmodified reference CUDA build (oracle/_ref/libref_bvh.so) via tests/golden/bvh_*.npz.
 * Compile with -ffp-contract=off (oracle/build.sh).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- bvh/__init__.py:31-57 + utils/general_utils.py:82-103 (build_rotation) ------------------ */
void oracle_bvh_leaf_aabbs(int P, const float* means3D, const float* scales, const float* rotations,
                           float* aabb /* [P][6] min xyz, max xyz */) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        const float* rq = rotations + 4 * i;
        const float norm = sqrtf(rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3]);
        const float r = rq[0] / norm, x = rq[1] / norm, y = rq[2] / norm, z = rq[3] / norm;
        float R[3][3];
        R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - r * z); R[0][2] = 2 * (x * z + r * y);
        R[1][0] = 2 * (x * y + r * z); R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - r * x);
        R[2][0] = 2 * (x * z - r * y); R[2][1] = 2 * (y * z + r * x); R[2][2] = 1 - 2 * (x * x + y * y);
        const float sa = 3 * scales[3 * i], sb = 3 * scales[3 * i + 1], sc = 3 * scales[3 * i + 2];
        for (int k = 0; k < 3; ++k) {      /* a,b,c = columns 0,1,2 of R; component k */
            const float m = means3D[3 * i + k];
            const float pa = m + R[k][0] * sa, ma = m - R[k][0] * sa;
            const float tb = R[k][1] * sb, tc = R[k][2] * sc;
            const float c8[8] = {(pa + tb) + tc, (pa + tb) - tc, (pa - tb) + tc, (pa - tb) - tc,
                                 (ma + tb) + tc, (ma + tb) - tc, (ma - tb) + tc, (ma - tb) - tc};
            float lo = c8[0], hi = c8[0];
            for (int j = 1; j < 8; ++j) { lo = fminf(lo, c8[j]); hi = fmaxf(hi, c8[j]); }
            aabb[6 * i + k] = lo;
            aabb[6 * i + 3 + k] = hi;
        }
    }
}

/* ---- construct.cu:7-51 ---------------------------------------------------------------------- */
static uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static uint32_t morton_code(float x, float y, float z) {
    const float res = 1024.0f;
    x = fminf(fmaxf(x * res, 0.0f), res - 1.0f);
    y = fminf(fmaxf(y * res, 0.0f), res - 1.0f);
    z = fminf(fmaxf(z * res, 0.0f), res - 1.0f);
    return expand_bits((uint32_t)x) * 4 + expand_bits((uint32_t)y) * 2 + expand_bits((uint32_t)z);
}
static int clz64(uint64_t v) { return v ? __builtin_clzll(v) : 64; }
static int common_upper_bits(uint64_t a, uint64_t b) { return clz64(a ^ b); }

/* construct.cu:53-112 */
static void determine_range(const uint64_t* code, uint32_t n, uint32_t idx, uint32_t* lo, uint32_t* hi) {
    if (idx == 0) { *lo = 0; *hi = n - 1; return; }
    const uint64_t self = code[idx];
    const int Ld = common_upper_bits(self, code[idx - 1]), Rd = common_upper_bits(self, code[idx + 1]);
    const int d = (Rd > Ld) ? 1 : -1;
    const int dmin = Ld < Rd ? Ld : Rd;
    int l_max = 2, delta = -1;
    long i_tmp = (long)idx + d * l_max;
    if (0 <= i_tmp && i_tmp < (long)n) delta = common_upper_bits(self, code[i_tmp]);
    while (delta > dmin) {
        l_max <<= 1;
        i_tmp = (long)idx + (long)d * l_max;
        delta = -1;
        if (0 <= i_tmp && i_tmp < (long)n) delta = common_upper_bits(self, code[i_tmp]);
    }
    int l = 0, t = l_max >> 1;
    while (t > 0) {
        i_tmp = (long)idx + (long)(l + t) * d;
        delta = -1;
        if (0 <= i_tmp && i_tmp < (long)n) delta = common_upper_bits(self, code[i_tmp]);
        if (delta > dmin) l += t;
        t >>= 1;
    }
    uint32_t jdx = idx + l * d;
    if (d < 0) { uint32_t tmp = idx; idx = jdx; jdx = tmp; }
    *lo = idx; *hi = jdx;
}
/* construct.cu:114-145 */
static int32_t find_split(const uint64_t* code, int32_t first, int32_t last) {
    const uint64_t fc = code[first], lc = code[last];
    if (fc == lc) return (first + last) >> 1;
    const int dn = common_upper_bits(fc, lc);
    int32_t split = first, stride = last - first;
    do {
        stride = (stride + 1) >> 1;
        const int middle = split + stride;
        if (middle < last && common_upper_bits(fc, code[middle]) > dn) split = middle;
    } while (stride > 1);
    return split;
}

/* construct_bvh: construct.cu:147-266.  nodes [2P-1][5] and aabbs [2P-1][6] pre-filled by the
 * caller like bvh/__init__.py:32-57 (leaf halves valid); both mutated in place; morton [P] out. */
typedef struct { uint32_t m; uint32_t idx; } mpair;
static int cmp_mpair(const void* a, const void* b) {
    const mpair* x = (const mpair*)a; const mpair* y = (const mpair*)b;
    if (x->m != y->m) return x->m < y->m ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);   /* == stable_sort_by_key */
}
void oracle_bvh_build(int P, int32_t* nodes, float* aabbs, uint64_t* morton) {
    const int NI = P - 1;
    float* leaf = aabbs + (size_t)NI * 6;
    float wlo[3] = {100000.f, 100000.f, 100000.f}, whi[3] = {-100000.f, -100000.f, -100000.f};
    for (int i = 0; i < P; ++i)
        for (int k = 0; k < 3; ++k) {
            wlo[k] = fminf(wlo[k], leaf[6 * i + k]);
            whi[k] = fmaxf(whi[k], leaf[6 * i + 3 + k]);
        }
    mpair* mp = (mpair*)malloc(sizeof(mpair) * (size_t)P);
    for (int i = 0; i < P; ++i) {
        float p[3];
        for (int k = 0; k < 3; ++k) {
            /* centroid: (upper + lower) * 0.5 (double literal; exact), then normalise */
            float c = (float)((double)(leaf[6 * i + 3 + k] + leaf[6 * i + k]) * 0.5);
            c -= wlo[k];
            c /= (whi[k] - wlo[k]);
            p[k] = c;
        }
        mp[i].m = morton_code(p[0], p[1], p[2]);
        mp[i].idx = (uint32_t)i;
    }
    qsort(mp, (size_t)P, sizeof(mpair), cmp_mpair);
    float* sorted = (float*)malloc(sizeof(float) * 6 * (size_t)P);
    for (int i = 0; i < P; ++i) {
        memcpy(sorted + 6 * i, leaf + 6 * (size_t)mp[i].idx, 6 * sizeof(float));
        morton[i] = ((uint64_t)mp[i].m << 31) | mp[i].idx;
        nodes[(size_t)(NI + i) * 5 + 3] = (int32_t)mp[i].idx;
    }
    memcpy(leaf, sorted, sizeof(float) * 6 * (size_t)P);
    free(sorted); free(mp);
    for (int idx = 0; idx < NI; ++idx) {
        int32_t* node = nodes + (size_t)idx * 5;
        node[3] = -1;
        uint32_t lo, hi;
        determine_range(morton, (uint32_t)P, (uint32_t)idx, &lo, &hi);
        const int32_t gamma = find_split(morton, (int32_t)lo, (int32_t)hi);
        node[1] = gamma; node[2] = gamma + 1;
        if ((int32_t)(lo < hi ? lo : hi) == gamma) node[1] += P - 1;
        if ((int32_t)(lo > hi ? lo : hi) == gamma + 1) node[2] += P - 1;
        nodes[(size_t)node[1] * 5] = idx;
        nodes[(size_t)node[2] * 5] = idx;
    }
    /* bottom-up refit + leaf counts (construct.cu:229-265); sequential emulation of the flags */
    uint8_t* flag = (uint8_t*)calloc((size_t)(NI > 0 ? NI : 1), 1);
    for (int leaf_i = NI; leaf_i < 2 * P - 1; ++leaf_i) {
        int32_t num = 1;
        int32_t parent = nodes[(size_t)leaf_i * 5];
        while (parent != -1) {
            nodes[(size_t)parent * 5 + 4] += num;
            if (!flag[parent]) { flag[parent] = 1; break; }
            int32_t* pn = nodes + (size_t)parent * 5;
            const float* lb = aabbs + (size_t)pn[1] * 6; const float* rb = aabbs + (size_t)pn[2] * 6;
            float* pb = aabbs + (size_t)parent * 6;
            for (int k = 0; k < 3; ++k) { pb[k] = fminf(lb[k], rb[k]); pb[3 + k] = fmaxf(lb[3 + k], rb[3 + k]); }
            num = pn[4];
            parent = pn[0];
        }
    }
    free(flag);
}

/* ---- utility.cuh:35-82 slab test (IEEE divisions by possibly-zero components kept) -------- */
static void ray_box(const float* b, const float* o, const float* d, float* tmin_o, float* tmax_o) {
    float tmin = (b[0] - o[0]) / d[0], tmax = (b[3] - o[0]) / d[0];
    if (tmin > tmax) { float t = tmin; tmin = tmax; tmax = t; }
    float tymin = (b[1] - o[1]) / d[1], tymax = (b[4] - o[1]) / d[1];
    if (tymin > tymax) { float t = tymin; tymin = tymax; tymax = t; }
    if (tmin > tymax || tymin > tmax) { *tmin_o = -1.f; *tmax_o = -1.f; return; }
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (b[2] - o[2]) / d[2], tzmax = (b[5] - o[2]) / d[2];
    if (tzmin > tzmax) { float t = tzmin; tzmin = tzmax; tzmax = t; }
    if (tmin > tzmax || tzmin > tmax) { *tmin_o = -1.f; *tmax_o = -1.f; return; }
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    *tmin_o = tmin; *tmax_o = tmax;
}

/* trace_bvh_opacity_cuda: trace.cu:196-287.  One ray per iteration; the stack discipline (far
 * child pushed first) fixes the visiting order and hence the product order of (1 - alpha). */
void oracle_bvh_trace_opacity(int64_t num_rays, const int32_t* nodes, const float* aabbs,
                              const float* rays_o, const float* rays_d, const float* means3D,
                              const float* covs3D, const float* opacities, const float* normals,
                              int32_t* num_contributes /* zero-init */, float* rendered_opacity /* one-init */) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t idx = 0; idx < num_rays; ++idx) {
        int32_t stack[64]; int sp = 0;
        stack[sp++] = 0;
        int32_t count = 0;
        const float* o = rays_o + 3 * idx; const float* d = rays_d + 3 * idx;
        float ray_opacity = 1.0f;
        int terminated = 0;
        while (sp > 0 && !terminated) {
            const int32_t node_id = stack[--sp];
            const int32_t* node = nodes + (size_t)node_id * 5;
            if (node[4] <= 1) {
                const int32_t g = node[3];
                if (opacities[g] < 1.f / 255.f) continue;
                const float* nrm = normals + 3 * (size_t)g;
                if (nrm[0] * d[0] + nrm[1] * d[1] + nrm[2] * d[2] > 0) continue;
                const float* mu = means3D + 3 * (size_t)g; const float* ci = covs3D + 6 * (size_t)g;
                const float m0 = mu[0] - o[0], m1 = mu[1] - o[1], m2 = mu[2] - o[2];
                const float t1 = ci[0] * m0 * d[0] + ci[1] * m0 * d[1] + ci[2] * m0 * d[2] +
                                 ci[1] * m1 * d[0] + ci[3] * m1 * d[1] + ci[4] * m1 * d[2] +
                                 ci[2] * m2 * d[0] + ci[4] * m2 * d[1] + ci[5] * m2 * d[2];
                const float t2 = ci[0] * d[0] * d[0] + ci[1] * d[0] * d[1] + ci[2] * d[0] * d[2] +
                                 ci[1] * d[1] * d[0] + ci[3] * d[1] * d[1] + ci[4] * d[1] * d[2] +
                                 ci[2] * d[2] * d[0] + ci[4] * d[2] * d[1] + ci[5] * d[2] * d[2];
                const float t = t1 / t2;
                if (t < 0.01f) continue;
                const float e0 = mu[0] - (o[0] + t * d[0]), e1 = mu[1] - (o[1] + t * d[1]), e2 = mu[2] - (o[2] + t * d[2]);
                const float power = -0.5f * (e0 * e0 * ci[0] + e1 * e1 * ci[3] + e2 * e2 * ci[5] +
                                             2 * e0 * e1 * ci[1] + 2 * e0 * e2 * ci[2] + 2 * e1 * e2 * ci[4]);
                if (power > 0) continue;
                count += 1;
                const float alpha = opacities[g] * expf(power);
                ray_opacity *= 1 - alpha;
                if (ray_opacity < 0.9f) { rendered_opacity[idx] = 0.0f; terminated = 1; }
            } else {
                const int32_t lid = node[1], rid = node[2];
                float lmin, lmax, rmin, rmax;
                ray_box(aabbs + (size_t)lid * 6, o, d, &lmin, &lmax);
                ray_box(aabbs + (size_t)rid * 6, o, d, &rmin, &rmax);
                if (lmax > rmax) {
                    if (lmax > 0) stack[sp++] = lid;
                    if (rmax > 0) stack[sp++] = rid;
                } else {
                    if (rmax > 0) stack[sp++] = rid;
                    if (lmax > 0) stack[sp++] = lid;
                }
            }
        }
        if (!terminated) { num_contributes[idx] = count; rendered_opacity[idx] = ray_opacity; }
    }
}
