// knn_search.cuh -- the search half of distCUDA2 (mean squared distance to the 3 nearest neighbours), written so that the same
// code compiles for the device (knn.cu) and for the host (tests/knn_search_host.cpp: the search logic is checked on CPU against a
// brute force before it ever sees a GPU).
//
// What the reference does (submodules/simple-knn/simple_knn.cu:147-183, boxMeanDist): sort by a 30-bit Morton code, cut the sorted
// list into boxes of 1024 points, and let EVERY point test EVERY box (P * P/1024 box tests, plus 1024 distance evaluations for each
// box that survives).  The result is the exact 3-NN mean; the arithmetic of one distance is fma(dz,dz, fma(dx,dx, dy*dy)) with
// d = other - query, and the result is ((b0 + b1) + b2) / 3 (cuobjdump of the reference build, recorded in DESIGN.md).
//
// What this does instead: 63-bit Morton codes (21 bits per axis) over the cubic bounding box, sorted once.  Every octree cell of
// every level is then one contiguous range of the sorted list (a prefix of the code), found through a dense table for the first
// KnnIndex::T levels and a short binary search inside the table range below that.  A query
//   1. picks the finest level at which its own cell holds >= 4 points (common prefix with its sorted neighbours j-3 .. j+3),
//   2. scans the 3x3x3 block of cells around it at that level (own cell first, the others pruned by box distance),
//   3. is done when its third-best distance is no larger than the distance to the nearest face of the block that has cells behind
//      it; otherwise it repeats one level up (at most twice in practice, because the own cell already holds 3 neighbours).
// The work per query is a few hundred distance evaluations whatever the density, instead of P/1024 box tests + thousands of
// evaluations, and it does not degrade when a few far outliers blow up the bounding box (the level is chosen per query).
// Results are bit-identical to the reference: same distance arithmetic, and the set of three smallest values does not depend on the
// visiting order.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KNN_HD __host__ __device__ __forceinline__
#else
#define KNN_HD inline
#endif

#define KNN_BITS 21            // bits per axis of the Morton code
#define KNN_LMAX 19            // finest level a query may start at (cells of 4 code units: the 1-unit rounding margin stays < 25 %)
#define KNN_MARGIN 1.0f        // bound on the error of a difference of two computed cell coordinates, in code units (see knn_cell)

struct alignas(16) KnnPoint { float x, y, z; uint32_t idx; };   // 16 bytes (one 128-bit load): position + original index, in Morton order

struct KnnGrid {               // cubic bounding box: u = (p - o) * scale in [0, 2^21], unit = side * 2^-21
    float ox, oy, oz, scale, unit;
};

struct KnnIndex {
    const uint64_t* code;      // [P] sorted Morton codes
    const KnnPoint* pts;       // [P] points in the same order
    const uint32_t* table;     // [8^T + 1] first sorted position whose T-level prefix is >= t
    int P, T;
};

KNN_HD float knn_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
KNN_HD float knn_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;     // host build uses -ffp-contract=off as well; volatile keeps it a separate rounding anyway
    return r;
#endif
}
KNN_HD float knn_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
KNN_HD int knn_clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

// 21 bits -> every third bit of 63
KNN_HD uint64_t knn_expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
KNN_HD uint64_t knn_morton(uint32_t cx, uint32_t cy, uint32_t cz) { return knn_expand21(cx) | (knn_expand21(cy) << 1) | (knn_expand21(cz) << 2); }

KNN_HD KnnGrid knn_make_grid(float mnx, float mny, float mnz, float mxx, float mxy, float mxz) {
    KnnGrid g;
    g.ox = mnx; g.oy = mny; g.oz = mnz;
    const float side = fmaxf(fmaxf(mxx - mnx, mxy - mny), mxz - mnz);
    if (side > 1e-30f && side < 3e38f) { g.scale = 2097152.0f / side; g.unit = side * (1.0f / 2097152.0f); }
    else { g.scale = 0.0f; g.unit = 0.0f; }          // one point, or all points identical: everything lands in cell 0
    return g;
}

// Cell coordinates of a point, in code units.  Every kernel calls this with the same inputs, so the assignment is consistent.
// Error of u against the exact (p - o) / unit: the subtraction and the product round once each, 2^-24 relative to a value
// <= 2^21, i.e. <= 0.25 units per point and <= 0.5 for a difference of two -- KNN_MARGIN = 1 covers it twice.
KNN_HD void knn_cell(const KnnGrid& g, float x, float y, float z, float* u, uint32_t* c) {
    u[0] = knn_mul(knn_sub(x, g.ox), g.scale); u[1] = knn_mul(knn_sub(y, g.oy), g.scale); u[2] = knn_mul(knn_sub(z, g.oz), g.scale);
    for (int a = 0; a < 3; ++a) {
        float v = u[a];
        if (!(v > 0.0f)) v = 0.0f;
        uint32_t ci = v >= 2097151.0f ? 2097151u : (uint32_t)v;
        c[a] = ci;
    }
}

KNN_HD void knn_insert(float* best, float d) {        // simple_knn.cu:130-144 updateKBest<3>
    for (int j = 0; j < 3; ++j) {
        if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
    }
}

KNN_HD uint32_t knn_lower_bound(const uint64_t* code, uint32_t a, uint32_t b, uint64_t key) {
    while (a < b) {
        const uint32_t m = a + ((b - a) >> 1);
        if (code[m] < key) a = m + 1; else b = m;
    }
    return a;
}

// [lo, hi) of the level-`level` cell whose Morton prefix (3 * level bits) is `key`
KNN_HD void knn_range(const KnnIndex& ix, int level, uint64_t key, uint32_t* lo, uint32_t* hi) {
    if (level <= ix.T) {
        const int sh = 3 * (ix.T - level);
        *lo = ix.table[key << sh];
        *hi = ix.table[(key + 1) << sh];
    } else {
        const uint64_t anc = key >> (3 * (level - ix.T));
        const uint32_t a = ix.table[anc], b = ix.table[anc + 1];
        if (a == b) { *lo = *hi = a; return; }
        const int sh = 63 - 3 * level;
        *lo = knn_lower_bound(ix.code, a, b, key << sh);
        *hi = knn_lower_bound(ix.code, *lo, b, (key + 1) << sh);
    }
}

// finest level (<= KNN_LMAX) at which the cell of sorted position j holds at least 4 points
KNN_HD int knn_start_level(const KnnIndex& ix, int j) {
    int level = 0;
    const uint64_t* c = ix.code;
    for (int a = j - 3; a <= j; ++a) {
        const int b = a + 3;
        if (a < 0 || b >= ix.P) continue;
        const uint64_t x = c[a] ^ c[b];
        const int common = x ? knn_clz64(x) - 1 : 63;          // shared leading bits of the 63-bit codes
        const int l = common / 3;
        if (l > level) level = l;
    }
    return level < KNN_LMAX ? level : KNN_LMAX;
}

struct KnnStats { uint32_t evals, cells, rounds; };            // host-side instrumentation (tests); the device passes nullptr

KNN_HD float knn_query(const KnnIndex& ix, const KnnGrid& g, int j, KnnStats* st) {
    const KnnPoint q = ix.pts[j];
    float u[3];
    uint32_t c0[3];
    knn_cell(g, q.x, q.y, q.z, u, c0);
    int level = knn_start_level(ix, j);
    float best[3];
    const float unit2 = g.unit * g.unit;
    for (;;) {
        const int shift = KNN_BITS - level;
        const float S = (float)(1u << shift);                  // cell side in code units
        const int G = 1 << level;
        int ck[3];
        float lo_d[3], hi_d[3];                                // distance (code units) from the query to the low / high face of its cell
        float m = FLT_MAX;                                     // distance to the nearest block face that has cells behind it
        for (int a = 0; a < 3; ++a) {
            ck[a] = (int)(c0[a] >> shift);
            lo_d[a] = u[a] - (float)ck[a] * S;
            hi_d[a] = (float)(ck[a] + 1) * S - u[a];
            if (ck[a] >= 2) m = fminf(m, lo_d[a] + S);
            if (ck[a] + 2 <= G - 1) m = fminf(m, hi_d[a] + S);
        }
        best[0] = best[1] = best[2] = FLT_MAX;
        if (st) st->rounds++;
        for (int t = 0; t < 27; ++t) {                          // offsets in the order 0, -1, +1 per axis: own cell first
            const int o[3] = {(t % 3 == 0) ? 0 : (t % 3 == 1 ? -1 : 1), ((t / 3) % 3 == 0) ? 0 : ((t / 3) % 3 == 1 ? -1 : 1),
                              (t / 9 == 0) ? 0 : (t / 9 == 1 ? -1 : 1)};
            const int nx = ck[0] + o[0], ny = ck[1] + o[1], nz = ck[2] + o[2];
            if (nx < 0 || ny < 0 || nz < 0 || nx >= G || ny >= G || nz >= G) continue;
            if (t > 0) {                                        // box distance, shrunk by the rounding margin per axis
                float bd2 = 0.0f;
                for (int a = 0; a < 3; ++a) {
                    float d = o[a] == 0 ? 0.0f : (o[a] < 0 ? lo_d[a] : hi_d[a]);
                    d = fmaxf(d - KNN_MARGIN, 0.0f);
                    bd2 += d * d;
                }
                if (bd2 * unit2 * 0.99999f > best[2]) continue;
            }
            uint32_t lo, hi;
            knn_range(ix, level, knn_morton((uint32_t)nx, (uint32_t)ny, (uint32_t)nz), &lo, &hi);
            if (st) st->cells++;
            for (uint32_t i = lo; i < hi; ++i) {
                const KnnPoint p = ix.pts[i];
                if (p.idx == q.idx) continue;                   // simple_knn.cu:156,174: only the query itself is skipped
                const float dx = knn_sub(p.x, q.x), dy = knn_sub(p.y, q.y), dz = knn_sub(p.z, q.z);
                const float d = knn_fma(dz, dz, knn_fma(dx, dx, knn_mul(dy, dy)));   // nvcc's contraction of simple_knn.cu:134
                knn_insert(best, d);
                if (st) st->evals++;
                if (best[2] == 0.0f) break;                     // three coincident points: nothing can improve
            }
            if (best[2] == 0.0f) break;
        }
        if (m == FLT_MAX || level == 0 || best[2] == 0.0f) break;   // the block covered everything there is
        const float lb = (m - KNN_MARGIN) * g.unit * 0.999999f;     // every point outside the block is at least this far
        if (lb > 0.0f && best[2] <= lb * lb) break;
        --level;
    }
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(__fadd_rn(__fadd_rn(best[0], best[1]), best[2]), 3.0f);
#else
    return ((best[0] + best[1]) + best[2]) / 3.0f;
#endif
}
