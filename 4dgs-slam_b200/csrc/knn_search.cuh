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
//   1. takes the third-best distance among its Morton neighbours j-3 .. j+3 as an upper bound r^2 (the reference's first loop),
//   2. picks the level whose cells are just wider than r, so that the 3x3x3 block of cells around it contains the ball of radius r,
//   3. visits the block (own cell first; every cell pruned by its box distance against min(r^2, third best so far)); a cell that
//      holds more than KNN_LEAF points is opened instead of scanned -- a stackless nearest-first walk over its sub-cells, each
//      pruned the same way -- so a sparse query next to a dense region does not pay for the region,
//   4. is done when its third-best distance is no larger than the distance to the nearest face of the block that has cells behind
//      it (always true after one round up to rounding; otherwise it repeats one level up).
// The work per query is a few dozen distance evaluations whatever the density, instead of P/1024 box tests + thousands of
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
#define KNN_LEAF 32            // a cell with more points than this is opened (its children pruned one by one), not scanned
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
// inverse: every third bit of 63 -> 21 bits
KNN_HD uint32_t knn_compact21(uint64_t x) {
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x1fffffull;
    return (uint32_t)x;
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

// finest level at which the cell of sorted position j holds at least 4 points (common prefix with the sorted neighbours)
KNN_HD int knn_window_level(const KnnIndex& ix, int j) {
    int level = -1;
    const uint64_t* c = ix.code;
    for (int a = j - 3; a <= j; ++a) {
        const int b = a + 3;
        if (a < 0 || b >= ix.P) continue;
        const uint64_t x = c[a] ^ c[b];
        const int common = x ? knn_clz64(x) - 1 : 63;          // shared leading bits of the 63-bit codes
        const int l = common / 3;
        if (l > level) level = l;
    }
    return level;                                              // -1: fewer than 4 points in total
}

// level whose cells are at least as wide as the ball of squared radius r2: the 3x3x3 block around the query then contains the
// ball.  KNN_LMAX bounds it from below (dense spots: the cells are opened, not scanned).
KNN_HD int knn_level_for(const KnnGrid& g, float r2) {
    if (!(g.unit > 0.0f) || !(r2 < FLT_MAX)) return 0;
    const float r_units = sqrtf(r2) / g.unit * 1.0001f + 2.0f * KNN_MARGIN + 1.0f;
    if (!(r_units < 2097152.0f)) return 0;
    const uint32_t need = (uint32_t)r_units + 1u;              // cell side in code units, rounded up ...
    int shift = 0;
    while ((1u << shift) < need) ++shift;                      // ... to a power of two
    const int level = KNN_BITS - shift;
    return level > KNN_LMAX ? KNN_LMAX : (level < 0 ? 0 : level);
}

KNN_HD float knn_dist2(const KnnPoint& p, const KnnPoint& q) {
    const float dx = knn_sub(p.x, q.x), dy = knn_sub(p.y, q.y), dz = knn_sub(p.z, q.z);
    return knn_fma(dz, dz, knn_fma(dx, dx, knn_mul(dy, dy)));          // nvcc's contraction of simple_knn.cu:134 (FMUL on y, FFMA x, FFMA z)
}

struct KnnStats { uint32_t evals, cells, rounds, nodes; };     // host-side instrumentation (tests); the device passes nullptr

struct KnnQuery {
    KnnPoint q;
    float u[3];              // position in code units
    float best[3];
    float reject;            // third-best distance among the Morton neighbours j-3 .. j+3: an upper bound of the answer (simple_knn.cu:153-161)
    float unit2;
};

// squared distance (code units, shrunk by the rounding margin per axis) from the query to the cell `key` of `level`
KNN_HD float knn_box_dist2(const KnnQuery& Q, int level, uint64_t key) {
    const float S = (float)(1u << (KNN_BITS - level));
    float bd2 = 0.0f;
    for (int a = 0; a < 3; ++a) {
        const float lo = (float)knn_compact21(key >> a) * S;
        float d = fmaxf(lo - Q.u[a], Q.u[a] - (lo + S));
        d = fmaxf(d - KNN_MARGIN, 0.0f);
        bd2 += d * d;
    }
    return bd2;
}

KNN_HD void knn_scan(const KnnIndex& ix, KnnQuery& Q, uint32_t lo, uint32_t hi, KnnStats* st) {
    for (uint32_t i = lo; i < hi; ++i) {
        const KnnPoint p = ix.pts[i];
        if (p.idx == Q.q.idx) continue;                         // simple_knn.cu:156,174: only the query itself is skipped
        knn_insert(Q.best, knn_dist2(p, Q.q));
        if (st) st->evals++;
        if (Q.best[2] == 0.0f) return;                          // three coincident points: nothing can improve
    }
}

// child of the cell (level, key) that is nearest to the query: per axis, the half the query's coordinate falls into (or faces)
KNN_HD uint32_t knn_octant(const KnnQuery& Q, int level, uint64_t key) {
    const float Sc = (float)(1u << (KNN_BITS - level - 1));     // child side
    uint32_t oct = 0;
    for (int a = 0; a < 3; ++a) {
        const float centre = (float)(2u * knn_compact21(key >> a) + 1u) * Sc;
        if (Q.u[a] >= centre) oct |= 1u << a;
    }
    return oct;
}

// Everything in the cell (level0, key0) = sorted range [rlo, rhi) that can still improve the answer.  A cell with more than
// KNN_LEAF points is not scanned but opened: its 8 children are visited nearest first (child index v ^ octant of the query,
// v = 0 .. 7), each pruned by its box distance, so the bound tightens before the far children are looked at.  The traversal needs
// no stack: `vkey` holds the visit counters v of the path (3 bits per level) next to the real Morton key, the next node is the
// next counter value at this level or, when it wraps, at the parent's.  A sub-cell's range comes from the prefix table
// (levels <= T) or a binary search inside its T-level ancestor (or inside the root, when the root is deeper than T).
KNN_HD void knn_visit(const KnnIndex& ix, KnnQuery& Q, int level0, uint64_t key0, uint32_t rlo, uint32_t rhi, KnnStats* st) {
    int level = level0;
    uint64_t key = key0, vkey = 0;
    for (;;) {
        bool descend = false;
        if (st) st->nodes++;
        if (level == level0 || !(knn_box_dist2(Q, level, key) * Q.unit2 * 0.99999f > fminf(Q.best[2], Q.reject))) {
            uint32_t lo = rlo, hi = rhi;
            if (level != level0) {
                if (level <= ix.T || level0 < ix.T) knn_range(ix, level, key, &lo, &hi);     // table, or search inside the T-level ancestor
                else {                                                                        // the root is the tighter bracket
                    const int sh = 63 - 3 * level;
                    lo = knn_lower_bound(ix.code, rlo, rhi, key << sh);
                    hi = knn_lower_bound(ix.code, lo, rhi, (key + 1) << sh);
                }
            }
            if (hi - lo > KNN_LEAF && level < KNN_BITS) descend = true;
            else knn_scan(ix, Q, lo, hi, st);
        }
        if (Q.best[2] == 0.0f) return;
        if (descend) {
            key = (key << 3) | knn_octant(Q, level, key);
            vkey <<= 3;
            ++level;
            continue;
        }
        for (;;) {                                              // next node: nearest-first order among siblings, then up
            if (level == level0) return;
            const uint32_t v = (uint32_t)(vkey & 7u) + 1u;
            if (v < 8u) {
                const uint64_t parent = key >> 3;
                key = (parent << 3) | (v ^ knn_octant(Q, level - 1, parent));
                vkey = (vkey & ~7ull) | v;
                break;
            }
            key >>= 3;
            vkey >>= 3;
            --level;
        }
    }
}

KNN_HD float knn_query(const KnnIndex& ix, const KnnGrid& g, int j, KnnStats* st) {
    KnnQuery Q;
    Q.q = ix.pts[j];
    uint32_t c0[3];
    knn_cell(g, Q.q.x, Q.q.y, Q.q.z, Q.u, c0);
    Q.unit2 = g.unit * g.unit;
    float start2;            // squared radius the first round is sized for
    {   // upper bound from the neighbours in Morton order, like the reference's first loop
        float r[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        const int a = j - 3 < 0 ? 0 : j - 3, b = j + 3 > ix.P - 1 ? ix.P - 1 : j + 3;
        for (int i = a; i <= b; ++i) {
            if (i != j) knn_insert(r, knn_dist2(ix.pts[i], Q.q));
        }
        Q.reject = r[2];
        // second bound: 4 points share the query's cell at the window level, so three of them are within its diagonal
        const int wl = knn_window_level(ix, j);
        if (wl >= 0) {
            const float side = ((float)(1u << (KNN_BITS - wl)) + KNN_MARGIN) * g.unit;
            Q.reject = fminf(Q.reject, 3.0f * side * side * 1.00001f);
        }
        // Where the Morton curve jumps (cell corners), the third neighbour along the curve can be far away although the real
        // neighbours sit just across the cell face: then both bounds are useless (measured: 3.5 k cells opened for one query).
        // The nearest curve neighbour is usually still a real neighbour, so the first round is sized for 4x its distance; a
        // round that does not settle the query tightens the bound and the next one is sized from it.
        const float d1 = r[0] > 0.0f ? r[0] : (r[1] > 0.0f ? r[1] : r[2]);
        start2 = fminf(Q.reject, 16.0f * d1);
    }
    int level = knn_level_for(g, start2);
    for (;;) {
        const int shift = KNN_BITS - level;
        const float S = (float)(1u << shift);                  // cell side in code units
        const int G = 1 << level;
        int ck[3];
        float m = FLT_MAX;                                     // distance to the nearest block face that has cells behind it
        // per axis and offset (0, -1, +1): squared box distance in code units (shrunk by the rounding margin), the axis' share of
        // the Morton key, and whether the cell exists -- a block cell then costs three selects, not three bit de-interleaves
        float ad2[3][3];
        uint64_t ek[3][3];
        bool in[3][3];
        for (int a = 0; a < 3; ++a) {
            ck[a] = (int)(c0[a] >> shift);
            const float lo_d = Q.u[a] - (float)ck[a] * S, hi_d = (float)(ck[a] + 1) * S - Q.u[a];
            if (ck[a] >= 2) m = fminf(m, lo_d + S);
            if (ck[a] + 2 <= G - 1) m = fminf(m, hi_d + S);
            const float dl = fmaxf(lo_d - KNN_MARGIN, 0.0f), dh = fmaxf(hi_d - KNN_MARGIN, 0.0f);
            ad2[a][0] = 0.0f; ad2[a][1] = dl * dl; ad2[a][2] = dh * dh;
            in[a][0] = true; in[a][1] = ck[a] >= 1; in[a][2] = ck[a] + 1 <= G - 1;
            ek[a][0] = knn_expand21((uint32_t)ck[a]) << a;
            ek[a][1] = knn_expand21((uint32_t)(ck[a] - 1) & 0x1fffffu) << a;
            ek[a][2] = knn_expand21((uint32_t)(ck[a] + 1) & 0x1fffffu) << a;
        }
        Q.best[0] = Q.best[1] = Q.best[2] = FLT_MAX;
        if (st) st->rounds++;
        for (int t = 0; t < 27; ++t) {                          // offsets in the order 0, -1, +1 per axis: own cell first
            const int ix0 = t % 3, iy0 = (t / 3) % 3, iz0 = t / 9;
#define KNN_PICK(arr, a, i) ((i) == 0 ? arr[a][0] : ((i) == 1 ? arr[a][1] : arr[a][2]))
            if (!(KNN_PICK(in, 0, ix0) && KNN_PICK(in, 1, iy0) && KNN_PICK(in, 2, iz0))) continue;
            const float bd2 = (KNN_PICK(ad2, 0, ix0) + KNN_PICK(ad2, 1, iy0)) + KNN_PICK(ad2, 2, iz0);
            if (t > 0 && bd2 * Q.unit2 * 0.99999f > fminf(Q.best[2], Q.reject)) continue;
            const uint64_t key = KNN_PICK(ek, 0, ix0) | KNN_PICK(ek, 1, iy0) | KNN_PICK(ek, 2, iz0);
#undef KNN_PICK
            uint32_t lo, hi;
            knn_range(ix, level, key, &lo, &hi);
            if (st) st->cells++;
            if (lo == hi) continue;
            knn_visit(ix, Q, level, key, lo, hi, st);
            if (Q.best[2] == 0.0f) break;
        }
        if (m == FLT_MAX || level == 0 || Q.best[2] == 0.0f) break;   // the block covered everything there is
        const float lb = (m - KNN_MARGIN) * g.unit * 0.999999f;       // every point outside the block is at least this far
        if (lb > 0.0f && Q.best[2] <= lb * lb) break;
        if (Q.best[2] < FLT_MAX) {                              // three real points seen: a valid (and now tight) bound
            Q.reject = fminf(Q.reject, Q.best[2]);
            const int l2 = knn_level_for(g, Q.reject);
            level = l2 < level - 1 ? l2 : level - 1;
        } else {
            level = level > 2 ? level - 2 : 0;
        }
    }
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(__fadd_rn(__fadd_rn(Q.best[0], Q.best[1]), Q.best[2]), 3.0f);
#else
    return ((Q.best[0] + Q.best[1]) + Q.best[2]) / 3.0f;
#endif
}
