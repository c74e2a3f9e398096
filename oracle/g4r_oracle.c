/*
 * g4r_oracle.c -- CPU restatement of the reference rasterizer.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's
 * shared objects.  The product (4dgs-slam_b200/) never links, imports or calls it.
 *
 * It restates, stage by stage, /root/reference/submodules/diff-gaussian-rasterization ("DGR/"):
 *   oracle_project      <- preprocessCUDA fwd      DGR/cuda_rasterizer/forward.cu:157-258
 *                          in_frustum               DGR/cuda_rasterizer/auxiliary.h:139-164
 *                          computeCov3D/Cov2D/SH    DGR/cuda_rasterizer/forward.cu:22-154
 *                          ndc2Pix/getRect          DGR/cuda_rasterizer/auxiliary.h:41-56
 *   oracle_bin          <- InclusiveSum, duplicateWithKeys, SortPairs, identifyTileRanges
 *                                                   DGR/cuda_rasterizer/rasterizer_impl.cu:70-138,280-321
 *   oracle_composite    <- renderCUDA fwd           DGR/cuda_rasterizer/forward.cu:263-392
 *   oracle_composite_bw <- renderCUDA bwd           DGR/cuda_rasterizer/backward.cu:563-787
 *   oracle_gaussian_bw  <- computeCov2DCUDA         DGR/cuda_rasterizer/backward.cu:150-346
 *                          preprocessCUDA bwd        DGR/cuda_rasterizer/backward.cu:418-539
 *                          SH / cov3D backward       DGR/cuda_rasterizer/backward.cu:21-145,350-413
 *                          sum of dL_dtau over P     DGR/diff_gaussian_rasterization/__init__.py:152-154
 *
 * Parity status: PINNED against outputs of the unmodified reference build run on a B200
 * (the .npz files under tests/golden, generated on a B200 by tools/gpu_check.py --golden; see tests/test_oracle_golden.py).
 * The reference itself ships no tests or golden vectors.
 *
 * Built twice (oracle/Makefile): REAL=float -> libg4r_oracle_f32.so (bit-faithful on the integer
 * path: radii, tile rectangles, depth keys -- every operation there uses the rounding sequence the
 * reference executes on sm_100a, via fmaf() and -ffp-contract=off), and REAL=double ->
 * libg4r_oracle_f64.so (used for finite-difference checks of the analytic backward).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORACLE_F64
typedef float REAL;
#define R_FMA fmaf
#define R_SQRT sqrtf
#define R_EXP expf
#define R_CEIL ceilf
#define R_FMIN fminf
#define R_FMAX fmaxf
#define ORACLE_NAME(x) x##_f32
#else
typedef double REAL;
#define R_FMA fma
#define R_SQRT sqrt
#define R_EXP exp
#define R_CEIL ceil
#define R_FMIN fmin
#define R_FMAX fmax
#define ORACLE_NAME(x) x##_f64
#endif
#define RC(x) ((REAL)(x))

#define TILE 16

typedef struct OracleScene {
    int32_t P, D, M, W, H;
    REAL tan_fovx, tan_fovy, scale_modifier;
    const REAL *bg, *viewmatrix, *projmatrix, *projmatrix_raw, *campos;
    const REAL *means3D, *opacities, *shs, *colors_precomp, *scales, *rotations, *cov3D_precomp;
} OracleScene;

/* per-Gaussian forward state (the reference's GeometryState, rasterizer_impl.h:29-45) */
typedef struct OracleGeom {
    REAL* depths;          /* [P]   */
    REAL* means2D;         /* [P,2] */
    REAL* cov3D;           /* [P,6] */
    REAL* conic_opacity;   /* [P,4] */
    REAL* rgb;             /* [P,3] */
    uint8_t* clamped;      /* [P,3] */
    int32_t* radii;        /* [P]   */
    uint32_t* tiles_touched; /* [P] */
} OracleGeom;

static const REAL SH_C0 = RC(0.28209479177387814);
static const REAL SH_C1 = RC(0.4886025119029199);
static const REAL SH_C2[5] = {RC(1.0925484305920792), RC(-1.0925484305920792), RC(0.31539156525252005), RC(-1.0925484305920792),
                              RC(0.5462742152960396)};
static const REAL SH_C3[7] = {RC(-0.5900435899266435), RC(2.890611442640554), RC(-0.4570457994644658), RC(0.3731763325901154),
                              RC(-0.4570457994644658), RC(1.445305721320277), RC(-0.5900435899266435)};

/* float -> int32 like the GPU's cvt.rzi.s32.f32: truncate, saturate, NaN -> 0 */
static int32_t f2i_sat(REAL v) {
    if (v != v) return 0;
    if (v >= RC(2147483648.0)) return INT32_MAX;
    if (v <= RC(-2147483648.0)) return INT32_MIN;
    return (int32_t)v;
}
static uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
static int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

/* getRect (auxiliary.h:46-56): float arithmetic, one rounding per operation */
static void get_rect(REAL px, REAL py, int32_t radius, uint32_t gx, uint32_t gy, uint32_t* x0, uint32_t* y0, uint32_t* x1, uint32_t* y1) {
    const REAL rf = (REAL)radius;
    *x0 = umin(gx, (uint32_t)imax(0, f2i_sat((px - rf) * RC(0.0625))));
    *y0 = umin(gy, (uint32_t)imax(0, f2i_sat((py - rf) * RC(0.0625))));
    *x1 = umin(gx, (uint32_t)imax(0, f2i_sat((((px + rf) + RC(16.0)) + RC(-1.0)) * RC(0.0625))));
    *y1 = umin(gy, (uint32_t)imax(0, f2i_sat((((py + rf) + RC(16.0)) + RC(-1.0)) * RC(0.0625))));
}

/* one row of transformPoint4x3 / 4x4 (auxiliary.h:58-78) as the reference build evaluates it */
static REAL row_affine(REAL a, REAL b, REAL c, REAL d, REAL x, REAL y, REAL z) {
    REAL t = y * b;
    t = R_FMA(x, a, t);
    t = R_FMA(z, c, t);
    return d + t;
}
/* p*q + r*s + u*v with the middle product rounded first (every 3-term glm product) */
static REAL dot3m(REAL p, REAL q, REAL r, REAL s, REAL u, REAL v) {
    REAL t = r * s;
    t = R_FMA(p, q, t);
    return R_FMA(u, v, t);
}

/* ------------------------------------------------------------------------------------------- */
/* projection                                                                                   */
/* ------------------------------------------------------------------------------------------- */
int ORACLE_NAME(oracle_project)(const OracleScene* sc, OracleGeom* g) {
    const REAL* V = sc->viewmatrix;
    const REAL* Q = sc->projmatrix;
    const uint32_t gx = (uint32_t)((sc->W + TILE - 1) / TILE), gy = (uint32_t)((sc->H + TILE - 1) / TILE);
    const REAL focal_y = (REAL)sc->H / (RC(2.0) * sc->tan_fovy);
    const REAL focal_x = (REAL)sc->W / (RC(2.0) * sc->tan_fovx);
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < sc->P; ++i) {
        g->radii[i] = 0;
        g->tiles_touched[i] = 0;
        const REAL x = sc->means3D[3 * i], y = sc->means3D[3 * i + 1], z = sc->means3D[3 * i + 2];
        const REAL depth = row_affine(V[2], V[6], V[10], V[14], x, y, z);
        if (depth <= RC(0.2)) continue;
        const REAL hx = row_affine(Q[0], Q[4], Q[8], Q[12], x, y, z);
        const REAL hy = row_affine(Q[1], Q[5], Q[9], Q[13], x, y, z);
        const REAL hw = row_affine(Q[3], Q[7], Q[11], Q[15], x, y, z);
        const REAL pw = RC(1.0) / (hw + RC(0.0000001));
        const REAL ndc_x = hx * pw, ndc_y = hy * pw;

        REAL c[6];
        if (sc->cov3D_precomp) {
            for (int k = 0; k < 6; ++k) c[k] = sc->cov3D_precomp[6 * i + k];
        } else {
            const REAL sx = sc->scale_modifier * sc->scales[3 * i], sy = sc->scale_modifier * sc->scales[3 * i + 1],
                       sz = sc->scale_modifier * sc->scales[3 * i + 2];
            const REAL qr = sc->rotations[4 * i], qx = sc->rotations[4 * i + 1], qy = sc->rotations[4 * i + 2], qz = sc->rotations[4 * i + 3];
            const REAL yy = qy * qy, zz = qz * qz, xz = qx * qz, rz = qr * qz, rx = qr * qx;
            const REAL yy_zz = yy + zz, xx_zz = R_FMA(qx, qx, zz), xx_yy = R_FMA(qx, qx, yy);
            const REAL xy_m_rz = R_FMA(qx, qy, -rz), xy_p_rz = R_FMA(qx, qy, rz);
            const REAL xz_p_ry = R_FMA(qr, qy, xz), xz_m_ry = R_FMA(-qr, qy, xz);
            const REAL yz_m_rx = R_FMA(qy, qz, -rx), yz_p_rx = R_FMA(qy, qz, rx);
            const REAL m00 = sx * (RC(1.0) - (yy_zz + yy_zz)), m01 = sy * (xy_m_rz + xy_m_rz), m02 = sz * (xz_p_ry + xz_p_ry);
            const REAL m10 = sx * (xy_p_rz + xy_p_rz), m11 = sy * (RC(1.0) - (xx_zz + xx_zz)), m12 = sz * (yz_m_rx + yz_m_rx);
            const REAL m20 = sx * (xz_m_ry + xz_m_ry), m21 = sy * (yz_p_rx + yz_p_rx), m22 = sz * (RC(1.0) - (xx_yy + xx_yy));
            c[0] = dot3m(m00, m00, m01, m01, m02, m02);
            c[1] = dot3m(m10, m00, m11, m01, m12, m02);
            c[2] = dot3m(m20, m00, m21, m01, m22, m02);
            c[3] = dot3m(m10, m10, m11, m11, m12, m12);
            c[4] = dot3m(m20, m10, m21, m11, m22, m12);
            c[5] = dot3m(m20, m20, m21, m21, m22, m22);
            for (int k = 0; k < 6; ++k) g->cov3D[6 * i + k] = c[k];
        }

        const REAL tx = row_affine(V[0], V[4], V[8], V[12], x, y, z);
        const REAL ty = row_affine(V[1], V[5], V[9], V[13], x, y, z);
        const REAL tz = depth;
        const REAL limx = sc->tan_fovx * RC(1.3), limy = sc->tan_fovy * RC(1.3);
        const REAL cx = R_FMIN(limx, R_FMAX(-limx, tx / tz));
        const REAL cy = R_FMIN(limy, R_FMAX(-limy, ty / tz));
        const REAL tz2 = tz * tz;
        const REAL j00 = focal_x / tz, j02 = (focal_x * (cx * -tz)) / tz2;
        const REAL j11 = focal_y / tz, j12 = (focal_y * (cy * -tz)) / tz2;
        const REAL u0 = R_FMA(V[2], j02, V[0] * j00), u1 = R_FMA(V[6], j02, V[4] * j00), u2 = R_FMA(V[10], j02, V[8] * j00);
        const REAL w0 = R_FMA(V[2], j12, V[1] * j11), w1 = R_FMA(V[6], j12, V[5] * j11), w2 = R_FMA(V[10], j12, V[9] * j11);
        const REAL au0 = dot3m(u0, c[0], u1, c[1], u2, c[2]), au1 = dot3m(u0, c[1], u1, c[3], u2, c[4]), au2 = dot3m(u0, c[2], u1, c[4], u2, c[5]);
        const REAL aw0 = dot3m(w0, c[0], w1, c[1], w2, c[2]), aw1 = dot3m(w0, c[1], w1, c[3], w2, c[4]), aw2 = dot3m(w0, c[2], w1, c[4], w2, c[5]);
        const REAL cov_a = dot3m(u0, au0, u1, au1, u2, au2) + RC(0.3);
        const REAL cov_b = dot3m(u0, aw0, u1, aw1, u2, aw2);
        const REAL cov_c = dot3m(w0, aw0, w1, aw1, w2, aw2) + RC(0.3);

        const REAL det = R_FMA(cov_a, cov_c, -(cov_b * cov_b));
        if (det == RC(0.0)) continue;
        const REAL det_inv = RC(1.0) / det;
        const REAL con_x = cov_c * det_inv, con_y = det_inv * -cov_b, con_z = cov_a * det_inv;
        const REAL mid = (cov_a + cov_c) * RC(0.5);
        const REAL disc = R_SQRT(R_FMAX(R_FMA(mid, mid, -det), RC(0.1)));
        const REAL lam = R_FMAX(mid + disc, mid - disc);
        const REAL radius_f = R_CEIL(R_SQRT(lam) * RC(3.0));
        /* ndc2Pix: double arithmetic with a fused multiply-add (auxiliary.h:41-44 as compiled) */
        const REAL px = (REAL)(fma((double)ndc_x + 1.0, (double)sc->W, -1.0) * 0.5);
        const REAL py = (REAL)(fma((double)ndc_y + 1.0, (double)sc->H, -1.0) * 0.5);
        const int32_t radius = f2i_sat(radius_f);
        uint32_t x0, y0, x1, y1;
        get_rect(px, py, radius, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0u) continue;

        if (!sc->colors_precomp) {
            REAL dx = x - sc->campos[0], dy = y - sc->campos[1], dz = z - sc->campos[2];
            const REAL len = R_SQRT(dx * dx + dy * dy + dz * dz);
            dx /= len; dy /= len; dz /= len;
            const REAL* sh = sc->shs + (size_t)i * sc->M * 3;
            for (int ch = 0; ch < 3; ++ch) {
#define SH(k) sh[(k) * 3 + ch]
                REAL r = SH_C0 * SH(0);
                if (sc->D > 0) {
                    r = r - SH_C1 * dy * SH(1) + SH_C1 * dz * SH(2) - SH_C1 * dx * SH(3);
                    if (sc->D > 1) {
                        const REAL xx = dx * dx, yy = dy * dy, zz = dz * dz, xy = dx * dy, yz = dy * dz, xz = dx * dz;
                        r = r + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) + SH_C2[2] * (RC(2.0) * zz - xx - yy) * SH(6) +
                            SH_C2[3] * xz * SH(7) + SH_C2[4] * (xx - yy) * SH(8);
                        if (sc->D > 2) {
                            r = r + SH_C3[0] * dy * (RC(3.0) * xx - yy) * SH(9) + SH_C3[1] * xy * dz * SH(10) +
                                SH_C3[2] * dy * (RC(4.0) * zz - xx - yy) * SH(11) +
                                SH_C3[3] * dz * (RC(2.0) * zz - RC(3.0) * xx - RC(3.0) * yy) * SH(12) +
                                SH_C3[4] * dx * (RC(4.0) * zz - xx - yy) * SH(13) + SH_C3[5] * dz * (xx - yy) * SH(14) +
                                SH_C3[6] * dx * (xx - RC(3.0) * yy) * SH(15);
                        }
                    }
                }
#undef SH
                r += RC(0.5);
                g->clamped[3 * i + ch] = (uint8_t)(r < RC(0.0));
                g->rgb[3 * i + ch] = r < RC(0.0) ? RC(0.0) : r;
            }
        }
        g->depths[i] = depth;
        g->radii[i] = radius;
        g->means2D[2 * i] = px;
        g->means2D[2 * i + 1] = py;
        g->conic_opacity[4 * i] = con_x;
        g->conic_opacity[4 * i + 1] = con_y;
        g->conic_opacity[4 * i + 2] = con_z;
        g->conic_opacity[4 * i + 3] = sc->opacities[i];
        g->tiles_touched[i] = (y1 - y0) * (x1 - x0);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* binning: keys (tile, depth bits) sorted stably over instances emitted in Gaussian order       */
/* ------------------------------------------------------------------------------------------- */
typedef struct { uint32_t tile; REAL depth; uint32_t id; } Inst;
static int inst_cmp(const void* a, const void* b) {
    const Inst* p = (const Inst*)a;
    const Inst* q = (const Inst*)b;
    if (p->tile != q->tile) return p->tile < q->tile ? -1 : 1;
    /* depth > 0.2 here, so IEEE bit order == numeric order; stable radix sort => ties by id */
    if (p->depth != q->depth) return p->depth < q->depth ? -1 : 1;
    if (p->id != q->id) return p->id < q->id ? -1 : 1;
    return 0;
}

/* returns N; *point_list is malloc'ed (free with oracle_free); ranges is [tiles,2] */
int64_t ORACLE_NAME(oracle_bin)(const OracleScene* sc, const OracleGeom* g, uint32_t** point_list, uint32_t* ranges) {
    const uint32_t gx = (uint32_t)((sc->W + TILE - 1) / TILE), gy = (uint32_t)((sc->H + TILE - 1) / TILE);
    int64_t N = 0;
    for (int32_t i = 0; i < sc->P; ++i) N += g->tiles_touched[i];
    Inst* inst = (Inst*)malloc(sizeof(Inst) * (size_t)(N > 0 ? N : 1));
    int64_t off = 0;
    for (int32_t i = 0; i < sc->P; ++i) {
        if (g->radii[i] <= 0) continue;
        uint32_t x0, y0, x1, y1;
        get_rect(g->means2D[2 * i], g->means2D[2 * i + 1], g->radii[i], gx, gy, &x0, &y0, &x1, &y1);
        for (uint32_t ty = y0; ty < y1; ++ty)
            for (uint32_t tx = x0; tx < x1; ++tx) {
                inst[off].tile = ty * gx + tx;
                inst[off].depth = g->depths[i];
                inst[off].id = (uint32_t)i;
                ++off;
            }
    }
    qsort(inst, (size_t)N, sizeof(Inst), inst_cmp);
    memset(ranges, 0, sizeof(uint32_t) * 2 * gx * gy);
    uint32_t* pl = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N > 0 ? N : 1));
    for (int64_t k = 0; k < N; ++k) {
        pl[k] = inst[k].id;
        const uint32_t t = inst[k].tile;
        if (k == 0) ranges[2 * t] = 0;
        else if (inst[k - 1].tile != t) { ranges[2 * inst[k - 1].tile + 1] = (uint32_t)k; ranges[2 * t] = (uint32_t)k; }
        if (k == N - 1) ranges[2 * t + 1] = (uint32_t)N;
    }
    free(inst);
    *point_list = pl;
    return N;
}
void ORACLE_NAME(oracle_free)(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------- */
/* forward composite                                                                             */
/* ------------------------------------------------------------------------------------------- */
static REAL splat_power(REAL dx, REAL dy, REAL A, REAL B, REAL C) {
    /* forward.cu:345 as the reference build evaluates it */
    const REAL q = R_FMA(dx, dx * A, dy * (dy * C));
    return R_FMA(q, RC(-0.5), -(dy * (dx * B)));
}

int ORACLE_NAME(oracle_composite)(const OracleScene* sc, const OracleGeom* g, const uint32_t* point_list, const uint32_t* ranges,
                                  REAL* out_color, REAL* out_depth, REAL* out_opacity, REAL* final_T, uint32_t* n_contrib, int32_t* n_touched) {
    const int W = sc->W, H = sc->H;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const REAL* feat = sc->colors_precomp ? sc->colors_precomp : g->rgb;
    memset(n_touched, 0, sizeof(int32_t) * (size_t)sc->P);
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; ++tile) {
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        const int bx = (tile % gx) * TILE, by = (tile / gx) * TILE;
        for (int py = by; py < by + TILE && py < H; ++py)
            for (int px = bx; px < bx + TILE && px < W; ++px) {
                const REAL pxf = (REAL)px, pyf = (REAL)py;
                REAL T = RC(1.0), C0 = 0, C1 = 0, C2 = 0, Dz = 0;
                uint32_t contributor = 0, last = 0;
                for (uint32_t k = r0; k < r1; ++k) {
                    ++contributor;
                    const uint32_t id = point_list[k];
                    const REAL dx = g->means2D[2 * id] - pxf, dy = g->means2D[2 * id + 1] - pyf;
                    const REAL* co = g->conic_opacity + 4 * (size_t)id;
                    const REAL power = splat_power(dx, dy, co[0], co[1], co[2]);
                    if (power > RC(0.0)) continue;
                    const REAL alpha = R_FMIN(RC(0.99), co[3] * R_EXP(power));
                    if (alpha < RC(1.0) / RC(255.0)) continue;
                    const REAL test_T = T * (RC(1.0) - alpha);
                    if (test_T < RC(0.0001)) break;
                    C0 = R_FMA(T, alpha * feat[3 * id], C0);
                    C1 = R_FMA(T, alpha * feat[3 * id + 1], C1);
                    C2 = R_FMA(T, alpha * feat[3 * id + 2], C2);
                    Dz = R_FMA(T, alpha * g->depths[id], Dz);
                    if (test_T > RC(0.5)) {
#pragma omp atomic
                        n_touched[id] += 1;
                    }
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;
                final_T[pix] = T;
                n_contrib[pix] = last;
                out_color[pix] = R_FMA(T, sc->bg[0], C0);
                out_color[plane + pix] = R_FMA(T, sc->bg[1], C1);
                out_color[2 * plane + pix] = R_FMA(T, sc->bg[2], C2);
                out_depth[pix] = Dz;
                out_opacity[pix] = RC(1.0) - T;
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* backward composite: per-Gaussian sums of the per-pixel partials (accumulated in double)       */
/*   acc[P,10] = {dmean2D.x, dmean2D.y, dconic.x, dconic.y, dconic.w, dopacity, dcol.r, dcol.g, dcol.b, ddepth} */
/* ------------------------------------------------------------------------------------------- */
int ORACLE_NAME(oracle_composite_bw)(const OracleScene* sc, const OracleGeom* g, const uint32_t* point_list, const uint32_t* ranges,
                                     const REAL* final_T, const uint32_t* n_contrib, const REAL* dL_dpix, const REAL* dL_dpix_depth,
                                     double* acc) {
    const int W = sc->W, H = sc->H;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const REAL* feat = sc->colors_precomp ? sc->colors_precomp : g->rgb;
    const REAL ddelx_dx = RC(0.5) * W, ddely_dy = RC(0.5) * H;
    memset(acc, 0, sizeof(double) * 10 * (size_t)sc->P);
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; ++tile) {
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        const int bx = (tile % gx) * TILE, by = (tile / gx) * TILE;
        for (int py = by; py < by + TILE && py < H; ++py)
            for (int px = bx; px < bx + TILE && px < W; ++px) {
                const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;
                const REAL pxf = (REAL)px, pyf = (REAL)py;
                const REAL T_final = final_T[pix];
                REAL T = T_final;
                const uint32_t last_contributor = n_contrib[pix];
                const REAL dp[3] = {dL_dpix[pix], dL_dpix[plane + pix], dL_dpix[2 * plane + pix]};
                const REAL dpd = dL_dpix_depth[pix];
                REAL accum[3] = {0, 0, 0}, accum_d = 0, last_alpha = 0, last_c[3] = {0, 0, 0}, last_d = 0;
                uint32_t contributor = r1 - r0;
                for (uint32_t k = r1; k-- > r0;) {
                    --contributor;
                    if (contributor >= last_contributor) continue;
                    const uint32_t id = point_list[k];
                    const REAL dx = g->means2D[2 * id] - pxf, dy = g->means2D[2 * id + 1] - pyf;
                    const REAL* co = g->conic_opacity + 4 * (size_t)id;
                    const REAL power = splat_power(dx, dy, co[0], co[1], co[2]);
                    if (power > RC(0.0)) continue;
                    const REAL G = R_EXP(power);
                    const REAL alpha = R_FMIN(RC(0.99), co[3] * G);
                    if (alpha < RC(1.0) / RC(255.0)) continue;
                    T = T / (RC(1.0) - alpha);
                    const REAL w = alpha * T;
                    REAL dL_dalpha = 0;
                    REAL dcol[3];
                    for (int ch = 0; ch < 3; ++ch) {
                        const REAL c = feat[3 * id + ch];
                        accum[ch] = last_alpha * last_c[ch] + (RC(1.0) - last_alpha) * accum[ch];
                        last_c[ch] = c;
                        dL_dalpha += (c - accum[ch]) * dp[ch];
                        dcol[ch] = w * dp[ch];
                    }
                    const REAL dep = g->depths[id];
                    accum_d = last_alpha * last_d + (RC(1.0) - last_alpha) * accum_d;
                    last_d = dep;
                    dL_dalpha += (dep - accum_d) * dpd;
                    const REAL ddep = w * dpd;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    REAL bg_dot = 0;
                    for (int ch = 0; ch < 3; ++ch) bg_dot += sc->bg[ch] * dp[ch];
                    dL_dalpha += (-T_final / (RC(1.0) - alpha)) * bg_dot;
                    const REAL dL_dG = co[3] * dL_dalpha;
                    const REAL gdx = G * dx, gdy = G * dy;
                    const REAL dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const REAL dG_ddely = -gdy * co[2] - gdx * co[1];
                    const double v[10] = {dL_dG * dG_ddelx * ddelx_dx, dL_dG * dG_ddely * ddely_dy, RC(-0.5) * gdx * dx * dL_dG,
                                          RC(-0.5) * gdx * dy * dL_dG, RC(-0.5) * gdy * dy * dL_dG, G * dL_dalpha,
                                          dcol[0], dcol[1], dcol[2], ddep};
                    double* a = acc + 10 * (size_t)id;
                    for (int q = 0; q < 10; ++q) {
#pragma omp atomic
                        a[q] += v[q];
                    }
                }
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* per-Gaussian backward                                                                         */
/* ------------------------------------------------------------------------------------------- */
typedef struct OracleGrads {
    REAL* dL_dmeans3D;   /* [P,3] */
    REAL* dL_dmeans2D;   /* [P,3] */
    REAL* dL_dopacity;   /* [P]   */
    REAL* dL_dcolors;    /* [P,3] gradient w.r.t. the per-Gaussian RGB (colors_precomp or SH output) */
    REAL* dL_dcov3D;     /* [P,6] */
    REAL* dL_dshs;       /* [P,M,3] or NULL */
    REAL* dL_dscales;    /* [P,3] or NULL */
    REAL* dL_drots;      /* [P,4] or NULL */
    REAL* dL_dtau_rows;  /* [P,6] per-Gaussian pose gradient (the reference's dL_dtau tensor) */
    REAL* dL_dtau;       /* [6]   sum over Gaussians */
} OracleGrads;

int ORACLE_NAME(oracle_gaussian_bw)(const OracleScene* sc, const OracleGeom* g, const double* acc, OracleGrads* o) {
    const REAL* V = sc->viewmatrix;
    const REAL* Pm = sc->projmatrix;
    const REAL* Praw = sc->projmatrix_raw;
    const int P = sc->P, M = sc->M;
    const REAL h_x = (REAL)sc->W / (RC(2.0) * sc->tan_fovx), h_y = (REAL)sc->H / (RC(2.0) * sc->tan_fovy);
    memset(o->dL_dmeans3D, 0, sizeof(REAL) * 3 * P);
    memset(o->dL_dmeans2D, 0, sizeof(REAL) * 3 * P);
    memset(o->dL_dopacity, 0, sizeof(REAL) * P);
    memset(o->dL_dcolors, 0, sizeof(REAL) * 3 * P);
    memset(o->dL_dcov3D, 0, sizeof(REAL) * 6 * P);
    memset(o->dL_dtau_rows, 0, sizeof(REAL) * 6 * P);
    if (o->dL_dshs) memset(o->dL_dshs, 0, sizeof(REAL) * 3 * (size_t)M * P);
    if (o->dL_dscales) memset(o->dL_dscales, 0, sizeof(REAL) * 3 * P);
    if (o->dL_drots) memset(o->dL_drots, 0, sizeof(REAL) * 4 * P);

#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        /* screen-space gradients exist for every Gaussian (zeros when untouched) */
        const double* a = acc + 10 * (size_t)i;
        o->dL_dmeans2D[3 * i] = (REAL)a[0];
        o->dL_dmeans2D[3 * i + 1] = (REAL)a[1];
        o->dL_dopacity[i] = (REAL)a[5];
        o->dL_dcolors[3 * i] = (REAL)a[6]; o->dL_dcolors[3 * i + 1] = (REAL)a[7]; o->dL_dcolors[3 * i + 2] = (REAL)a[8];
        if (!(g->radii[i] > 0)) continue;
        REAL* tau = o->dL_dtau_rows + 6 * (size_t)i;
        const REAL dconic[3] = {(REAL)a[2], (REAL)a[3], (REAL)a[4]};
        const REAL ddepth = (REAL)a[9];
        const REAL mx = sc->means3D[3 * i], my = sc->means3D[3 * i + 1], mz = sc->means3D[3 * i + 2];
        const REAL* cov3D = sc->cov3D_precomp ? sc->cov3D_precomp + 6 * (size_t)i : g->cov3D + 6 * (size_t)i;

        /* ---- computeCov2DCUDA (backward.cu:150-346) ---- */
        REAL t[3] = {V[0] * mx + V[4] * my + V[8] * mz + V[12], V[1] * mx + V[5] * my + V[9] * mz + V[13], V[2] * mx + V[6] * my + V[10] * mz + V[14]};
        const REAL limx = RC(1.3) * sc->tan_fovx, limy = RC(1.3) * sc->tan_fovy;
        const REAL txtz = t[0] / t[2], tytz = t[1] / t[2];
        t[0] = R_FMIN(limx, R_FMAX(-limx, txtz)) * t[2];
        t[1] = R_FMIN(limy, R_FMAX(-limy, tytz)) * t[2];
        const REAL x_grad_mul = (txtz < -limx || txtz > limx) ? 0 : 1;
        const REAL y_grad_mul = (tytz < -limy || tytz > limy) ? 0 : 1;
        /* J (2x3 non-zero part), Wr = rotation of the view matrix with Wr[r][c] = V[r + 4c] */
        const REAL J00 = h_x / t[2], J02 = -(h_x * t[0]) / (t[2] * t[2]), J11 = h_y / t[2], J12 = -(h_y * t[1]) / (t[2] * t[2]);
        REAL Wr[3][3];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Wr[r][c] = V[r + 4 * c];
        /* Tm = J * Wr  (2x3): Tm[0][k] = J00*Wr[0][k] + J02*Wr[2][k];  Tm[1][k] = J11*Wr[1][k] + J12*Wr[2][k] */
        REAL Tm[2][3];
        for (int k = 0; k < 3; ++k) { Tm[0][k] = J00 * Wr[0][k] + J02 * Wr[2][k]; Tm[1][k] = J11 * Wr[1][k] + J12 * Wr[2][k]; }
        const REAL S[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
        REAL ST[2][3];   /* ST[r] = S * Tm[r] */
        for (int r = 0; r < 2; ++r) for (int k = 0; k < 3; ++k) ST[r][k] = S[k][0] * Tm[r][0] + S[k][1] * Tm[r][1] + S[k][2] * Tm[r][2];
        const REAL ca = Tm[0][0] * ST[0][0] + Tm[0][1] * ST[0][1] + Tm[0][2] * ST[0][2] + RC(0.3);
        const REAL cb = Tm[0][0] * ST[1][0] + Tm[0][1] * ST[1][1] + Tm[0][2] * ST[1][2];
        const REAL cc = Tm[1][0] * ST[1][0] + Tm[1][1] * ST[1][1] + Tm[1][2] * ST[1][2] + RC(0.3);
        const REAL denom = ca * cc - cb * cb;
        const REAL denom2inv = RC(1.0) / ((denom * denom) + RC(0.0000001));
        REAL dL_da = 0, dL_db = 0, dL_dc = 0;
        REAL* dcov = o->dL_dcov3D + 6 * (size_t)i;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-cc * cc * dconic[0] + 2 * cb * cc * dconic[1] + (denom - ca * cc) * dconic[2]);
            dL_dc = denom2inv * (-ca * ca * dconic[2] + 2 * ca * cb * dconic[1] + (denom - ca * cc) * dconic[0]);
            dL_db = denom2inv * 2 * (cb * cc * dconic[0] - (denom + 2 * cb * cb) * dconic[1] + ca * cb * dconic[2]);
            static const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
            for (int p = 0; p < 3; ++p)
                for (int q = p; q < 3; ++q) {
                    if (p == q) dcov[idx[p][q]] = Tm[0][p] * Tm[0][p] * dL_da + Tm[0][p] * Tm[1][p] * dL_db + Tm[1][p] * Tm[1][p] * dL_dc;
                    else dcov[idx[p][q]] = 2 * Tm[0][p] * Tm[0][q] * dL_da + (Tm[0][p] * Tm[1][q] + Tm[0][q] * Tm[1][p]) * dL_db + 2 * Tm[1][p] * Tm[1][q] * dL_dc;
                }
        }
        REAL dT[2][3];
        for (int k = 0; k < 3; ++k) {
            dT[0][k] = 2 * ST[0][k] * dL_da + ST[1][k] * dL_db;
            dT[1][k] = 2 * ST[1][k] * dL_dc + ST[0][k] * dL_db;
        }
        REAL dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int k = 0; k < 3; ++k) { dJ00 += Wr[0][k] * dT[0][k]; dJ02 += Wr[2][k] * dT[0][k]; dJ11 += Wr[1][k] * dT[1][k]; dJ12 += Wr[2][k] * dT[1][k]; }
        const REAL tz = RC(1.0) / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const REAL dtx = x_grad_mul * -h_x * tz2 * dJ02;
        const REAL dty = y_grad_mul * -h_y * tz2 * dJ12;
        const REAL dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * t[0]) * tz3 * dJ02 + (2 * h_y * t[1]) * tz3 * dJ12;
        /* pose through t: d p_C/d rho = I, d p_C/d theta = -[t]x (clamped t), backward.cu:273-288 */
        tau[0] += dtx; tau[1] += dty; tau[2] += dtz;
        tau[3] += -t[2] * dty + t[1] * dtz;
        tau[4] += t[2] * dtx - t[0] * dtz;
        tau[5] += -t[1] * dtx + t[0] * dty;
        REAL dmean[3] = {V[0] * dtx + V[1] * dty + V[2] * dtz, V[4] * dtx + V[5] * dty + V[6] * dtz, V[8] * dtx + V[9] * dty + V[10] * dtz};
        /* pose through the rotation inside Tm (backward.cu:299-343): dL/dWr[r][k] */
        {
            REAL dWr[3][3];
            for (int k = 0; k < 3; ++k) { dWr[0][k] = J00 * dT[0][k]; dWr[1][k] = J11 * dT[1][k]; dWr[2][k] = J02 * dT[0][k] + J12 * dT[1][k]; }
            /* column k of the W2C rotation is c_k = (V[4k], V[4k+1], V[4k+2]); its gradient vector is g_k = dWr[:,k] */
            for (int k = 0; k < 3; ++k) {
                const REAL c[3] = {V[4 * k], V[4 * k + 1], V[4 * k + 2]};
                const REAL gk[3] = {dWr[0][k], dWr[1][k], dWr[2][k]};
                tau[3] += -gk[1] * c[2] + gk[2] * c[1];
                tau[4] += gk[0] * c[2] - gk[2] * c[0];
                tau[5] += -gk[0] * c[1] + gk[1] * c[0];
            }
        }

        /* ---- preprocessCUDA backward (backward.cu:446-528) ---- */
        const REAL hx = Pm[0] * mx + Pm[4] * my + Pm[8] * mz + Pm[12];
        const REAL hy = Pm[1] * mx + Pm[5] * my + Pm[9] * mz + Pm[13];
        const REAL hw = Pm[3] * mx + Pm[7] * my + Pm[11] * mz + Pm[15];
        const REAL m_w = RC(1.0) / (hw + RC(0.0000001));
        const REAL mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
        const REAL d2x = (REAL)a[0], d2y = (REAL)a[1];
        dmean[0] += (Pm[0] * m_w - Pm[3] * mul1) * d2x + (Pm[1] * m_w - Pm[3] * mul2) * d2y;
        dmean[1] += (Pm[4] * m_w - Pm[7] * mul1) * d2x + (Pm[5] * m_w - Pm[7] * mul2) * d2y;
        dmean[2] += (Pm[8] * m_w - Pm[11] * mul1) * d2x + (Pm[9] * m_w - Pm[11] * mul2) * d2y;
        {
            const REAL al = m_w, be = -hx * m_w * m_w, ga = -hy * m_w * m_w;
            const REAL pa = Praw[0], pb = Praw[5], pe = Praw[11];
            const REAL pC[3] = {V[0] * mx + V[4] * my + V[8] * mz + V[12], V[1] * mx + V[5] * my + V[9] * mz + V[13], V[2] * mx + V[6] * my + V[10] * mz + V[14]};
            const REAL d1[3] = {al * pa, 0, be * pe}, d2[3] = {0, al * pb, ga * pe};
            /* (-[pC]x)^T d = [pC]x d = pC x d */
            const REAL d1t[3] = {pC[1] * d1[2] - pC[2] * d1[1], pC[2] * d1[0] - pC[0] * d1[2], pC[0] * d1[1] - pC[1] * d1[0]};
            const REAL d2t[3] = {pC[1] * d2[2] - pC[2] * d2[1], pC[2] * d2[0] - pC[0] * d2[2], pC[0] * d2[1] - pC[1] * d2[0]};
            for (int k = 0; k < 3; ++k) { tau[k] += d2x * d1[k] + d2y * d2[k]; tau[3 + k] += d2x * d1t[k] + d2y * d2t[k]; }
            dmean[0] += ddepth * V[2]; dmean[1] += ddepth * V[6]; dmean[2] += ddepth * V[10];
            tau[2] += ddepth;
            tau[3] += ddepth * pC[1];
            tau[4] += ddepth * -pC[0];
        }

        /* ---- SH backward (backward.cu:21-145) ---- */
        if (sc->shs) {
            const REAL dRGB[3] = {g->clamped[3 * i] ? 0 : (REAL)a[6], g->clamped[3 * i + 1] ? 0 : (REAL)a[7], g->clamped[3 * i + 2] ? 0 : (REAL)a[8]};
            const REAL dir_o[3] = {mx - sc->campos[0], my - sc->campos[1], mz - sc->campos[2]};
            const REAL len = R_SQRT(dir_o[0] * dir_o[0] + dir_o[1] * dir_o[1] + dir_o[2] * dir_o[2]);
            const REAL x = dir_o[0] / len, y = dir_o[1] / len, z = dir_o[2] / len;
            const REAL* sh = sc->shs + (size_t)i * M * 3;
            REAL* dsh = o->dL_dshs + (size_t)i * M * 3;
            REAL dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
#define DSH(k, f) for (int ch = 0; ch < 3; ++ch) dsh[(k) * 3 + ch] = (f) * dRGB[ch]
#define ADD3(dst, f, k) for (int ch = 0; ch < 3; ++ch) dst[ch] += (f) * sh[(k) * 3 + ch]
            DSH(0, SH_C0);
            if (sc->D > 0) {
                DSH(1, -SH_C1 * y); DSH(2, SH_C1 * z); DSH(3, -SH_C1 * x);
                ADD3(dRGBdx, -SH_C1, 3); ADD3(dRGBdy, -SH_C1, 1); ADD3(dRGBdz, SH_C1, 2);
                if (sc->D > 1) {
                    const REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    DSH(4, SH_C2[0] * xy); DSH(5, SH_C2[1] * yz); DSH(6, SH_C2[2] * (RC(2.0) * zz - xx - yy)); DSH(7, SH_C2[3] * xz);
                    DSH(8, SH_C2[4] * (xx - yy));
                    ADD3(dRGBdx, SH_C2[0] * y, 4); ADD3(dRGBdx, SH_C2[2] * RC(2.0) * -x, 6); ADD3(dRGBdx, SH_C2[3] * z, 7); ADD3(dRGBdx, SH_C2[4] * RC(2.0) * x, 8);
                    ADD3(dRGBdy, SH_C2[0] * x, 4); ADD3(dRGBdy, SH_C2[1] * z, 5); ADD3(dRGBdy, SH_C2[2] * RC(2.0) * -y, 6); ADD3(dRGBdy, SH_C2[4] * RC(2.0) * -y, 8);
                    ADD3(dRGBdz, SH_C2[1] * y, 5); ADD3(dRGBdz, SH_C2[2] * RC(4.0) * z, 6); ADD3(dRGBdz, SH_C2[3] * x, 7);
                    if (sc->D > 2) {
                        DSH(9, SH_C3[0] * y * (RC(3.0) * xx - yy)); DSH(10, SH_C3[1] * xy * z); DSH(11, SH_C3[2] * y * (RC(4.0) * zz - xx - yy));
                        DSH(12, SH_C3[3] * z * (RC(2.0) * zz - RC(3.0) * xx - RC(3.0) * yy)); DSH(13, SH_C3[4] * x * (RC(4.0) * zz - xx - yy));
                        DSH(14, SH_C3[5] * z * (xx - yy)); DSH(15, SH_C3[6] * x * (xx - RC(3.0) * yy));
                        ADD3(dRGBdx, SH_C3[0] * RC(6.0) * xy, 9); ADD3(dRGBdx, SH_C3[1] * yz, 10); ADD3(dRGBdx, SH_C3[2] * RC(-2.0) * xy, 11);
                        ADD3(dRGBdx, SH_C3[3] * RC(-6.0) * xz, 12); ADD3(dRGBdx, SH_C3[4] * (RC(-3.0) * xx + RC(4.0) * zz - yy), 13);
                        ADD3(dRGBdx, SH_C3[5] * RC(2.0) * xz, 14); ADD3(dRGBdx, SH_C3[6] * RC(3.0) * (xx - yy), 15);
                        ADD3(dRGBdy, SH_C3[0] * RC(3.0) * (xx - yy), 9); ADD3(dRGBdy, SH_C3[1] * xz, 10);
                        ADD3(dRGBdy, SH_C3[2] * (RC(-3.0) * yy + RC(4.0) * zz - xx), 11); ADD3(dRGBdy, SH_C3[3] * RC(-6.0) * yz, 12);
                        ADD3(dRGBdy, SH_C3[4] * RC(-2.0) * xy, 13); ADD3(dRGBdy, SH_C3[5] * RC(-2.0) * yz, 14); ADD3(dRGBdy, SH_C3[6] * RC(-6.0) * xy, 15);
                        ADD3(dRGBdz, SH_C3[1] * xy, 10); ADD3(dRGBdz, SH_C3[2] * RC(8.0) * yz, 11); ADD3(dRGBdz, SH_C3[3] * RC(3.0) * (RC(2.0) * zz - xx - yy), 12);
                        ADD3(dRGBdz, SH_C3[4] * RC(8.0) * xz, 13); ADD3(dRGBdz, SH_C3[5] * (xx - yy), 14);
                    }
                }
            }
#undef DSH
#undef ADD3
            const REAL dd[3] = {dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1] + dRGBdx[2] * dRGB[2],
                                dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1] + dRGBdy[2] * dRGB[2],
                                dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1] + dRGBdz[2] * dRGB[2]};
            /* dnormvdv (auxiliary.h:109-120) */
            const REAL sum2 = dir_o[0] * dir_o[0] + dir_o[1] * dir_o[1] + dir_o[2] * dir_o[2];
            const REAL inv32 = RC(1.0) / R_SQRT(sum2 * sum2 * sum2);
            const REAL dm[3] = {((sum2 - dir_o[0] * dir_o[0]) * dd[0] - dir_o[1] * dir_o[0] * dd[1] - dir_o[2] * dir_o[0] * dd[2]) * inv32,
                                (-dir_o[0] * dir_o[1] * dd[0] + (sum2 - dir_o[1] * dir_o[1]) * dd[1] - dir_o[2] * dir_o[1] * dd[2]) * inv32,
                                (-dir_o[0] * dir_o[2] * dd[0] - dir_o[1] * dir_o[2] * dd[1] + (sum2 - dir_o[2] * dir_o[2]) * dd[2]) * inv32};
            for (int k = 0; k < 3; ++k) { dmean[k] += dm[k]; tau[k] += -dm[k]; }
        }

        /* ---- cov3D backward (backward.cu:350-413) ---- */
        if (sc->scales) {
            const REAL s[3] = {sc->scale_modifier * sc->scales[3 * i], sc->scale_modifier * sc->scales[3 * i + 1], sc->scale_modifier * sc->scales[3 * i + 2]};
            const REAL r = sc->rotations[4 * i], x = sc->rotations[4 * i + 1], y = sc->rotations[4 * i + 2], z = sc->rotations[4 * i + 3];
            /* Rm[a][k]: a = glm column, k = glm row */
            const REAL Rm[3][3] = {{RC(1.0) - RC(2.0) * (y * y + z * z), RC(2.0) * (x * y - r * z), RC(2.0) * (x * z + r * y)},
                                   {RC(2.0) * (x * y + r * z), RC(1.0) - RC(2.0) * (x * x + z * z), RC(2.0) * (y * z - r * x)},
                                   {RC(2.0) * (x * z - r * y), RC(2.0) * (y * z + r * x), RC(1.0) - RC(2.0) * (x * x + y * y)}};
            const REAL dS[3][3] = {{dcov[0], RC(0.5) * dcov[1], RC(0.5) * dcov[2]}, {RC(0.5) * dcov[1], dcov[3], RC(0.5) * dcov[4]},
                                   {RC(0.5) * dcov[2], RC(0.5) * dcov[4], dcov[5]}};
            /* dMt[j][i] = dL_dMt[col j][row i] = 2 * sum_k (s_j Rm[k][j]) dS[i][k] */
            REAL dMt[3][3];
            for (int j = 0; j < 3; ++j) for (int ii = 0; ii < 3; ++ii) dMt[j][ii] = 2 * s[j] * (Rm[0][j] * dS[ii][0] + Rm[1][j] * dS[ii][1] + Rm[2][j] * dS[ii][2]);
            REAL* dsc = o->dL_dscales + 3 * (size_t)i;
            for (int j = 0; j < 3; ++j) dsc[j] = Rm[0][j] * dMt[j][0] + Rm[1][j] * dMt[j][1] + Rm[2][j] * dMt[j][2];
            for (int j = 0; j < 3; ++j) for (int ii = 0; ii < 3; ++ii) dMt[j][ii] *= s[j];
            REAL* dq = o->dL_drots + 4 * (size_t)i;
            dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
        }
        for (int k = 0; k < 3; ++k) o->dL_dmeans3D[3 * i + k] = dmean[k];
    }
    /* pose-gradient reduction (reference __init__.py:152-154) */
    for (int k = 0; k < 6; ++k) {
        double s = 0;
        for (int i = 0; i < P; ++i) s += o->dL_dtau_rows[6 * (size_t)i + k];
        o->dL_dtau[k] = (REAL)s;
    }
    return 0;
}

int ORACLE_NAME(oracle_mark_visible)(int32_t P, const REAL* means3D, const REAL* V, uint8_t* present) {
    for (int32_t i = 0; i < P; ++i) {
        const REAL d = row_affine(V[2], V[6], V[10], V[14], means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
        present[i] = d <= RC(0.2) ? 0 : 1;
    }
    return 0;
}
