// project.cu -- per-Gaussian projection / EWA splat / SH colour (forward "preprocess").
//
// Replaces preprocessCUDA<3> (DGR/cuda_rasterizer/forward.cu:157-258) together with
// in_frustum (auxiliary.h:139-164), computeCov3D (forward.cu:120-154), computeCov2D
// (forward.cu:76-115), computeColorFromSH (forward.cu:22-73), ndc2Pix/getRect
// (auxiliary.h:41-56), and the per-Gaussian half of the tile binning (the reference's
// tiles_touched + InclusiveSum; here a per-TILE histogram, see binning.cu).
//
// ARITHMETIC CONTRACT.  `radii`, the tile rectangle and the depth bits decide the
// integer outputs (point_list, ranges) that must be bit-exact against the reference.
// Every operation on that path is therefore written with explicit-rounding intrinsics
// (__fmul_rn/__fmaf_rn/... never re-contracted by nvcc/ptxas) in exactly the operation
// order the reference build executes on sm_100a (established from its PTX + SASS, see
// DESIGN.md "Arithmetic contract").  oracle/g4r_oracle.c follows the same contract with
// fmaf()/-ffp-contract=off on the CPU.  Notation: fma(a,b,c) = a*b+c with one rounding.
#include "g4r_common.cuh"
#include <math_constants.h>

struct ProjectParams {
    int P, D, M, W, H;
    uint32_t gx, gy;
    float tan_fovx, tan_fovy, focal_x, focal_y, scale_modifier;
    const float *means3D, *opacities, *shs, *shs_rest, *colors_precomp, *scales, *rotations, *cov3D_precomp;
    int scale_dim;                // raw mode: 1 = isotropic _scaling [P,1]
    const uint8_t* mask;          // static mask / dynamic offsets (include/g4r.h), all NULL on the reference surface
    const int32_t* dyn_slot;
    const float *dx, *ds, *dr;
    const float *viewmatrix, *projmatrix, *campos;
    float4* rec;
    uint8_t* clamped;
    int32_t* radii;
    int32_t* n_touched;
    uint32_t* tile_counts;
};

// a*x + b*y + c*z + d exactly as the reference evaluates transformPoint4x3/4x4 rows
// (auxiliary.h:58-78): t = y*b; t = fma(x,a,t); t = fma(z,c,t); t = d + t.
static __device__ __forceinline__ float affine_row(float a, float b, float c, float d, float x, float y, float z) {
    float t = __fmul_rn(y, b);
    t = __fmaf_rn(x, a, t);
    t = __fmaf_rn(z, c, t);
    return __fadd_rn(d, t);
}
// p*q + r*s + u*v the way every 3-term glm dot/matrix entry is evaluated there:
// t = r*s; t = fma(p,q,t); t = fma(u,v,t).
static __device__ __forceinline__ float dot3_mid_first(float p, float q, float r, float s, float u, float v) {
    float t = __fmul_rn(r, s);
    t = __fmaf_rn(p, q, t);
    return __fmaf_rn(u, v, t);
}

// Stage a [rows,3] float array for this CTA's 256 rows into shared memory with 128-bit loads.
template <bool kVec>
static __device__ __forceinline__ void stage_rows3(float* s_dst, const float* __restrict__ src, int row0, int P) {
    const int rows = min(G4R_BLOCK, P - row0);
    const int nfl = rows * 3;
    const float* base = src + (size_t)row0 * 3;
    if (kVec) {
        const int nvec = nfl >> 2;
        const float4* b4 = reinterpret_cast<const float4*>(base);   // row0*12 B is 16 B-aligned (row0 % 256 == 0)
        for (int i = threadIdx.x; i < nvec; i += G4R_BLOCK) reinterpret_cast<float4*>(s_dst)[i] = __ldg(b4 + i);
        for (int i = (nvec << 2) + threadIdx.x; i < nfl; i += G4R_BLOCK) s_dst[i] = __ldg(base + i);
    } else {
        for (int i = threadIdx.x; i < nfl; i += G4R_BLOCK) s_dst[i] = __ldg(base + i);
    }
}

// kRaw: the inputs are GaussianModel's raw parameters (G4R_ACT_RAW); the activations are applied right after the loads.
template <bool kVec, bool kRaw = false>
__global__ void __launch_bounds__(G4R_BLOCK) project_kernel(const ProjectParams p) {
    __shared__ __align__(16) float s_mean[G4R_BLOCK * 3];
    __shared__ __align__(16) float s_aux[G4R_BLOCK * 3];   // scales
    __shared__ float s_cam[32];                            // view matrix [0,16), full projection [16,32): broadcast reads
    const int row0 = blockIdx.x * G4R_BLOCK;
    const int i = row0 + threadIdx.x;
    const bool has_scale = p.cov3D_precomp == nullptr;

    const bool iso = kRaw && p.scale_dim == 1;
    if (threadIdx.x < 32) s_cam[threadIdx.x] = __ldg((threadIdx.x < 16 ? p.viewmatrix : p.projmatrix - 16) + threadIdx.x);
    // The per-Gaussian loads that the arithmetic only needs late (rotation, opacity) are issued here, together with the staging
    // loads, so that their latency overlaps instead of adding up behind the cull branches (ncu: 45 % of the stall samples of
    // the round-1 kernel were long-scoreboard waits on loads issued one after the other).
    float4 q_pre = make_float4(0.f, 0.f, 0.f, 0.f);
    float o_pre = 0.0f;
    if (i < p.P) {
        if (has_scale) {
            if (kVec) q_pre = __ldg(reinterpret_cast<const float4*>(p.rotations) + i);
            else { const float* q = p.rotations + (size_t)i * 4; q_pre = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3)); }
        }
        o_pre = __ldg(p.opacities + i);
    }
    stage_rows3<kVec>(s_mean, p.means3D, row0, p.P);
    if (has_scale && !iso) stage_rows3<kVec>(s_aux, p.scales, row0, p.P);
    __syncthreads();
    if (i >= p.P) return;
    p.n_touched[i] = 0;                          // accumulated by composite_forward_kernel
    if (p.mask != nullptr && p.mask[i] == 0) { p.radii[i] = 0; return; }     // masked out: as if culled

    float x = s_mean[threadIdx.x * 3 + 0], y = s_mean[threadIdx.x * 3 + 1], z = s_mean[threadIdx.x * 3 + 2];
    const int slot = p.dyn_slot != nullptr ? p.dyn_slot[i] : -1;              // >= 0: row of the dynamic offsets
    if (slot >= 0 && p.dx != nullptr) {
        x += __ldg(p.dx + (size_t)slot * 3); y += __ldg(p.dx + (size_t)slot * 3 + 1); z += __ldg(p.dx + (size_t)slot * 3 + 2);
    }
    const float* __restrict__ Q = s_cam + 16;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = s_cam[k];

    int radius_i = 0;
    // ---- near cull: p_view.z <= 0.2 (auxiliary.h:152-162) -----------------------------------
    const float depth = affine_row(v[2], v[6], v[10], v[14], x, y, z);
    if (depth <= 0.2f) { p.radii[i] = 0; return; }

    // ---- homogeneous projection (forward.cu:199-202) -------------------------------------------
    const float hx_ = affine_row(Q[0], Q[4], Q[8], Q[12], x, y, z);
    const float hy_ = affine_row(Q[1], Q[5], Q[9], Q[13], x, y, z);
    const float hw_ = affine_row(Q[3], Q[7], Q[11], Q[15], x, y, z);
    const float pw = __frcp_rn(__fadd_rn(hw_, 0.0000001f));
    const float ndc_x = __fmul_rn(hx_, pw);
    const float ndc_y = __fmul_rn(hy_, pw);

    // ---- 3D covariance (forward.cu:120-154), or the precomputed one (forward.cu:207-215) --------
    float c0, c1, c2, c3, c4, c5;
    if (has_scale) {
        float s0 = iso ? __ldg(p.scales + i) : s_aux[threadIdx.x * 3 + 0];
        float s1 = iso ? s0 : s_aux[threadIdx.x * 3 + 1];
        float s2 = iso ? s0 : s_aux[threadIdx.x * 3 + 2];
        if (kRaw) { s0 = expf(s0); s1 = iso ? s0 : expf(s1); s2 = iso ? s0 : expf(s2); }     // scaling_activation = exp
        if (slot >= 0 && p.ds != nullptr) {          // scales + dscale (gaussian_renderer/__init__.py:167-170)
            s0 += __ldg(p.ds + (size_t)slot * 3); s1 += __ldg(p.ds + (size_t)slot * 3 + 1); s2 += __ldg(p.ds + (size_t)slot * 3 + 2);
        }
        const float sx = __fmul_rn(p.scale_modifier, s0);
        const float sy = __fmul_rn(p.scale_modifier, s1);
        const float sz = __fmul_rn(p.scale_modifier, s2);
        float qr = q_pre.x, qx = q_pre.y, qy = q_pre.z, qz = q_pre.w;     // (r,x,y,z), NOT normalised (forward.cu:129)
        if (kRaw) {                                  // rotation_activation = normalize
            const float n = g4r_quat_norm(qr, qx, qy, qz);
            qr = __fdiv_rn(qr, n); qx = __fdiv_rn(qx, n); qy = __fdiv_rn(qy, n); qz = __fdiv_rn(qz, n);
        }
        if (slot >= 0 && p.dr != nullptr) {          // get_rotation + drot (:171-174): added AFTER the normalisation
            qr += __ldg(p.dr + (size_t)slot * 4); qx += __ldg(p.dr + (size_t)slot * 4 + 1);
            qy += __ldg(p.dr + (size_t)slot * 4 + 2); qz += __ldg(p.dr + (size_t)slot * 4 + 3);
        }
        const float yy = __fmul_rn(qy, qy), zz = __fmul_rn(qz, qz);
        const float xz = __fmul_rn(qx, qz), rz = __fmul_rn(qr, qz), rx = __fmul_rn(qr, qx);
        const float yy_zz = __fadd_rn(yy, zz);
        const float xx_zz = __fmaf_rn(qx, qx, zz);
        const float xx_yy = __fmaf_rn(qx, qx, yy);
        const float xy_m_rz = __fmaf_rn(qx, qy, -rz), xy_p_rz = __fmaf_rn(qx, qy, rz);
        const float xz_p_ry = __fmaf_rn(qr, qy, xz), xz_m_ry = __fmaf_rn(-qr, qy, xz);
        const float yz_m_rx = __fmaf_rn(qy, qz, -rx), yz_p_rx = __fmaf_rn(qy, qz, rx);
        // rows of M = S*R^T-layout of the reference: m[a][k] = s_k * R(col a, row k)
        const float m00 = __fmul_rn(sx, __fsub_rn(1.0f, __fadd_rn(yy_zz, yy_zz)));
        const float m01 = __fmul_rn(sy, __fadd_rn(xy_m_rz, xy_m_rz));
        const float m02 = __fmul_rn(sz, __fadd_rn(xz_p_ry, xz_p_ry));
        const float m10 = __fmul_rn(sx, __fadd_rn(xy_p_rz, xy_p_rz));
        const float m11 = __fmul_rn(sy, __fsub_rn(1.0f, __fadd_rn(xx_zz, xx_zz)));
        const float m12 = __fmul_rn(sz, __fadd_rn(yz_m_rx, yz_m_rx));
        const float m20 = __fmul_rn(sx, __fadd_rn(xz_m_ry, xz_m_ry));
        const float m21 = __fmul_rn(sy, __fadd_rn(yz_p_rx, yz_p_rx));
        const float m22 = __fmul_rn(sz, __fsub_rn(1.0f, __fadd_rn(xx_yy, xx_yy)));
        c0 = dot3_mid_first(m00, m00, m01, m01, m02, m02);
        c1 = dot3_mid_first(m10, m00, m11, m01, m12, m02);
        c2 = dot3_mid_first(m20, m00, m21, m01, m22, m02);
        c3 = dot3_mid_first(m10, m10, m11, m11, m12, m12);
        c4 = dot3_mid_first(m20, m10, m21, m11, m22, m12);
        c5 = dot3_mid_first(m20, m20, m21, m21, m22, m22);
    } else {
        const float* c = p.cov3D_precomp + (size_t)i * 6;
        c0 = __ldg(c + 0); c1 = __ldg(c + 1); c2 = __ldg(c + 2); c3 = __ldg(c + 3); c4 = __ldg(c + 4); c5 = __ldg(c + 5);
    }

    // ---- EWA 2D covariance (forward.cu:76-115) ----------------------------------------------------
    const float tx = affine_row(v[0], v[4], v[8], v[12], x, y, z);
    const float ty = affine_row(v[1], v[5], v[9], v[13], x, y, z);
    const float tz = depth;
    const float limx = __fmul_rn(p.tan_fovx, 1.3f), limy = __fmul_rn(p.tan_fovy, 1.3f);
    const float cx = fminf(limx, fmaxf(-limx, __fdiv_rn(tx, tz)));
    const float cy = fminf(limy, fmaxf(-limy, __fdiv_rn(ty, tz)));
    const float tz2 = __fmul_rn(tz, tz);
    const float j00 = __fdiv_rn(p.focal_x, tz);
    const float j02 = __fdiv_rn(__fmul_rn(p.focal_x, __fmul_rn(cx, -tz)), tz2);
    const float j11 = __fdiv_rn(p.focal_y, tz);
    const float j12 = __fdiv_rn(__fmul_rn(p.focal_y, __fmul_rn(cy, -tz)), tz2);
    // rows u, w of the 2x3 matrix J*R_w2c:  u_k = fma(V[2+4k], j02, V[0+4k]*j00), w_k = fma(V[2+4k], j12, V[1+4k]*j11)
    const float u0 = __fmaf_rn(v[2], j02, __fmul_rn(v[0], j00));
    const float u1 = __fmaf_rn(v[6], j02, __fmul_rn(v[4], j00));
    const float u2 = __fmaf_rn(v[10], j02, __fmul_rn(v[8], j00));
    const float w0 = __fmaf_rn(v[2], j12, __fmul_rn(v[1], j11));
    const float w1 = __fmaf_rn(v[6], j12, __fmul_rn(v[5], j11));
    const float w2 = __fmaf_rn(v[10], j12, __fmul_rn(v[9], j11));
    // A = (rows) * Sigma, then cov = A * rows^T; each 3-term sum is "middle product first".
    const float au0 = dot3_mid_first(u0, c0, u1, c1, u2, c2);
    const float au1 = dot3_mid_first(u0, c1, u1, c3, u2, c4);
    const float au2 = dot3_mid_first(u0, c2, u1, c4, u2, c5);
    const float aw0 = dot3_mid_first(w0, c0, w1, c1, w2, c2);
    const float aw1 = dot3_mid_first(w0, c1, w1, c3, w2, c4);
    const float aw2 = dot3_mid_first(w0, c2, w1, c4, w2, c5);
    const float cov_a = __fadd_rn(dot3_mid_first(u0, au0, u1, au1, u2, au2), 0.3f);
    const float cov_b = dot3_mid_first(u0, aw0, u1, aw1, u2, aw2);
    const float cov_c = __fadd_rn(dot3_mid_first(w0, aw0, w1, aw1, w2, aw2), 0.3f);

    // ---- conic + radius (forward.cu:221-234) --------------------------------------------------------
    const float det = __fmaf_rn(cov_a, cov_c, -__fmul_rn(cov_b, cov_b));
    if (det == 0.0f) { p.radii[i] = 0; return; }
    const float det_inv = __frcp_rn(det);
    const float con_x = __fmul_rn(cov_c, det_inv);
    const float con_y = __fmul_rn(det_inv, -cov_b);
    const float con_z = __fmul_rn(cov_a, det_inv);
    const float mid = __fmul_rn(__fadd_rn(cov_a, cov_c), 0.5f);
    const float disc = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
    const float lam = fmaxf(__fadd_rn(mid, disc), __fsub_rn(mid, disc));
    const float radius_f = ceilf(__fmul_rn(__fsqrt_rn(lam), 3.0f));

    // ---- pixel centre (auxiliary.h:41-44: double arithmetic, fused multiply-add) ---------------------
    const float px = (float)(__fma_rn((double)ndc_x + 1.0, (double)p.W, -1.0) * 0.5);
    const float py = (float)(__fma_rn((double)ndc_y + 1.0, (double)p.H, -1.0) * 0.5);
    radius_i = f2i_rz(radius_f);
    const TileRect r = tile_rect(px, py, radius_i, p.gx, p.gy);
    if ((r.x1 - r.x0) * (r.y1 - r.y0) == 0u) { p.radii[i] = 0; return; }

    // ---- colour: SH evaluation (forward.cu:22-73) or precomputed -------------------------------------
    float cr, cg, cb;
    if (p.colors_precomp != nullptr) {
        const float* c = p.colors_precomp + (size_t)i * 3;
        cr = __ldg(c + 0); cg = __ldg(c + 1); cb = __ldg(c + 2);
    } else {
        float dx = x - __ldg(p.campos + 0), dy = y - __ldg(p.campos + 1), dz = z - __ldg(p.campos + 2);
        const float len = sqrtf(dx * dx + dy * dy + dz * dz);
        dx = dx / len; dy = dy / len; dz = dz / len;
        // raw mode: coefficient 0 lives in _features_dc [P,1,3], the others in _features_rest [P,M-1,3]
        const float* sh = kRaw ? p.shs + (size_t)i * 3 : p.shs + (size_t)i * p.M * 3;
        // coefficient k >= 1 of Gaussian i sits at shs_rest[(i*(M-1) + k-1)*3]; the base is shifted by one coefficient so that the
        // SH(k, c) macro can index with k directly (signed arithmetic: the shift is negative for i == 0)
        const float* shr = (kRaw && p.M > 1) ? p.shs_rest + ((ptrdiff_t)i * (p.M - 1) - 1) * 3 : sh;
#define SH(k, c) __ldg(((k) == 0 ? sh : shr) + (k) * 3 + (c))
        float col[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float r_ = G4R_SH_C0 * SH(0, c);
            if (p.D > 0) {
                r_ = r_ - G4R_SH_C1 * dy * SH(1, c) + G4R_SH_C1 * dz * SH(2, c) - G4R_SH_C1 * dx * SH(3, c);
                if (p.D > 1) {
                    const float xx = dx * dx, yy = dy * dy, zz = dz * dz, xy = dx * dy, yz = dy * dz, xz = dx * dz;
                    r_ = r_ + G4R_SH_C2_0 * xy * SH(4, c) + G4R_SH_C2_1 * yz * SH(5, c) +
                         G4R_SH_C2_2 * (2.0f * zz - xx - yy) * SH(6, c) + G4R_SH_C2_3 * xz * SH(7, c) +
                         G4R_SH_C2_4 * (xx - yy) * SH(8, c);
                    if (p.D > 2) {
                        r_ = r_ + G4R_SH_C3_0 * dy * (3.0f * xx - yy) * SH(9, c) + G4R_SH_C3_1 * xy * dz * SH(10, c) +
                             G4R_SH_C3_2 * dy * (4.0f * zz - xx - yy) * SH(11, c) +
                             G4R_SH_C3_3 * dz * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SH(12, c) +
                             G4R_SH_C3_4 * dx * (4.0f * zz - xx - yy) * SH(13, c) + G4R_SH_C3_5 * dz * (xx - yy) * SH(14, c) +
                             G4R_SH_C3_6 * dx * (xx - 3.0f * yy) * SH(15, c);
                    }
                }
            }
            col[c] = r_ + 0.5f;
        }
#undef SH
        const uint8_t cl = (uint8_t)((col[0] < 0.0f ? 1 : 0) | (col[1] < 0.0f ? 2 : 0) | (col[2] < 0.0f ? 4 : 0));
        p.clamped[i] = cl;
        cr = fmaxf(col[0], 0.0f); cg = fmaxf(col[1], 0.0f); cb = fmaxf(col[2], 0.0f);
    }

    // ---- conservative cull threshold for the composite kernels -----------------------------------------
    // A pixel can only receive alpha >= 1/255 from this splat if q(d) = A dx^2 + 2B dx dy + C dy^2 <= 2*ln(255*o).
    // The composite kernels minimise q over a warp's 8x4 pixel patch and skip the splat when the minimum exceeds
    // cull_q; the small padding absorbs float rounding of the per-pixel evaluation, so a skipped (warp, splat)
    // pair is always one the reference would have evaluated to alpha < 1/255 for all 32 pixels.
    const float o = kRaw ? g4r_sigmoid(o_pre) : o_pre;                                     // opacity_activation = sigmoid
    float cull_q = CUDART_INF_F;                    // no culling (degenerate conic / NaN)
    if (o < (1.0f / 255.0f)) {
        cull_q = -1.0f;                             // alpha = o*exp(power<=0) < 1/255 everywhere
    } else {
        const float detq = fmaf(con_x, con_z, -con_y * con_y);
        if (con_x > 0.0f && con_z > 0.0f && detq > 1e-4f * con_x * con_z) cull_q = fmaf(2.0f * logf(255.0f * o), 1.002f, 0.05f);
    }

    float4* rec = p.rec + (size_t)i * 3;
    rec[0] = make_float4(px, py, con_x, con_y);
    rec[1] = make_float4(con_z, o, depth, cr);
    rec[2] = make_float4(cg, cb, cull_q, 0.0f);
    p.radii[i] = radius_i;

    // ---- per-tile instance histogram (replaces tiles_touched + InclusiveSum) ----------------------------
    if (p.tile_counts == nullptr) return;           // sharded render: tiles are counted after the all-gather
    for (uint32_t ty_ = r.y0; ty_ < r.y1; ++ty_)
        for (uint32_t tx_ = r.x0; tx_ < r.x1; ++tx_) atomicAdd(p.tile_counts + (size_t)(ty_ * p.gx + tx_) * G4R_COUNT_STRIDE, 1u);
}

int launch_project(const G4RFrame& f, const G4RGaussians& g, void* geom, void* img, int32_t* radii, int32_t* n_touched,
                   cudaStream_t s) {
    const GeomLayout gl(g.P);
    const ImageLayout il(f.width, f.height);
    ProjectParams p;
    p.P = g.P; p.D = f.sh_degree; p.M = f.sh_coeffs; p.W = f.width; p.H = f.height;
    p.gx = (uint32_t)il.tiles_x; p.gy = (uint32_t)il.tiles_y;
    p.tan_fovx = f.tan_fovx; p.tan_fovy = f.tan_fovy;
    // rasterizer_impl.cu:225-226: float division of the int extent by (2.0f * tan)
    p.focal_x = (float)f.width / (2.0f * f.tan_fovx);
    p.focal_y = (float)f.height / (2.0f * f.tan_fovy);
    p.scale_modifier = f.scale_modifier;
    p.means3D = g.means3D; p.opacities = g.opacities; p.shs = g.shs; p.shs_rest = g.shs_rest; p.colors_precomp = g.colors_precomp;
    p.scales = g.scales; p.rotations = g.rotations; p.cov3D_precomp = g.cov3D_precomp;
    p.scale_dim = g.scale_dim == 1 ? 1 : 3;
    p.mask = g.mask; p.dyn_slot = g.dyn_slot; p.dx = g.dx; p.ds = g.ds; p.dr = g.dr;
    p.viewmatrix = f.viewmatrix; p.projmatrix = f.projmatrix; p.campos = f.campos;
    p.rec = reinterpret_cast<float4*>((char*)geom + gl.rec);
    p.clamped = reinterpret_cast<uint8_t*>((char*)geom + gl.clamped);
    p.radii = radii;
    p.n_touched = n_touched;
    p.tile_counts = img ? reinterpret_cast<uint32_t*>((char*)img + il.counts) : nullptr;
    const int blocks = (g.P + G4R_BLOCK - 1) / G4R_BLOCK;
    const bool raw = g.activation == G4R_ACT_RAW;
    const bool vec = (((uintptr_t)g.means3D | (uintptr_t)g.scales | (uintptr_t)g.rotations) & 15u) == 0;
    g4r_stage_begin(ST_PROJECT, s);
    if (raw) {
        if (vec) project_kernel<true, true><<<blocks, G4R_BLOCK, 0, s>>>(p);
        else     project_kernel<false, true><<<blocks, G4R_BLOCK, 0, s>>>(p);
    } else {
        if (vec) project_kernel<true><<<blocks, G4R_BLOCK, 0, s>>>(p);
        else     project_kernel<false><<<blocks, G4R_BLOCK, 0, s>>>(p);
    }
    g4r_stage_end(ST_PROJECT, s);
    G4R_LAUNCH_OK("project_kernel");
    return G4R_OK;
}

// ---- markVisible (rasterizer_impl.cu:54-66,141-153): present = p_view.z > 0.2 -----------------------------
__global__ void __launch_bounds__(G4R_BLOCK) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                                  const float* __restrict__ V, uint8_t* __restrict__ present) {
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (i >= P) return;
    const float x = __ldg(means3D + 3 * (size_t)i), y = __ldg(means3D + 3 * (size_t)i + 1), z = __ldg(means3D + 3 * (size_t)i + 2);
    const float depth = affine_row(__ldg(V + 2), __ldg(V + 6), __ldg(V + 10), __ldg(V + 14), x, y, z);
    present[i] = (depth <= 0.2f) ? 0 : 1;
}

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s) {
    mark_visible_kernel<<<(P + G4R_BLOCK - 1) / G4R_BLOCK, G4R_BLOCK, 0, s>>>(P, means3D, viewmatrix, present);
    G4R_LAUNCH_OK("mark_visible_kernel");
    return G4R_OK;
}
