// gaussian_bwd.cu -- fused per-Gaussian backward.
//
// One kernel replaces computeCov2DCUDA (DGR/cuda_rasterizer/backward.cu:150-346), the backward
// preprocessCUDA<3> (:418-539), computeColorFromSH backward (:21-145), computeCov3D backward
// (:350-413), the 11 torch::zeros fills of RasterizeGaussiansBackwardCUDA
// (DGR/rasterize_points.cu:160-170) and the (P,6)->(6) pose-gradient reduction of
// DGR/diff_gaussian_rasterization/__init__.py:152-154.
//
// Every output row is written exactly once (zeros for Gaussians with radii == 0, which both
// reference kernels skip: backward.cu:163,443), so the caller never pre-zeroes anything.  The
// SE(3) pose gradient dL/dtau is reduced warp -> CTA -> 6 global atomics per CTA.
//
// The SH derivative sums (the DSH / dd{x,y,z} block) and dL/dq are the closed-form derivatives written term by term in the order
// of the reference's backward.cu:48-123,404-408 -- bug-compatible gradients need the same terms; the kernel structure around
// them (one fused kernel, write-once rows, optional outputs, in-kernel pose reduction, raw-parameter chain rule, dynamic
// offsets) is this repo's.
//
// Semantics follow the reference including its approximations:
// the pose Jacobian of the 2D mean uses proj_raw entries a, b, e only (:465-484); the
// SH view-direction term contributes -dL/dmean to rho only (:141-143); -[t]x is built from
// the CLAMPED t (:179-180,278); dL/drot has no normalisation Jacobian (:412).
#include "g4r_common.cuh"

struct GaussBwdParams {
    int P, D, M, W, H;
    float tan_fovx, tan_fovy, focal_x, focal_y, scale_modifier;
    const float *means3D, *shs, *shs_rest, *colors_precomp, *scales, *rotations, *cov3D_precomp;
    int scale_dim;                // raw mode: 1 = isotropic _scaling [P,1]
    const int32_t* dyn_slot;      // dynamic offsets (include/g4r.h); NULL on the reference surface
    const float *dx, *ds, *dr;
    float *dL_ddx, *dL_dds, *dL_ddr;
    const float *viewmatrix, *projmatrix, *projmatrix_raw, *campos;
    const int32_t* radii;
    const float4* rec;
    const uint8_t* clamped;
    const float* acc;
    float *dL_dmeans3D, *dL_dmeans2D, *dL_dopacity, *dL_dshs, *dL_dshs_rest, *dL_dcolors_precomp, *dL_dscales, *dL_drotations, *dL_dcov3D;
    float* dL_dtau;
};

struct V3 { float x, y, z; };
static __device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
static __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// kRaw (G4R_ACT_RAW): inputs are GaussianModel's raw parameters; the activations are re-applied after the loads and the
// gradients leave through the chain rule of exp / sigmoid / normalize (what autograd does for the reference's prelude,
// gaussian_model.py:100-128): d/d_scaling = dL/ds * s, d/d_opacity = dL/do * o (1 - o), d/d_rotation = (g - q (q.g)) / |raw|;
// SH gradients are split into the [P,1,3] and [P,M-1,3] parameter tensors.
template <bool kRaw>
__global__ void __launch_bounds__(G4R_BLOCK, 4) gaussian_backward_kernel(const GaussBwdParams p) {
    __shared__ float s_tau[G4R_BLOCK / 32][6];
    __shared__ float v[16], pm[16];      // view / full projection matrices: broadcast reads instead of 32 live registers
    if (threadIdx.x < 16) { v[threadIdx.x] = __ldg(p.viewmatrix + threadIdx.x); pm[threadIdx.x] = __ldg(p.projmatrix + threadIdx.x); }
    __syncthreads();
    const int i = blockIdx.x * G4R_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float tau[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    const bool in_range = i < p.P;
    const bool visible = in_range && p.radii[i] > 0;
    const int K = (p.D + 1) * (p.D + 1);

    if (in_range && !visible) {
        // invisible Gaussian: all gradients are zero
        const size_t o3 = (size_t)i * 3;
        if (p.dL_dmeans3D) { p.dL_dmeans3D[o3] = 0.f; p.dL_dmeans3D[o3 + 1] = 0.f; p.dL_dmeans3D[o3 + 2] = 0.f; }
        if (p.dL_dmeans2D) { p.dL_dmeans2D[o3] = 0.f; p.dL_dmeans2D[o3 + 1] = 0.f; p.dL_dmeans2D[o3 + 2] = 0.f; }
        if (p.dL_dopacity) p.dL_dopacity[i] = 0.f;
        if (kRaw) {
            if (p.dL_dshs) { float* d = p.dL_dshs + o3; d[0] = 0.f; d[1] = 0.f; d[2] = 0.f; }
            if (p.dL_dshs_rest) { float* r = p.dL_dshs_rest + (size_t)i * (p.M - 1) * 3; for (int k = 0; k < (p.M - 1) * 3; ++k) r[k] = 0.f; }
        } else if (p.dL_dshs) { float* d = p.dL_dshs + (size_t)i * p.M * 3; for (int k = 0; k < p.M * 3; ++k) d[k] = 0.f; }
        if (p.dL_dcolors_precomp) { p.dL_dcolors_precomp[o3] = 0.f; p.dL_dcolors_precomp[o3 + 1] = 0.f; p.dL_dcolors_precomp[o3 + 2] = 0.f; }
        if (p.dL_dscales) {
            if (kRaw && p.scale_dim == 1) p.dL_dscales[i] = 0.f;
            else { p.dL_dscales[o3] = 0.f; p.dL_dscales[o3 + 1] = 0.f; p.dL_dscales[o3 + 2] = 0.f; }
        }
        if (p.dL_drotations) reinterpret_cast<float4*>(p.dL_drotations)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.dL_dcov3D) { float* d = p.dL_dcov3D + (size_t)i * 6; for (int k = 0; k < 6; ++k) d[k] = 0.f; }
    }

    const int slot = (in_range && p.dyn_slot != nullptr) ? p.dyn_slot[i] : -1;
    if (in_range && !visible && slot >= 0) {     // a dynamic Gaussian that was culled / masked: its offset rows get zeros
        if (p.dL_ddx) { float* d = p.dL_ddx + (size_t)slot * 3; d[0] = 0.f; d[1] = 0.f; d[2] = 0.f; }
        if (p.dL_dds) { float* d = p.dL_dds + (size_t)slot * 3; d[0] = 0.f; d[1] = 0.f; d[2] = 0.f; }
        if (p.dL_ddr) reinterpret_cast<float4*>(p.dL_ddr)[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    if (visible) {
        float mx = __ldg(p.means3D + (size_t)i * 3), my = __ldg(p.means3D + (size_t)i * 3 + 1), mz = __ldg(p.means3D + (size_t)i * 3 + 2);
        if (slot >= 0 && p.dx != nullptr) {
            mx += __ldg(p.dx + (size_t)slot * 3); my += __ldg(p.dx + (size_t)slot * 3 + 1); mz += __ldg(p.dx + (size_t)slot * 3 + 2);
        }

        // accumulated screen-space gradients from composite_backward_kernel
        const float4* arow = reinterpret_cast<const float4*>(p.acc + (size_t)i * G4R_ACC_STRIDE);
        const float4 g0 = __ldg(arow), g1 = __ldg(arow + 1), g2 = __ldg(arow + 2);
        const float dmean2D_x = g0.x, dmean2D_y = g0.y;
        const float dconic_x = g0.z, dconic_y = g0.w, dconic_w = g1.x;
        const float dopacity = g1.y;
        const V3 dcolor = v3(g1.z, g1.w, g2.x);
        const float ddepth = g2.y;

        // ---- 3D covariance: recompute from scale/rotation, or take the precomputed one ------------
        const bool has_scale = p.cov3D_precomp == nullptr;
        float c0, c1, c2, c3, c4, c5;
        float sx = 0.f, sy = 0.f, sz = 0.f, qr = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
        float nr = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;    // the normalised quaternion before the dynamic offset (raw mode chain rule)
        float act_s[3] = {0.f, 0.f, 0.f}, qnorm = 1.f;   // raw mode: activated scales (before the modifier), |_rotation|
        float R[3][3];   // R[a][k] = reference's glm R[col a][row k]; M[a][k] = s_k * R[a][k]
        if (has_scale) {
            const bool iso = kRaw && p.scale_dim == 1;
            float s0 = iso ? __ldg(p.scales + i) : __ldg(p.scales + (size_t)i * 3);
            float s1 = iso ? s0 : __ldg(p.scales + (size_t)i * 3 + 1);
            float s2 = iso ? s0 : __ldg(p.scales + (size_t)i * 3 + 2);
            if (kRaw) { s0 = expf(s0); s1 = iso ? s0 : expf(s1); s2 = iso ? s0 : expf(s2); act_s[0] = s0; act_s[1] = s1; act_s[2] = s2; }
            if (slot >= 0 && p.ds != nullptr) {
                s0 += __ldg(p.ds + (size_t)slot * 3); s1 += __ldg(p.ds + (size_t)slot * 3 + 1); s2 += __ldg(p.ds + (size_t)slot * 3 + 2);
            }
            sx = p.scale_modifier * s0;
            sy = p.scale_modifier * s1;
            sz = p.scale_modifier * s2;
            qr = __ldg(p.rotations + (size_t)i * 4); qx = __ldg(p.rotations + (size_t)i * 4 + 1);
            qy = __ldg(p.rotations + (size_t)i * 4 + 2); qz = __ldg(p.rotations + (size_t)i * 4 + 3);
            if (kRaw) {
                qnorm = g4r_quat_norm(qr, qx, qy, qz);
                qr = __fdiv_rn(qr, qnorm); qx = __fdiv_rn(qx, qnorm); qy = __fdiv_rn(qy, qnorm); qz = __fdiv_rn(qz, qnorm);
            }
            nr = qr; nx = qx; ny = qy; nz = qz;
            if (slot >= 0 && p.dr != nullptr) {
                qr += __ldg(p.dr + (size_t)slot * 4); qx += __ldg(p.dr + (size_t)slot * 4 + 1);
                qy += __ldg(p.dr + (size_t)slot * 4 + 2); qz += __ldg(p.dr + (size_t)slot * 4 + 3);
            }
            R[0][0] = 1.f - 2.f * (qy * qy + qz * qz); R[0][1] = 2.f * (qx * qy - qr * qz); R[0][2] = 2.f * (qx * qz + qr * qy);
            R[1][0] = 2.f * (qx * qy + qr * qz); R[1][1] = 1.f - 2.f * (qx * qx + qz * qz); R[1][2] = 2.f * (qy * qz - qr * qx);
            R[2][0] = 2.f * (qx * qz - qr * qy); R[2][1] = 2.f * (qy * qz + qr * qx); R[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
            float M[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a) { M[a][0] = sx * R[a][0]; M[a][1] = sy * R[a][1]; M[a][2] = sz * R[a][2]; }
            c0 = M[0][0] * M[0][0] + M[0][1] * M[0][1] + M[0][2] * M[0][2];
            c1 = M[0][0] * M[1][0] + M[0][1] * M[1][1] + M[0][2] * M[1][2];
            c2 = M[0][0] * M[2][0] + M[0][1] * M[2][1] + M[0][2] * M[2][2];
            c3 = M[1][0] * M[1][0] + M[1][1] * M[1][1] + M[1][2] * M[1][2];
            c4 = M[1][0] * M[2][0] + M[1][1] * M[2][1] + M[1][2] * M[2][2];
            c5 = M[2][0] * M[2][0] + M[2][1] * M[2][1] + M[2][2] * M[2][2];
        } else {
            const float* c = p.cov3D_precomp + (size_t)i * 6;
            c0 = __ldg(c); c1 = __ldg(c + 1); c2 = __ldg(c + 2); c3 = __ldg(c + 3); c4 = __ldg(c + 4); c5 = __ldg(c + 5);
        }

        // ---- EWA covariance backward (backward.cu:171-297) ---------------------------------------------
        V3 t = v3(v[0] * mx + v[4] * my + v[8] * mz + v[12], v[1] * mx + v[5] * my + v[9] * mz + v[13],
                  v[2] * mx + v[6] * my + v[10] * mz + v[14]);
        const float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
        const float txtz = t.x / t.z, tytz = t.y / t.z;
        t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
        t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float j00 = p.focal_x / t.z, j02 = -(p.focal_x * t.x) / (t.z * t.z);
        const float j11 = p.focal_y / t.z, j12 = -(p.focal_y * t.y) / (t.z * t.z);
        // view rotation rows: Wk = (V[k], V[k+4], V[k+8]) = k-th row of R_w2c
        const V3 W0 = v3(v[0], v[4], v[8]), W1 = v3(v[1], v[5], v[9]), W2 = v3(v[2], v[6], v[10]);
        // 2x3 matrix Tm = J * R_w2c, rows u (x) and w (y)
        const V3 u = v3(W0.x * j00 + W2.x * j02, W0.y * j00 + W2.y * j02, W0.z * j00 + W2.z * j02);
        const V3 w = v3(W1.x * j11 + W2.x * j12, W1.y * j11 + W2.y * j12, W1.z * j11 + W2.z * j12);
        // Sigma * u, Sigma * w
        const V3 Su = v3(c0 * u.x + c1 * u.y + c2 * u.z, c1 * u.x + c3 * u.y + c4 * u.z, c2 * u.x + c4 * u.y + c5 * u.z);
        const V3 Sw = v3(c0 * w.x + c1 * w.y + c2 * w.z, c1 * w.x + c3 * w.y + c4 * w.z, c2 * w.x + c4 * w.y + c5 * w.z);
        const float a = dot(u, Su) + 0.3f, b = dot(u, Sw), c = dot(w, Sw) + 0.3f;
        const float denom = a * c - b * b;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (denom2inv != 0.f) {
            dL_da = denom2inv * (-c * c * dconic_x + 2.f * b * c * dconic_y + (denom - a * c) * dconic_w);
            dL_dc = denom2inv * (-a * a * dconic_w + 2.f * a * b * dconic_y + (denom - a * c) * dconic_x);
            dL_db = denom2inv * 2.f * (b * c * dconic_x - (denom + 2.f * b * b) * dconic_y + a * b * dconic_w);
            dcov[0] = u.x * u.x * dL_da + u.x * w.x * dL_db + w.x * w.x * dL_dc;
            dcov[3] = u.y * u.y * dL_da + u.y * w.y * dL_db + w.y * w.y * dL_dc;
            dcov[5] = u.z * u.z * dL_da + u.z * w.z * dL_db + w.z * w.z * dL_dc;
            dcov[1] = 2.f * u.x * u.y * dL_da + (u.x * w.y + u.y * w.x) * dL_db + 2.f * w.x * w.y * dL_dc;
            dcov[2] = 2.f * u.x * u.z * dL_da + (u.x * w.z + u.z * w.x) * dL_db + 2.f * w.x * w.z * dL_dc;
            dcov[4] = 2.f * u.z * u.y * dL_da + (u.y * w.z + u.z * w.y) * dL_db + 2.f * w.y * w.z * dL_dc;
        }
        // dL/dTm rows
        const V3 du = v3(2.f * Su.x * dL_da + Sw.x * dL_db, 2.f * Su.y * dL_da + Sw.y * dL_db, 2.f * Su.z * dL_da + Sw.z * dL_db);
        const V3 dw = v3(2.f * Sw.x * dL_dc + Su.x * dL_db, 2.f * Sw.y * dL_dc + Su.y * dL_db, 2.f * Sw.z * dL_dc + Su.z * dL_db);
        const float dJ00 = dot(W0, du), dJ02 = dot(W2, du), dJ11 = dot(W1, dw), dJ12 = dot(W2, dw);
        const float tzi = 1.f / t.z, tz2 = tzi * tzi, tz3 = tz2 * tzi;
        const float dtx = x_grad_mul * -p.focal_x * tz2 * dJ02;
        const float dty = y_grad_mul * -p.focal_y * tz2 * dJ12;
        const float dtz = -p.focal_x * tz2 * dJ00 - p.focal_y * tz2 * dJ11 + (2.f * p.focal_x * t.x) * tz3 * dJ02 + (2.f * p.focal_y * t.y) * tz3 * dJ12;
        // pose: d p_C / d rho = I, d p_C / d theta = -[t]x with the clamped t (backward.cu:273-288)
        tau[0] += dtx; tau[1] += dty; tau[2] += dtz;
        tau[3] += -t.z * dty + t.y * dtz;
        tau[4] += t.z * dtx - t.x * dtz;
        tau[5] += -t.y * dtx + t.x * dty;
        // mean gradient through t (assignment in the reference, :292-297)
        V3 dmean = v3(v[0] * dtx + v[1] * dty + v[2] * dtz, v[4] * dtx + v[5] * dty + v[6] * dtz, v[8] * dtx + v[9] * dty + v[10] * dtz);
        // rotation part of the view matrix inside Tm (backward.cu:299-343): g_k = dL/d(column k of W)
        {
            const V3 gk[3] = {v3(j00 * du.x, j11 * dw.x, j02 * du.x + j12 * dw.x), v3(j00 * du.y, j11 * dw.y, j02 * du.y + j12 * dw.y),
                              v3(j00 * du.z, j11 * dw.z, j02 * du.z + j12 * dw.z)};
            const V3 ck[3] = {v3(v[0], v[1], v[2]), v3(v[4], v[5], v[6]), v3(v[8], v[9], v[10])};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                tau[3] += -gk[k].y * ck[k].z + gk[k].z * ck[k].y;
                tau[4] += gk[k].x * ck[k].z - gk[k].z * ck[k].x;
                tau[5] += -gk[k].x * ck[k].y + gk[k].y * ck[k].x;
            }
        }

        // ---- 2D mean -> 3D mean and pose (backward.cu:446-512) -----------------------------------------
        const float hx = pm[0] * mx + pm[4] * my + pm[8] * mz + pm[12];
        const float hy = pm[1] * mx + pm[5] * my + pm[9] * mz + pm[13];
        const float hw = pm[3] * mx + pm[7] * my + pm[11] * mz + pm[15];
        const float m_w = 1.0f / (hw + 0.0000001f);
        const float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
        dmean.x += (pm[0] * m_w - pm[3] * mul1) * dmean2D_x + (pm[1] * m_w - pm[3] * mul2) * dmean2D_y;
        dmean.y += (pm[4] * m_w - pm[7] * mul1) * dmean2D_x + (pm[5] * m_w - pm[7] * mul2) * dmean2D_y;
        dmean.z += (pm[8] * m_w - pm[11] * mul1) * dmean2D_x + (pm[9] * m_w - pm[11] * mul2) * dmean2D_y;
        {
            const float pa = __ldg(p.projmatrix_raw + 0), pb = __ldg(p.projmatrix_raw + 5), pe = __ldg(p.projmatrix_raw + 11);
            const float alpha = m_w, beta = -hx * m_w * m_w, gamma = -hy * m_w * m_w;
            const V3 pC = v3(v[0] * mx + v[4] * my + v[8] * mz + v[12], v[1] * mx + v[5] * my + v[9] * mz + v[13],
                             v[2] * mx + v[6] * my + v[10] * mz + v[14]);   // unclamped camera-space point
            const V3 d1 = v3(alpha * pa, 0.f, beta * pe), d2 = v3(0.f, alpha * pb, gamma * pe);
            // (-[pC]x)^T d = pC x d ... written out: rows of -[pC]x^T
            const V3 d1t = v3(-pC.z * d1.y + pC.y * d1.z, pC.z * d1.x - pC.x * d1.z, -pC.y * d1.x + pC.x * d1.y);
            const V3 d2t = v3(-pC.z * d2.y + pC.y * d2.z, pC.z * d2.x - pC.x * d2.z, -pC.y * d2.x + pC.x * d2.y);
            tau[0] += dmean2D_x * d1.x + dmean2D_y * d2.x;
            tau[1] += dmean2D_x * d1.y + dmean2D_y * d2.y;
            tau[2] += dmean2D_x * d1.z + dmean2D_y * d2.z;
            tau[3] += dmean2D_x * d1t.x + dmean2D_y * d2t.x;
            tau[4] += dmean2D_x * d1t.y + dmean2D_y * d2t.y;
            tau[5] += dmean2D_x * d1t.z + dmean2D_y * d2t.z;
            // depth = p_view.z (backward.cu:518-528)
            dmean.x += ddepth * v[2]; dmean.y += ddepth * v[6]; dmean.z += ddepth * v[10];
            tau[2] += ddepth;                 // d z / d rho = (0,0,1)
            tau[3] += ddepth * pC.y;          // row z of -[pC]x = (pC.y, -pC.x, 0)
            tau[4] += ddepth * -pC.x;
        }

        // ---- colour: SH backward (backward.cu:21-145) or precomputed colours -----------------------------
        if (p.dL_dcolors_precomp) {
            float* d = p.dL_dcolors_precomp + (size_t)i * 3;
            d[0] = dcolor.x; d[1] = dcolor.y; d[2] = dcolor.z;
        }
        if (p.shs != nullptr) {
            const uint8_t cl = p.clamped[i];
            const V3 dRGB = v3((cl & 1) ? 0.f : dcolor.x, (cl & 2) ? 0.f : dcolor.y, (cl & 4) ? 0.f : dcolor.z);
            const V3 dir_orig = v3(mx - __ldg(p.campos), my - __ldg(p.campos + 1), mz - __ldg(p.campos + 2));
            const float len = sqrtf(dot(dir_orig, dir_orig));
            const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
            // raw mode: coefficient 0 <-> _features_dc [P,1,3], coefficients 1.. <-> _features_rest [P,M-1,3]
            const bool split = kRaw && p.M > 1;
            const float* sh = kRaw ? p.shs + (size_t)i * 3 : p.shs + (size_t)i * p.M * 3;
            const float* shr = split ? p.shs_rest + ((ptrdiff_t)i * (p.M - 1) - 1) * 3 : sh;      // signed: negative shift for i == 0
            const bool want_dsh = p.dL_dshs != nullptr;      // NULL: the caller does not need dL/dSH (needs_input_grad)
            float* dsh = kRaw ? p.dL_dshs + (size_t)i * 3 : p.dL_dshs + (size_t)i * p.M * 3;
            float* dshr = split ? p.dL_dshs_rest + ((ptrdiff_t)i * (p.M - 1) - 1) * 3 : dsh;
#define SHV(k) v3(__ldg(((k) == 0 ? sh : shr) + (k) * 3), __ldg(((k) == 0 ? sh : shr) + (k) * 3 + 1), __ldg(((k) == 0 ? sh : shr) + (k) * 3 + 2))
#define DSH(k, f) { if (want_dsh) { const float f__ = (f); float* d__ = ((k) == 0 ? dsh : dshr) + (k) * 3; d__[0] = f__ * dRGB.x; d__[1] = f__ * dRGB.y; d__[2] = f__ * dRGB.z; } }
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;    // dL/d(dir) accumulated as dot(dRGB/d(dir), dRGB)
            DSH(0, G4R_SH_C0);
            if (p.D > 0) {
                DSH(1, -G4R_SH_C1 * y); DSH(2, G4R_SH_C1 * z); DSH(3, -G4R_SH_C1 * x);
                ddx += -G4R_SH_C1 * dot(SHV(3), dRGB);
                ddy += -G4R_SH_C1 * dot(SHV(1), dRGB);
                ddz += G4R_SH_C1 * dot(SHV(2), dRGB);
                if (p.D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    DSH(4, G4R_SH_C2_0 * xy); DSH(5, G4R_SH_C2_1 * yz); DSH(6, G4R_SH_C2_2 * (2.f * zz - xx - yy));
                    DSH(7, G4R_SH_C2_3 * xz); DSH(8, G4R_SH_C2_4 * (xx - yy));
                    const float s4 = dot(SHV(4), dRGB), s5 = dot(SHV(5), dRGB), s6 = dot(SHV(6), dRGB), s7 = dot(SHV(7), dRGB), s8 = dot(SHV(8), dRGB);
                    ddx += G4R_SH_C2_0 * y * s4 + G4R_SH_C2_2 * 2.f * -x * s6 + G4R_SH_C2_3 * z * s7 + G4R_SH_C2_4 * 2.f * x * s8;
                    ddy += G4R_SH_C2_0 * x * s4 + G4R_SH_C2_1 * z * s5 + G4R_SH_C2_2 * 2.f * -y * s6 + G4R_SH_C2_4 * 2.f * -y * s8;
                    ddz += G4R_SH_C2_1 * y * s5 + G4R_SH_C2_2 * 2.f * 2.f * z * s6 + G4R_SH_C2_3 * x * s7;
                    if (p.D > 2) {
                        DSH(9, G4R_SH_C3_0 * y * (3.f * xx - yy)); DSH(10, G4R_SH_C3_1 * xy * z);
                        DSH(11, G4R_SH_C3_2 * y * (4.f * zz - xx - yy)); DSH(12, G4R_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy));
                        DSH(13, G4R_SH_C3_4 * x * (4.f * zz - xx - yy)); DSH(14, G4R_SH_C3_5 * z * (xx - yy));
                        DSH(15, G4R_SH_C3_6 * x * (xx - 3.f * yy));
                        const float s9 = dot(SHV(9), dRGB), s10 = dot(SHV(10), dRGB), s11 = dot(SHV(11), dRGB), s12 = dot(SHV(12), dRGB),
                                    s13 = dot(SHV(13), dRGB), s14 = dot(SHV(14), dRGB), s15 = dot(SHV(15), dRGB);
                        ddx += G4R_SH_C3_0 * s9 * 3.f * 2.f * xy + G4R_SH_C3_1 * s10 * yz + G4R_SH_C3_2 * s11 * -2.f * xy +
                               G4R_SH_C3_3 * s12 * -3.f * 2.f * xz + G4R_SH_C3_4 * s13 * (-3.f * xx + 4.f * zz - yy) +
                               G4R_SH_C3_5 * s14 * 2.f * xz + G4R_SH_C3_6 * s15 * 3.f * (xx - yy);
                        ddy += G4R_SH_C3_0 * s9 * 3.f * (xx - yy) + G4R_SH_C3_1 * s10 * xz + G4R_SH_C3_2 * s11 * (-3.f * yy + 4.f * zz - xx) +
                               G4R_SH_C3_3 * s12 * -3.f * 2.f * yz + G4R_SH_C3_4 * s13 * -2.f * xy + G4R_SH_C3_5 * s14 * -2.f * yz +
                               G4R_SH_C3_6 * s15 * -3.f * 2.f * xy;
                        ddz += G4R_SH_C3_1 * s10 * xy + G4R_SH_C3_2 * s11 * 4.f * 2.f * yz + G4R_SH_C3_3 * s12 * 3.f * (2.f * zz - xx - yy) +
                               G4R_SH_C3_4 * s13 * 4.f * 2.f * xz + G4R_SH_C3_5 * s14 * (xx - yy);
                    }
                }
            }
            if (want_dsh) for (int k = K * 3; k < p.M * 3; ++k) dshr[k] = 0.f;    // inactive coefficients (reference: torch::zeros); K >= 1
#undef SHV
#undef DSH
            // d normalize(v)/dv applied to dL/d(dir) (auxiliary.h:109-120)
            const float sum2 = dot(dir_orig, dir_orig);
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            const V3 dm = v3(((sum2 - dir_orig.x * dir_orig.x) * ddx - dir_orig.y * dir_orig.x * ddy - dir_orig.z * dir_orig.x * ddz) * invsum32,
                             (-dir_orig.x * dir_orig.y * ddx + (sum2 - dir_orig.y * dir_orig.y) * ddy - dir_orig.z * dir_orig.y * ddz) * invsum32,
                             (-dir_orig.x * dir_orig.z * ddx - dir_orig.y * dir_orig.z * ddy + (sum2 - dir_orig.z * dir_orig.z) * ddz) * invsum32);
            dmean.x += dm.x; dmean.y += dm.y; dmean.z += dm.z;
            tau[0] -= dm.x; tau[1] -= dm.y; tau[2] -= dm.z;       // translation-only approximation (backward.cu:141-143)
        } else if (p.dL_dshs) {
            float* d = p.dL_dshs + (size_t)i * p.M * 3;
            for (int k = 0; k < p.M * 3; ++k) d[k] = 0.f;
        }

        // ---- 3D covariance -> scale / rotation (backward.cu:350-413), or hand dL/dcov3D out ----------------
        if (has_scale) {
            // dL/dSigma as a symmetric matrix (off-diagonals were accumulated doubled)
            const float S00 = dcov[0], S01 = 0.5f * dcov[1], S02 = 0.5f * dcov[2], S11 = dcov[3], S12 = 0.5f * dcov[4], S22 = dcov[5];
            // N[a][k] = 2 * s_k * sum_b dSigma[a][b] * R[b][k]  (= reference dL_dMt[k][a]); dL/ds_k = sum_a R[a][k]*N[a][k]/1
            float N[3][3];
            const float sk[3] = {sx, sy, sz};
            const float dS[3][3] = {{S00, S01, S02}, {S01, S11, S12}, {S02, S12, S22}};
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int k = 0; k < 3; ++k) N[a][k] = 2.f * sk[k] * (dS[a][0] * R[0][k] + dS[a][1] * R[1][k] + dS[a][2] * R[2][k]);
            if (p.dL_dscales || (slot >= 0 && p.dL_dds)) {
                float ds[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) ds[k] = R[0][k] * N[0][k] + R[1][k] * N[1][k] + R[2][k] * N[2][k];
                if (slot >= 0 && p.dL_dds) { float* d = p.dL_dds + (size_t)slot * 3; d[0] = ds[0]; d[1] = ds[1]; d[2] = ds[2]; }
                if (!p.dL_dscales) {
                } else if (kRaw) {                                            // d exp(x) = exp(x) dx
                    if (p.scale_dim == 1) p.dL_dscales[i] = (ds[0] + ds[1] + ds[2]) * act_s[0];
                    else { float* d = p.dL_dscales + (size_t)i * 3; d[0] = ds[0] * act_s[0]; d[1] = ds[1] * act_s[1]; d[2] = ds[2] * act_s[2]; }
                } else {
                    float* d = p.dL_dscales + (size_t)i * 3;
                    d[0] = ds[0]; d[1] = ds[1]; d[2] = ds[2];
                }
            }
            // G[k][a] = dL/dR entries scaled by s_k: reference dL_dMt[k] *= s_k  -> G[k][a] = s_k * N[a][k]
            float Gm[3][3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int a = 0; a < 3; ++a) Gm[k][a] = sk[k] * N[a][k];
            if (p.dL_drotations || (slot >= 0 && p.dL_ddr)) {
                float4 dq;
                dq.x = 2.f * qz * (Gm[0][1] - Gm[1][0]) + 2.f * qy * (Gm[2][0] - Gm[0][2]) + 2.f * qx * (Gm[1][2] - Gm[2][1]);
                dq.y = 2.f * qy * (Gm[1][0] + Gm[0][1]) + 2.f * qz * (Gm[2][0] + Gm[0][2]) + 2.f * qr * (Gm[1][2] - Gm[2][1]) - 4.f * qx * (Gm[2][2] + Gm[1][1]);
                dq.z = 2.f * qx * (Gm[1][0] + Gm[0][1]) + 2.f * qr * (Gm[2][0] - Gm[0][2]) + 2.f * qz * (Gm[1][2] + Gm[2][1]) - 4.f * qy * (Gm[2][2] + Gm[0][0]);
                dq.w = 2.f * qr * (Gm[0][1] - Gm[1][0]) + 2.f * qx * (Gm[2][0] + Gm[0][2]) + 2.f * qy * (Gm[1][2] + Gm[2][1]) - 4.f * qz * (Gm[1][1] + Gm[0][0]);
                if (slot >= 0 && p.dL_ddr) reinterpret_cast<float4*>(p.dL_ddr)[slot] = dq;      // w.r.t. the offset = w.r.t. the sum
                if (kRaw) {                                                   // d normalize(x) = (g - n (n.g)) / |x|, n = x / |x|
                    const float qg = nr * dq.x + nx * dq.y + ny * dq.z + nz * dq.w;
                    const float inv = 1.0f / qnorm;
                    dq.x = (dq.x - nr * qg) * inv; dq.y = (dq.y - nx * qg) * inv; dq.z = (dq.z - ny * qg) * inv; dq.w = (dq.w - nz * qg) * inv;
                }
                if (p.dL_drotations) reinterpret_cast<float4*>(p.dL_drotations)[i] = dq;
            }
        } else if (p.dL_dcov3D) {
            float* d = p.dL_dcov3D + (size_t)i * 6;
#pragma unroll
            for (int k = 0; k < 6; ++k) d[k] = dcov[k];
        }

        const size_t o3 = (size_t)i * 3;
        if (p.dL_dmeans3D) { p.dL_dmeans3D[o3] = dmean.x; p.dL_dmeans3D[o3 + 1] = dmean.y; p.dL_dmeans3D[o3 + 2] = dmean.z; }
        if (slot >= 0 && p.dL_ddx) { float* d = p.dL_ddx + (size_t)slot * 3; d[0] = dmean.x; d[1] = dmean.y; d[2] = dmean.z; }
        if (p.dL_dmeans2D) { p.dL_dmeans2D[o3] = dmean2D_x; p.dL_dmeans2D[o3 + 1] = dmean2D_y; p.dL_dmeans2D[o3 + 2] = 0.f; }
        if (p.dL_dopacity) {
            if (kRaw) {                                                      // d sigmoid(x) = o (1 - o) dx; o as the forward stored it
                const float o = __ldg(reinterpret_cast<const float*>(p.rec + (size_t)i * 3 + 1) + 1);
                p.dL_dopacity[i] = dopacity * o * (1.0f - o);
            } else {
                p.dL_dopacity[i] = dopacity;
            }
        }
    }

    // ---- pose gradient: warp shuffle -> CTA -> 6 atomics ------------------------------------------------------
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float s = tau[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) s_tau[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < G4R_BLOCK / 32; ++w) s += s_tau[w][threadIdx.x];
        if (s != 0.f) atomicAdd(p.dL_dtau + threadIdx.x, s);
    }
}

int launch_gaussian_backward(const G4RFrame& f, const G4RGaussians& g, const int32_t* radii, const void* geom,
                             const float* acc, const G4RBackwardIO& io, cudaStream_t s) {
    const GeomLayout gl(g.P);
    GaussBwdParams p;
    p.P = g.P; p.D = f.sh_degree; p.M = f.sh_coeffs; p.W = f.width; p.H = f.height;
    p.tan_fovx = f.tan_fovx; p.tan_fovy = f.tan_fovy;
    p.focal_x = (float)f.width / (2.0f * f.tan_fovx);
    p.focal_y = (float)f.height / (2.0f * f.tan_fovy);
    p.scale_modifier = f.scale_modifier;
    p.means3D = g.means3D; p.shs = g.shs; p.shs_rest = g.shs_rest; p.colors_precomp = g.colors_precomp; p.scales = g.scales;
    p.rotations = g.rotations; p.cov3D_precomp = g.cov3D_precomp;
    p.scale_dim = g.scale_dim == 1 ? 1 : 3;
    p.dyn_slot = g.dyn_slot; p.dx = g.dx; p.ds = g.ds; p.dr = g.dr;
    p.dL_ddx = io.dL_ddx; p.dL_dds = io.dL_dds; p.dL_ddr = io.dL_ddr;
    p.viewmatrix = f.viewmatrix; p.projmatrix = f.projmatrix; p.projmatrix_raw = f.projmatrix_raw; p.campos = f.campos;
    p.radii = radii;
    p.rec = (const float4*)((const char*)geom + gl.rec);
    p.clamped = (const uint8_t*)((const char*)geom + gl.clamped);
    p.acc = acc;
    p.dL_dmeans3D = io.dL_dmeans3D; p.dL_dmeans2D = io.dL_dmeans2D; p.dL_dopacity = io.dL_dopacity; p.dL_dshs = io.dL_dshs;
    p.dL_dshs_rest = io.dL_dshs_rest;
    p.dL_dcolors_precomp = io.dL_dcolors_precomp; p.dL_dscales = io.dL_dscales; p.dL_drotations = io.dL_drotations;
    p.dL_dcov3D = io.dL_dcov3D; p.dL_dtau = io.dL_dtau;
    g4r_stage_begin(ST_GAUSSIAN_BWD, s);
    if (g.activation == G4R_ACT_RAW) gaussian_backward_kernel<true><<<(g.P + G4R_BLOCK - 1) / G4R_BLOCK, G4R_BLOCK, 0, s>>>(p);
    else gaussian_backward_kernel<false><<<(g.P + G4R_BLOCK - 1) / G4R_BLOCK, G4R_BLOCK, 0, s>>>(p);
    g4r_stage_end(ST_GAUSSIAN_BWD, s);
    G4R_LAUNCH_OK("gaussian_backward_kernel");
    return G4R_OK;
}
