// warp.cu -- the control-node warp of 4DGS-SLAM's deformation step in two kernels.
//
// Reference: ControlNodeWarp.forward (utils/time_utils.py:1192-1275) with cal_nn_weight (:981-1015) for the configuration the SLAM
// runs (arguments.py:107-125: K = 3, node_num = 512, hyper_dim = 0, skinning off, with_node_weight, d_rot_as_res, local_frame; the
// other values of d_rot_as_res / local_frame are covered too).  Given the per-node outputs of the deformation MLP
// (d_xyz, d_rotation, d_scaling, local_rotation: [M, *]) it moves every dynamic Gaussian by a blend of its K nearest nodes:
//     nn_dist, nn_idx = knn_points(x, nodes[:, :3], K)                        (pytorch3d, squared distances ascending)   :998
//     a_k = exp(-nn_dist_k / (2 exp(log_radius[idx_k])^2)) * sigmoid(weight_logit[idx_k]) + 1e-7 ;  w_k = a_k / sum a    :1001-1007
//     translate = (sum_k w_k (R(local_rotation[idx_k] + (1,0,0,0)) (x - n_k) + n_k + d_xyz[idx_k]) - x) * motion_mask    :1208-1214
//                 (local_frame off: sum_k w_k d_xyz[idx_k] * motion_mask                                                  :1216)
//     rotation  = sum_k w_k d_rotation[idx_k] * motion_mask            (d_rot_as_res; else ((sum w (rot + b)) - b) m + b) :1252 / :1232
//     scale     = sum_k w_k d_scaling[idx_k] * motion_mask                                                               :1253 / :1247
// In torch that is pytorch3d's knn kernel plus ~35 element-wise / gather / einsum kernels forward and ~60 backward over [N, K, *]
// intermediates, for N up to a few 10^5 and M = 512: launch-bound.  Here: one forward kernel (node table in shared memory, top-K in
// registers, everything else fused) and one backward kernel (per-node gradient accumulators in shared memory, flushed once per CTA)
// plus a per-node epilogue that applies the quaternion -> matrix, exp and sigmoid chain rules.
// x and the node positions are constants for autograd exactly like in the reference (x.detach() :1196, nodes[..., :3].detach() :994).
#include "g4r_common.cuh"
#include <atomic>
#include <cfloat>

#define WARP_KMAX 8
#define WARP_CHUNK 2048                 // node positions staged per pass: 24 KB of shared memory
#define WARP_ACC 24                     // per-node accumulators: 3 d_xyz, 4 d_rotation, 3 d_scaling, 9 dL/dR, 1 log_radius, 1 weight_logit, 3 pad
#define WARP_ACC_SMEM_NODES 2048        // backward keeps the accumulators in shared memory up to this many nodes (192 KB)

struct WarpParams {
    int N, M, K, node_stride, d_rot_as_res, local_frame;
    const float *x, *nodes, *log_radius, *weight_logit, *d_xyz, *d_rotation, *d_scaling, *local_rotation, *motion_mask;
    // forward outputs / saved state
    float *translate, *rotation, *scale, *nn_dist, *nn_weight;
    int32_t* nn_idx;
    // backward
    const float *g_translate, *g_rotation, *g_scale;
    float* acc;                          // [M][WARP_ACC], zeroed
};

// utils/time_utils.py:115-132 quaternion_to_matrix (real part first, scaled by 2 / |q|^2)
static __device__ __forceinline__ void quat_to_mat(float r, float i, float j, float k, float* R) {
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    R[0] = 1.0f - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r);        R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r);        R[4] = 1.0f - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r);        R[7] = two_s * (j * k + i * r);        R[8] = 1.0f - two_s * (i * i + j * j);
}

__global__ void __launch_bounds__(G4R_BLOCK) node_warp_forward_kernel(const WarpParams p) {
    __shared__ float s_node[WARP_CHUNK][3];          // node positions; every thread reads the same entry (broadcast)
    const int n = blockIdx.x * G4R_BLOCK + threadIdx.x;
    const bool live = n < p.N;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) { px = __ldg(p.x + 3 * (size_t)n); py = __ldg(p.x + 3 * (size_t)n + 1); pz = __ldg(p.x + 3 * (size_t)n + 2); }
    float bd[WARP_KMAX];
    int bi[WARP_KMAX];
#pragma unroll
    for (int j = 0; j < WARP_KMAX; ++j) { bd[j] = FLT_MAX; bi[j] = -1; }
    float worst = FLT_MAX;                           // bd[K - 1]: most nodes fail this one comparison
    for (int base = 0; base < p.M; base += WARP_CHUNK) {
        const int cnt = min(WARP_CHUNK, p.M - base);
        __syncthreads();
        for (int m = threadIdx.x; m < cnt; m += G4R_BLOCK) {
            const float* nd = p.nodes + (size_t)(base + m) * p.node_stride;
            s_node[m][0] = __ldg(nd); s_node[m][1] = __ldg(nd + 1); s_node[m][2] = __ldg(nd + 2);
        }
        __syncthreads();
        if (live) {
            for (int m = 0; m < cnt; ++m) {
                // ((x - n) ** 2).sum(-1): three rounded squares added left to right, no contraction -- the restatement's arithmetic
                const float dx = __fsub_rn(px, s_node[m][0]), dy = __fsub_rn(py, s_node[m][1]), dz = __fsub_rn(pz, s_node[m][2]);
                float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d < worst) {
                    int id = base + m;
#pragma unroll
                    for (int j = 0; j < WARP_KMAX; ++j) {            // strict <: on equal distances the lower node index stays in front
                        if (j < p.K && d < bd[j]) {
                            const float td = bd[j]; const int ti = bi[j];
                            bd[j] = d; bi[j] = id; d = td; id = ti;
                        }
                        if (j == p.K - 1) worst = bd[j];
                    }
                }
            }
        }
    }
    if (!live) return;
    // weights (cal_nn_weight, gs_kernel branch)
    float a[WARP_KMAX], A = 0.f;
#pragma unroll
    for (int k = 0; k < WARP_KMAX; ++k) {
        a[k] = 0.f;
        if (k < p.K && bi[k] >= 0) {
            const float r = expf(__ldg(p.log_radius + bi[k]));
            float e = expf(-bd[k] / (2.0f * r * r));
            if (p.weight_logit) e *= g4r_sigmoid(__ldg(p.weight_logit + bi[k]));
            a[k] = e + 1e-7f;
            A += a[k];
        }
    }
    const float mask = p.motion_mask ? __ldg(p.motion_mask + n) : 1.0f;
    float T[3] = {0.f, 0.f, 0.f}, Q[4] = {0.f, 0.f, 0.f, 0.f}, S[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WARP_KMAX; ++k) {
        if (k < p.K && bi[k] >= 0) {
            const int id = bi[k];
            const float w = a[k] / A;
            p.nn_idx[(size_t)n * p.K + k] = id;
            p.nn_dist[(size_t)n * p.K + k] = bd[k];
            p.nn_weight[(size_t)n * p.K + k] = w;
            const float tx = __ldg(p.d_xyz + 3 * (size_t)id), ty = __ldg(p.d_xyz + 3 * (size_t)id + 1), tz = __ldg(p.d_xyz + 3 * (size_t)id + 2);
            if (p.local_frame) {
                const float* nd = p.nodes + (size_t)id * p.node_stride;
                const float nx = __ldg(nd), ny = __ldg(nd + 1), nz = __ldg(nd + 2);
                const float4 lq = __ldg(reinterpret_cast<const float4*>(p.local_rotation) + id);
                float R[9];
                quat_to_mat(lq.x + 1.0f, lq.y, lq.z, lq.w, R);
                const float ex = px - nx, ey = py - ny, ez = pz - nz;
                T[0] += w * (R[0] * ex + R[1] * ey + R[2] * ez + nx + tx);
                T[1] += w * (R[3] * ex + R[4] * ey + R[5] * ez + ny + ty);
                T[2] += w * (R[6] * ex + R[7] * ey + R[8] * ez + nz + tz);
            } else {
                T[0] += w * tx; T[1] += w * ty; T[2] += w * tz;
            }
            const float4 q = __ldg(reinterpret_cast<const float4*>(p.d_rotation) + id);
            const float bias = p.d_rot_as_res ? 0.0f : 1.0f;
            Q[0] += w * (q.x + bias); Q[1] += w * q.y; Q[2] += w * q.z; Q[3] += w * q.w;
            S[0] += w * __ldg(p.d_scaling + 3 * (size_t)id); S[1] += w * __ldg(p.d_scaling + 3 * (size_t)id + 1);
            S[2] += w * __ldg(p.d_scaling + 3 * (size_t)id + 2);
        } else if (k < p.K) {                                       // fewer than K nodes exist
            p.nn_idx[(size_t)n * p.K + k] = -1; p.nn_dist[(size_t)n * p.K + k] = FLT_MAX; p.nn_weight[(size_t)n * p.K + k] = 0.f;
        }
    }
    if (p.local_frame) { T[0] -= px; T[1] -= py; T[2] -= pz; }
    p.translate[3 * (size_t)n] = T[0] * mask; p.translate[3 * (size_t)n + 1] = T[1] * mask; p.translate[3 * (size_t)n + 2] = T[2] * mask;
    if (p.d_rot_as_res) {
        reinterpret_cast<float4*>(p.rotation)[n] = make_float4(Q[0] * mask, Q[1] * mask, Q[2] * mask, Q[3] * mask);
    } else {
        reinterpret_cast<float4*>(p.rotation)[n] = make_float4((Q[0] - 1.0f) * mask + 1.0f, Q[1] * mask, Q[2] * mask, Q[3] * mask);
    }
    p.scale[3 * (size_t)n] = S[0] * mask; p.scale[3 * (size_t)n + 1] = S[1] * mask; p.scale[3 * (size_t)n + 2] = S[2] * mask;
}

// One thread per Gaussian (grid-stride); the per-node sums live in shared memory when M is small enough, else in global memory.
template <bool kSmem>
__global__ void __launch_bounds__(G4R_BLOCK) node_warp_backward_kernel(const WarpParams p) {
    extern __shared__ float s_acc[];                 // [M][WARP_ACC] when kSmem
    float* acc = kSmem ? s_acc : p.acc;
    if (kSmem) {
        for (int i = threadIdx.x; i < p.M * WARP_ACC; i += G4R_BLOCK) s_acc[i] = 0.f;
        __syncthreads();
    }
    for (int n = blockIdx.x * G4R_BLOCK + threadIdx.x; n < p.N; n += gridDim.x * G4R_BLOCK) {
        const float mask = p.motion_mask ? __ldg(p.motion_mask + n) : 1.0f;
        const float gT[3] = {__ldg(p.g_translate + 3 * (size_t)n) * mask, __ldg(p.g_translate + 3 * (size_t)n + 1) * mask,
                             __ldg(p.g_translate + 3 * (size_t)n + 2) * mask};
        const float4 gq4 = __ldg(reinterpret_cast<const float4*>(p.g_rotation) + n);
        const float gQ[4] = {gq4.x * mask, gq4.y * mask, gq4.z * mask, gq4.w * mask};
        const float gS[3] = {__ldg(p.g_scale + 3 * (size_t)n) * mask, __ldg(p.g_scale + 3 * (size_t)n + 1) * mask,
                             __ldg(p.g_scale + 3 * (size_t)n + 2) * mask};
        const float px = __ldg(p.x + 3 * (size_t)n), py = __ldg(p.x + 3 * (size_t)n + 1), pz = __ldg(p.x + 3 * (size_t)n + 2);
        float w[WARP_KMAX], gw[WARP_KMAX], e[WARP_KMAX], sg[WARP_KMAX], d[WARP_KMAX], ir2[WARP_KMAX];
        int id[WARP_KMAX];
        float A = 0.f, wgw = 0.f;
#pragma unroll
        for (int k = 0; k < WARP_KMAX; ++k) {
            id[k] = -1; w[k] = gw[k] = e[k] = sg[k] = d[k] = ir2[k] = 0.f;
            if (k >= p.K) continue;
            id[k] = p.nn_idx[(size_t)n * p.K + k];
            if (id[k] < 0) continue;
            const int m = id[k];
            d[k] = p.nn_dist[(size_t)n * p.K + k];
            w[k] = p.nn_weight[(size_t)n * p.K + k];
            const float r = expf(__ldg(p.log_radius + m));
            ir2[k] = 1.0f / (r * r);
            e[k] = expf(-d[k] * 0.5f * ir2[k]);
            sg[k] = p.weight_logit ? g4r_sigmoid(__ldg(p.weight_logit + m)) : 1.0f;
            A += e[k] * sg[k] + 1e-7f;
            // dL/dw_k = g . (the blended quantity of node k), and the direct sums into the node attributes
            float* row = acc + (size_t)m * WARP_ACC;
            const float tx = __ldg(p.d_xyz + 3 * (size_t)m), ty = __ldg(p.d_xyz + 3 * (size_t)m + 1), tz = __ldg(p.d_xyz + 3 * (size_t)m + 2);
            float g = 0.f;
            if (p.local_frame) {
                const float* nd = p.nodes + (size_t)m * p.node_stride;
                const float nx = __ldg(nd), ny = __ldg(nd + 1), nz = __ldg(nd + 2);
                const float4 lq = __ldg(reinterpret_cast<const float4*>(p.local_rotation) + m);
                float R[9];
                quat_to_mat(lq.x + 1.0f, lq.y, lq.z, lq.w, R);
                const float ex = px - nx, ey = py - ny, ez = pz - nz;
                g = gT[0] * (R[0] * ex + R[1] * ey + R[2] * ez + nx + tx) + gT[1] * (R[3] * ex + R[4] * ey + R[5] * ez + ny + ty) +
                    gT[2] * (R[6] * ex + R[7] * ey + R[8] * ez + nz + tz);
                const float ev[3] = {ex, ey, ez};
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) atomicAdd(row + 10 + 3 * a + b, w[k] * gT[a] * ev[b]);
            } else {
                g = gT[0] * tx + gT[1] * ty + gT[2] * tz;
            }
            const float4 q = __ldg(reinterpret_cast<const float4*>(p.d_rotation) + m);
            const float bias = p.d_rot_as_res ? 0.0f : 1.0f;
            g += gQ[0] * (q.x + bias) + gQ[1] * q.y + gQ[2] * q.z + gQ[3] * q.w;
            g += gS[0] * __ldg(p.d_scaling + 3 * (size_t)m) + gS[1] * __ldg(p.d_scaling + 3 * (size_t)m + 1) + gS[2] * __ldg(p.d_scaling + 3 * (size_t)m + 2);
            gw[k] = g;
            wgw += w[k] * g;
#pragma unroll
            for (int a = 0; a < 3; ++a) atomicAdd(row + a, w[k] * gT[a]);
#pragma unroll
            for (int a = 0; a < 4; ++a) atomicAdd(row + 3 + a, w[k] * gQ[a]);
#pragma unroll
            for (int a = 0; a < 3; ++a) atomicAdd(row + 7 + a, w[k] * gS[a]);
        }
        // through w_k = a_k / A, a_k = e_k s_k + 1e-7, e_k = exp(-d_k / (2 r^2)), r = exp(log_radius), s = sigmoid(weight_logit)
#pragma unroll
        for (int k = 0; k < WARP_KMAX; ++k) {
            if (k >= p.K || id[k] < 0) continue;
            const float ga = (gw[k] - wgw) / A;
            float* row = acc + (size_t)id[k] * WARP_ACC;
            atomicAdd(row + 19, ga * sg[k] * e[k] * d[k] * ir2[k]);            // d(-d / (2 r^2)) / d log r = d / r^2
            if (p.weight_logit) atomicAdd(row + 20, ga * e[k] * sg[k] * (1.0f - sg[k]));
        }
    }
    if (kSmem) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.M * WARP_ACC; i += G4R_BLOCK) {
            const float v = s_acc[i];
            if (v != 0.f) atomicAdd(p.acc + i, v);
        }
    }
}

// per node: accumulators -> the gradients of the node tensors (quaternion_to_matrix chain rule for local_rotation)
__global__ void __launch_bounds__(G4R_BLOCK) node_warp_epilogue_kernel(const WarpParams p, float* g_d_xyz, float* g_d_rotation, float* g_d_scaling,
                                                                        float* g_local_rotation, float* g_log_radius, float* g_weight_logit) {
    const int m = blockIdx.x * G4R_BLOCK + threadIdx.x;
    if (m >= p.M) return;
    const float* row = p.acc + (size_t)m * WARP_ACC;
    if (g_d_xyz) { g_d_xyz[3 * m] = row[0]; g_d_xyz[3 * m + 1] = row[1]; g_d_xyz[3 * m + 2] = row[2]; }
    if (g_d_rotation) { g_d_rotation[4 * m] = row[3]; g_d_rotation[4 * m + 1] = row[4]; g_d_rotation[4 * m + 2] = row[5]; g_d_rotation[4 * m + 3] = row[6]; }
    if (g_d_scaling) { g_d_scaling[3 * m] = row[7]; g_d_scaling[3 * m + 1] = row[8]; g_d_scaling[3 * m + 2] = row[9]; }
    if (g_log_radius) g_log_radius[m] = row[19];
    if (g_weight_logit) g_weight_logit[m] = row[20];
    if (g_local_rotation) {
        float gq[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.local_frame) {
            const float* G = row + 10;                                  // dL/dR, row-major
            const float r = p.local_rotation[4 * m] + 1.0f, i = p.local_rotation[4 * m + 1], j = p.local_rotation[4 * m + 2], k = p.local_rotation[4 * m + 3];
            const float s2 = 2.0f / (r * r + i * i + j * j + k * k);
            // R = I + s2 U(q): dL/dq_c = s2 sum G dU/dq_c - s2^2 q_c sum G U
            const float W = G[0] * -(j * j + k * k) + G[1] * (i * j - k * r) + G[2] * (i * k + j * r) + G[3] * (i * j + k * r) + G[4] * -(i * i + k * k) +
                            G[5] * (j * k - i * r) + G[6] * (i * k - j * r) + G[7] * (j * k + i * r) + G[8] * -(i * i + j * j);
            const float dr = -k * G[1] + j * G[2] + k * G[3] - i * G[5] - j * G[6] + i * G[7];
            const float di = j * (G[1] + G[3]) + k * (G[2] + G[6]) - 2.0f * i * (G[4] + G[8]) + r * (G[7] - G[5]);
            const float dj = -2.0f * j * (G[0] + G[8]) + i * (G[1] + G[3]) + r * (G[2] - G[6]) + k * (G[5] + G[7]);
            const float dk = -2.0f * k * (G[0] + G[4]) + r * (G[3] - G[1]) + i * (G[2] + G[6]) + j * (G[5] + G[7]);
            const float c = s2 * s2 * W;
            gq[0] = s2 * dr - c * r; gq[1] = s2 * di - c * i; gq[2] = s2 * dj - c * j; gq[3] = s2 * dk - c * k;
        }
        g_local_rotation[4 * m] = gq[0]; g_local_rotation[4 * m + 1] = gq[1]; g_local_rotation[4 * m + 2] = gq[2]; g_local_rotation[4 * m + 3] = gq[3];
    }
}

static int warp_params(const G4RWarpIn* in, WarpParams* p) {
    if (!in) return g4r_set_error(G4R_EINVAL, "NULL argument");
    if (in->N < 0 || in->M <= 0) return g4r_set_error(G4R_EINVAL, "N = %d, M = %d: need N >= 0 and M > 0", in->N, in->M);
    if (in->K < 1 || in->K > WARP_KMAX) return g4r_set_error(G4R_EINVAL, "K = %d is outside [1, %d]", in->K, WARP_KMAX);
    if (in->node_stride < 3) return g4r_set_error(G4R_EINVAL, "node_stride = %d < 3", in->node_stride);
    if ((in->N > 0 && !in->x) || !in->nodes || !in->log_radius || !in->d_xyz || !in->d_rotation || !in->d_scaling)
        return g4r_set_error(G4R_EINVAL, "x, nodes, log_radius, d_xyz, d_rotation, d_scaling are required");
    if (in->local_frame && !in->local_rotation) return g4r_set_error(G4R_EINVAL, "local_frame needs local_rotation");
    p->N = in->N; p->M = in->M; p->K = in->K; p->node_stride = in->node_stride; p->d_rot_as_res = in->d_rot_as_res; p->local_frame = in->local_frame;
    p->x = in->x; p->nodes = in->nodes; p->log_radius = in->log_radius; p->weight_logit = in->weight_logit; p->d_xyz = in->d_xyz;
    p->d_rotation = in->d_rotation; p->d_scaling = in->d_scaling; p->local_rotation = in->local_rotation; p->motion_mask = in->motion_mask;
    return G4R_OK;
}

extern "C" int g4r_warp_forward(const G4RWarpIn* in, float* translate, float* rotation, float* scale, int32_t* nn_idx, float* nn_dist, float* nn_weight,
                                void* stream) {
    WarpParams p = {};
    const int rc = warp_params(in, &p);
    if (rc != G4R_OK) return rc;
    if (p.N == 0) return G4R_OK;
    if (!translate || !rotation || !scale || !nn_idx || !nn_dist || !nn_weight) return g4r_set_error(G4R_EINVAL, "NULL output");
    p.translate = translate; p.rotation = rotation; p.scale = scale; p.nn_idx = nn_idx; p.nn_dist = nn_dist; p.nn_weight = nn_weight;
    node_warp_forward_kernel<<<(p.N + G4R_BLOCK - 1) / G4R_BLOCK, G4R_BLOCK, 0, (cudaStream_t)stream>>>(p);
    G4R_LAUNCH_OK("node_warp_forward_kernel");
    return G4R_OK;
}

extern "C" size_t g4r_warp_scratch_bytes(int32_t M) { return M > 0 ? (size_t)M * WARP_ACC * sizeof(float) : 0; }

extern "C" int g4r_warp_backward(const G4RWarpIn* in, const int32_t* nn_idx, const float* nn_dist, const float* nn_weight, const float* dL_dtranslate,
                                 const float* dL_drotation, const float* dL_dscale, float* dL_dd_xyz, float* dL_dd_rotation, float* dL_dd_scaling,
                                 float* dL_dlocal_rotation, float* dL_dlog_radius, float* dL_dweight_logit, void* scratch, void* stream) {
    WarpParams p = {};
    const int rc = warp_params(in, &p);
    if (rc != G4R_OK) return rc;
    if (!scratch || (p.N > 0 && (!nn_idx || !nn_dist || !nn_weight || !dL_dtranslate || !dL_drotation || !dL_dscale)))
        return g4r_set_error(G4R_EINVAL, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    p.nn_idx = const_cast<int32_t*>(nn_idx); p.nn_dist = const_cast<float*>(nn_dist); p.nn_weight = const_cast<float*>(nn_weight);
    p.g_translate = dL_dtranslate; p.g_rotation = dL_drotation; p.g_scale = dL_dscale; p.acc = (float*)scratch;
    G4R_CUDA_OK(cudaMemsetAsync(scratch, 0, g4r_warp_scratch_bytes(p.M), s));
    if (p.N > 0) {
        int blocks = (p.N + G4R_BLOCK - 1) / G4R_BLOCK;
        if (p.M <= WARP_ACC_SMEM_NODES) {
            const size_t smem = (size_t)p.M * WARP_ACC * sizeof(float);
            static std::atomic<bool> configured{false};
            if (!configured.load()) {
                G4R_CUDA_OK(cudaFuncSetAttribute(node_warp_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 WARP_ACC_SMEM_NODES * WARP_ACC * (int)sizeof(float)));
                configured.store(true);
            }
            if (blocks > 148 * 2) blocks = 148 * 2;           // every CTA flushes M * 24 sums once: keep the grid small
            node_warp_backward_kernel<true><<<blocks, G4R_BLOCK, smem, s>>>(p);
        } else {
            if (blocks > 148 * 8) blocks = 148 * 8;
            node_warp_backward_kernel<false><<<blocks, G4R_BLOCK, 0, s>>>(p);
        }
        G4R_LAUNCH_OK("node_warp_backward_kernel");
    }
    node_warp_epilogue_kernel<<<(p.M + G4R_BLOCK - 1) / G4R_BLOCK, G4R_BLOCK, 0, s>>>(p, dL_dd_xyz, dL_dd_rotation, dL_dd_scaling, dL_dlocal_rotation,
                                                                                     dL_dlog_radius, dL_dweight_logit);
    G4R_LAUNCH_OK("node_warp_epilogue_kernel");
    return G4R_OK;
}
