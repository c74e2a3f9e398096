// loss.cu -- the RGB-D tracking / mapping losses of 4DGS-SLAM and their image gradients in ONE kernel.
//
// The reference evaluates get_loss_tracking_rgbd (utils/slam_utils.py:57-173) and get_loss_mapping_rgbd (:252-364, the static,
// non-split branch) with ~15 image-sized torch kernels each, and autograd runs as many again on the way back.  Both losses are
//      loss = alpha * mean_{3X}( w * | m_rgb * I_ab - m_rgb * I_gt | ) + (1 - alpha) * mean_X( | m_d * D - m_d * D_gt | )
// with I_ab = exp(a) * I + b (exposure), per-pixel 0/1 masks m_rgb, m_d and weight w:
//   tracking: m_rgb = (sum_c I_gt > thr) * grad_mask * motion_mask,  w = rendered opacity (a constant for autograd: the rasterizer
//             drops its gradient, DGR/diff_gaussian_rasterization/__init__.py:108),  m_d = (D_gt > 0.01)(D_gt < 1000)(opacity > 0.95) * motion_mask
//   mapping : m_rgb = (sum_c I_gt > thr) * motion_mask,  w = 1,  m_d = (D_gt > 0.01)(D_gt < 10000) * motion_mask
// One pass produces the loss, dL/dI, dL/dD and dL/da, dL/db; the last CTA to finish turns the block sums into the final scalars.
#include "g4r_common.cuh"

struct LossParams {
    int X;                 // W * H
    int mode;              // 0 tracking, 1 mapping
    float alpha, thr;
    const float *image, *depth, *opacity, *gt_image, *gt_depth, *exp_a, *exp_b;
    const uint8_t *motion_mask, *grad_mask;
    float *dL_dimage, *dL_ddepth;
    float* out;            // [0] loss [1] dL/da [2] dL/db [3] (unused)   -- written by the last CTA
    float* sums;           // [0] sum rgb terms [1] sum depth terms [2] raw dL/da [3] raw dL/db [4] ticket (as uint)
};

static __device__ __forceinline__ float sgn(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }   // d|v|/dv, 0 at 0 like torch

__global__ void __launch_bounds__(G4R_BLOCK) slam_loss_kernel(const LossParams p) {
    __shared__ float s_red[G4R_BLOCK / 32][4];
    const float ea = p.exp_a ? expf(__ldg(p.exp_a)) : 1.0f;
    const float eb = p.exp_b ? __ldg(p.exp_b) : 0.0f;
    const float k_rgb = p.alpha / (3.0f * (float)p.X), k_d = (1.0f - p.alpha) / (float)p.X;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int pix = blockIdx.x * G4R_BLOCK + threadIdx.x; pix < p.X; pix += gridDim.x * G4R_BLOCK) {
        const float g0 = __ldg(p.gt_image + pix), g1 = __ldg(p.gt_image + p.X + pix), g2 = __ldg(p.gt_image + 2 * (size_t)p.X + pix);
        const float mm = p.motion_mask ? (p.motion_mask[pix] ? 1.0f : 0.0f) : 1.0f;
        float m = ((g0 + g1) + g2 > p.thr ? 1.0f : 0.0f) * mm;
        float w = 1.0f, md;
        const float gd = __ldg(p.gt_depth + pix);
        if (p.mode == 0) {
            const float o = __ldg(p.opacity + pix);
            if (p.grad_mask) m *= p.grad_mask[pix] ? 1.0f : 0.0f;
            w = o;
            md = (gd > 0.01f && gd < 1000.0f && o > 0.95f) ? mm : 0.0f;
        } else {
            md = (gd > 0.01f && gd < 10000.0f) ? mm : 0.0f;
        }
        const float gt[3] = {g0, g1, g2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float I = __ldg(p.image + (size_t)c * p.X + pix);
            const float diff = (ea * I + eb) * m - gt[c] * m;
            const float s = sgn(diff) * m * w;                 // d term / d I_ab
            acc[0] += w * fabsf(diff);
            acc[2] += s * ea * I;                              // d I_ab / d a = exp(a) * I
            acc[3] += s;
            p.dL_dimage[(size_t)c * p.X + pix] = k_rgb * s * ea;
        }
        const float D = __ldg(p.depth + pix);
        const float dd = D * md - gd * md;
        acc[1] += fabsf(dd);
        p.dL_ddepth[pix] = k_d * sgn(dd) * md;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = acc[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) s_red[warp][k] = v;
    }
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x < 4) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < G4R_BLOCK / 32; ++w) v += s_red[w][threadIdx.x];
        atomicAdd(p.sums + threadIdx.x, v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(reinterpret_cast<unsigned int*>(p.sums + 4), 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const volatile float* sm = p.sums;
        p.out[0] = k_rgb * sm[0] + k_d * sm[1];
        p.out[1] = k_rgb * sm[2];
        p.out[2] = k_rgb * sm[3];
        p.out[3] = 0.0f;
    }
}

extern "C" int g4r_slam_loss(const G4RLossIn* in, float* dL_dimage, float* dL_ddepth, float* out4, void* scratch32, void* stream) {
    if (!in || !dL_dimage || !dL_ddepth || !out4 || !scratch32) return g4r_set_error(G4R_EINVAL, "NULL argument");
    if (in->width <= 0 || in->height <= 0) return g4r_set_error(G4R_EINVAL, "image size %dx%d is not positive", in->width, in->height);
    if (in->mode != 0 && in->mode != 1) return g4r_set_error(G4R_EINVAL, "mode must be 0 (tracking) or 1 (mapping)");
    if (!in->image || !in->depth || !in->gt_image || !in->gt_depth || (in->mode == 0 && !in->opacity))
        return g4r_set_error(G4R_EINVAL, "image / depth / gt_image / gt_depth (and opacity for tracking) are required");
    cudaStream_t s = (cudaStream_t)stream;
    LossParams p;
    p.X = in->width * in->height; p.mode = in->mode; p.alpha = in->alpha; p.thr = in->rgb_boundary_threshold;
    p.image = in->image; p.depth = in->depth; p.opacity = in->opacity; p.gt_image = in->gt_image; p.gt_depth = in->gt_depth;
    p.exp_a = in->exposure_a; p.exp_b = in->exposure_b; p.motion_mask = in->motion_mask; p.grad_mask = in->grad_mask;
    p.dL_dimage = dL_dimage; p.dL_ddepth = dL_ddepth; p.out = out4; p.sums = (float*)scratch32;
    G4R_CUDA_OK(cudaMemsetAsync(scratch32, 0, 32, s));
    int blocks = (p.X + G4R_BLOCK - 1) / G4R_BLOCK;
    if (blocks > 148 * 8) blocks = 148 * 8;
    slam_loss_kernel<<<blocks, G4R_BLOCK, 0, s>>>(p);
    G4R_LAUNCH_OK("slam_loss_kernel");
    return G4R_OK;
}
