// g4r_torch.cpp -- the host side of the standard rasterizer op in C++ (torch::autograd::Function over the C ABI of include/g4r.h).
//
// The reference's host side is a Python autograd.Function over a compiled torch extension
// (DGR/diff_gaussian_rasterization/__init__.py:48-171 over DGR/rasterize_points.cu:35-236).  This repo's is
// diff_gaussian_rasterization/__init__.py over ctypes; measured on B200 that Python costs ~0.18 ms of host time per fwd+bwd (a dozen
// and a half torch.empty calls, two ctypes structs each way, and the autograd engine re-entering the interpreter from its worker
// thread for the backward), which is what bounds small scenes, the tracking loop and the end-to-end bench loop (host 0.78 ms per
// step vs 0.62 ms of kernels).  This file is the same logic -- argument checks, buffer allocation, the two-phase forward with the
// speculative capacity and its re-run, saved state, optional gradient outputs -- with no interpreter in the loop: one pybind call
// in, C++ backward called straight from the engine's thread.  It only calls the C ABI; no kernels live here.
//
// Scope: the reference surface (_RasterizeGaussians).  debug=True, CUDA-graph capture, the raw / fused API, the sharded render and
// the inspection hooks stay on the Python path, which remains the specification (tests run both and compare bit for bit).
#include <torch/extension.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <mutex>
#include <unordered_map>

#include "g4r.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

void check(int64_t rc) {
    if (rc < 0) throw std::runtime_error(g4r_last_error());
}

// one native context (pinned word + event) per host thread and device, like _context() in __init__.py
G4RContext* context_for(int device) {
    thread_local std::unordered_map<int, G4RContext*> table;
    auto it = table.find(device);
    if (it != table.end()) return it->second;
    G4RContext* ctx = nullptr;
    check(g4r_context_create(&ctx));
    table[device] = ctx;
    return ctx;
}

thread_local int64_t t_last_num_rendered = -1;

Tensor dev_f32(const Tensor& t, const c10::Device& device) {          // _dev_f32
    if (t.scalar_type() == at::kFloat && t.device() == device) return t.is_contiguous() ? t : t.contiguous();
    return t.to(device, at::kFloat).contiguous();
}
Tensor dev_f32_or_empty(const Tensor& t, const c10::Device& device) { return t.numel() ? dev_f32(t, device) : t; }
const float* fptr(const Tensor& t) { return t.defined() && t.numel() ? t.data_ptr<float>() : nullptr; }

struct Settings {
    int64_t H, W, sh_degree;
    double tanfovx, tanfovy, scale_modifier;
    bool prefiltered;
};

G4RFrame make_frame(const Settings& s, int M, const Tensor& bg, const Tensor& view, const Tensor& proj, const Tensor& proj_raw, const Tensor& campos) {
    G4RFrame f{};
    f.width = (int32_t)s.W; f.height = (int32_t)s.H;
    f.tan_fovx = (float)s.tanfovx; f.tan_fovy = (float)s.tanfovy; f.scale_modifier = (float)s.scale_modifier;
    f.sh_degree = (int32_t)s.sh_degree; f.sh_coeffs = M; f.prefiltered = s.prefiltered ? 1 : 0;
    f.bg = bg.data_ptr<float>(); f.viewmatrix = view.data_ptr<float>(); f.projmatrix = proj.data_ptr<float>();
    f.projmatrix_raw = proj_raw.data_ptr<float>(); f.campos = campos.data_ptr<float>();
    f.tile_rank = 0; f.tile_world = 1; f.tile_row_begin = 0; f.tile_row_end = 0;
    return f;
}

class RasterizeFn : public torch::autograd::Function<RasterizeFn> {
public:
    static variable_list forward(AutogradContext* ctx, Tensor means3D, Tensor means2D, Tensor sh, Tensor colors_precomp, Tensor opacities, Tensor scales,
                                 Tensor rotations, Tensor cov3Ds_precomp, Tensor theta, Tensor rho, Tensor bg, Tensor viewmatrix, Tensor projmatrix,
                                 Tensor projmatrix_raw, Tensor campos, int64_t H, int64_t W, double tanfovx, double tanfovy, double scale_modifier,
                                 int64_t sh_degree, bool prefiltered, int64_t hint) {
        if (means3D.dim() != 2 || means3D.size(1) != 3) throw std::runtime_error("means3D must have dimensions (num_points, 3)");   // rasterize_points.cu:58-60
        if (!means3D.is_cuda())
            throw std::runtime_error("diff_gaussian_rasterization (B200-native): means3D must be a CUDA tensor; there is no CPU path");
        const c10::Device device = means3D.device();
        c10::cuda::CUDAGuard guard(device);
        const int P = (int)means3D.size(0);
        const Settings st{H, W, sh_degree, tanfovx, tanfovy, scale_modifier, prefiltered};

        Tensor m3 = dev_f32(means3D, device), op = dev_f32(opacities, device);
        Tensor shc = dev_f32_or_empty(sh, device), col = dev_f32_or_empty(colors_precomp, device), sc = dev_f32_or_empty(scales, device);
        Tensor rot = dev_f32_or_empty(rotations, device), cov = dev_f32_or_empty(cov3Ds_precomp, device);
        const int M = shc.numel() ? (int)shc.size(1) : 0;                                                          // rasterize_points.cu:87-91

        auto f32 = at::TensorOptions().dtype(at::kFloat).device(device);
        auto i32 = at::TensorOptions().dtype(at::kInt).device(device);
        auto u8 = at::TensorOptions().dtype(at::kByte).device(device);
        ctx->saved_data["P"] = (int64_t)P;
        ctx->saved_data["H"] = H; ctx->saved_data["W"] = W; ctx->saved_data["sh_degree"] = sh_degree;
        ctx->saved_data["tanfovx"] = tanfovx; ctx->saved_data["tanfovy"] = tanfovy; ctx->saved_data["scale_modifier"] = scale_modifier;
        ctx->saved_data["prefiltered"] = prefiltered;
        ctx->saved_data["opacities_shape"] = opacities.sizes().vec();
        if (P == 0) {                                                   // rasterize_points.cu:69-73,85: zero images, no kernels
            Tensor radii = at::zeros({0}, i32), n_touched = at::zeros({0}, i32);
            ctx->save_for_backward({col, m3, sc, rot, cov, radii, shc, at::empty({0}, u8), at::empty({0}, u8), at::empty({0}, u8), bg, viewmatrix,
                                    projmatrix, projmatrix_raw, campos});
            ctx->mark_non_differentiable({radii, n_touched});
            t_last_num_rendered = 0;
            return {at::zeros({3, H, W}, f32), radii, at::zeros({1, H, W}, f32), at::zeros({1, H, W}, f32), n_touched};
        }

        Tensor bgc = dev_f32(bg, device), view = dev_f32(viewmatrix, device), proj = dev_f32(projmatrix, device);
        Tensor praw = dev_f32(projmatrix_raw, device), cam = dev_f32(campos, device);
        Tensor ints = at::empty({2, P}, i32);                           // radii | n_touched
        Tensor radii = ints.select(0, 0), n_touched = ints.select(0, 1);
        Tensor geom = at::empty({(int64_t)g4r_geom_bytes(P)}, u8);
        Tensor img = at::empty({(int64_t)g4r_image_bytes((int32_t)W, (int32_t)H)}, u8);

        G4RContext* gctx = context_for(device.index());
        void* stream = (void*)at::cuda::getCurrentCUDAStream(device.index()).stream();
        const G4RFrame frame = make_frame(st, M, bgc, view, proj, praw, cam);
        G4RGaussians g{};
        g.P = P;
        g.means3D = fptr(m3); g.opacities = fptr(op); g.shs = fptr(shc); g.colors_precomp = fptr(col);
        g.scales = fptr(sc); g.rotations = fptr(rot); g.cov3D_precomp = fptr(cov);
        // phase 1 first: the projection kernel runs while the host allocates what phase 2 writes
        check(g4r_forward_project(gctx, &frame, &g, geom.data_ptr(), img.data_ptr(), radii.data_ptr<int32_t>(), n_touched.data_ptr<int32_t>(), stream));
        Tensor color = at::empty({3, H, W}, f32), depth = at::empty({1, H, W}, f32), opacity = at::empty({1, H, W}, f32);
        G4RForwardOut out{};
        out.color = color.data_ptr<float>(); out.depth = depth.data_ptr<float>(); out.opacity = opacity.data_ptr<float>();
        out.radii = radii.data_ptr<int32_t>(); out.n_touched = n_touched.data_ptr<int32_t>();
        // speculative capacity; N is checked after everything is in the stream (the reference blocks on a cudaMemcpy, rasterizer_impl.cu:284)
        int64_t cap = hint == 0 ? (int64_t)(std::max<int64_t>(hint, 4 * (int64_t)P) * 1.25) + 4096 : (int64_t)(hint * 1.25) + 4096;
        Tensor binning = at::empty({(int64_t)g4r_binning_bytes(cap)}, u8);
        {
            Tensor sort_scratch = at::empty({(int64_t)g4r_sort_scratch_bytes(cap)}, u8);     // dies here: the caching allocator recycles it stream-ordered
            check(g4r_forward_render(gctx, &frame, &g, geom.data_ptr(), img.data_ptr(), binning.data_ptr(), sort_scratch.data_ptr(), cap, &out, stream));
        }
        const int64_t N = g4r_wait_num_rendered(gctx);
        check(N);
        if (N > cap) {
            cap = N;
            binning = at::empty({(int64_t)g4r_binning_bytes(cap)}, u8);
            Tensor sort_scratch = at::empty({(int64_t)g4r_sort_scratch_bytes(cap)}, u8);
            check(g4r_forward_render(gctx, &frame, &g, geom.data_ptr(), img.data_ptr(), binning.data_ptr(), sort_scratch.data_ptr(), cap, &out, stream));
        }
        t_last_num_rendered = N;
        ctx->save_for_backward({col, m3, sc, rot, cov, radii, shc, geom, binning, img, bgc, view, proj, praw, cam});
        ctx->mark_non_differentiable({radii, n_touched});
        return {color, radii, depth, opacity, n_touched};
    }

    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        // like the reference, the gradients of opacity / radii / n_touched are dropped (DGR/diff_gaussian_rasterization/__init__.py:108-130)
        const auto saved = ctx->get_saved_variables();
        const Tensor &col = saved[0], &m3 = saved[1], &sc = saved[2], &rot = saved[3], &cov = saved[4], &radii = saved[5], &shc = saved[6];
        const Tensor &geom = saved[7], &binning = saved[8], &img = saved[9];
        const c10::Device device = m3.device();
        c10::cuda::CUDAGuard guard(device);
        const int P = (int)ctx->saved_data["P"].toInt();
        const Settings st{ctx->saved_data["H"].toInt(), ctx->saved_data["W"].toInt(), ctx->saved_data["sh_degree"].toInt(), ctx->saved_data["tanfovx"].toDouble(),
                          ctx->saved_data["tanfovy"].toDouble(), ctx->saved_data["scale_modifier"].toDouble(), ctx->saved_data["prefiltered"].toBool()};
        const int M = shc.numel() ? (int)shc.size(1) : 0;
        auto f32 = at::TensorOptions().dtype(at::kFloat).device(device);
        auto want = [&](int i) { return ctx->needs_input_grad(i); };
        Tensor tau = at::empty({8}, f32);
        Tensor g_m3 = want(0) ? at::empty({P, 3}, f32) : Tensor(), g_m2 = want(1) ? at::empty({P, 3}, f32) : Tensor();
        Tensor g_sh = (shc.numel() && want(2)) ? at::empty({P, M, 3}, f32) : Tensor();
        Tensor g_col = (col.numel() && want(3)) ? at::empty({P, 3}, f32) : Tensor();
        Tensor g_op = want(4) ? at::empty(ctx->saved_data["opacities_shape"].toIntVector(), f32) : Tensor();
        Tensor g_sc = (sc.numel() && want(5)) ? at::empty({P, 3}, f32) : Tensor();
        Tensor g_rot = (rot.numel() && want(6)) ? at::empty({P, 4}, f32) : Tensor();
        Tensor g_cov = (cov.numel() && want(7)) ? at::empty({P, 6}, f32) : Tensor();
        if (P == 0) {
            tau.zero_();
        } else {
            Tensor gc = dev_f32(grads[0], device), gd = dev_f32(grads[2], device);
            Tensor scratch = at::empty({(int64_t)g4r_backward_scratch_bytes(P)}, at::TensorOptions().dtype(at::kByte).device(device));
            const G4RFrame frame = make_frame(st, M, saved[10], saved[11], saved[12], saved[13], saved[14]);
            G4RGaussians g{};
            g.P = P;
            g.means3D = fptr(m3); g.opacities = fptr(m3);       // opacities are not read in backward (they live in the splat records)
            g.shs = fptr(shc); g.colors_precomp = fptr(col); g.scales = fptr(sc); g.rotations = fptr(rot); g.cov3D_precomp = fptr(cov);
            G4RBackwardIO io{};
            io.dL_dcolor = gc.data_ptr<float>(); io.dL_ddepth = gd.data_ptr<float>();
            auto out = [](Tensor& t) { return t.defined() && t.numel() ? t.data_ptr<float>() : nullptr; };
            io.dL_dmeans3D = out(g_m3); io.dL_dmeans2D = out(g_m2); io.dL_dopacity = out(g_op); io.dL_dshs = out(g_sh);
            io.dL_dcolors_precomp = out(g_col); io.dL_dscales = out(g_sc); io.dL_drotations = out(g_rot); io.dL_dcov3D = out(g_cov);
            io.dL_dtau = tau.data_ptr<float>();
            void* stream = (void*)at::cuda::getCurrentCUDAStream(device.index()).stream();
            check(g4r_backward(&frame, &g, radii.data_ptr<int32_t>(), geom.data_ptr(), img.data_ptr(), binning.data_ptr(), scratch.data_ptr(), &io, stream));
        }
        Tensor g_rho = want(9) ? tau.slice(0, 0, 3).view({1, -1}) : Tensor();
        Tensor g_theta = want(8) ? tau.slice(0, 3, 6).view({1, -1}) : Tensor();
        // one entry per forward argument: 10 differentiable tensors, then the camera tensors and the scalars
        return {g_m3, g_m2, g_sh, g_col, g_op, g_sc, g_rot, g_cov, g_theta, g_rho, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(),
                Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

// returns (color, radii, depth, opacity, n_touched, num_rendered)
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, int64_t> rasterize(Tensor means3D, Tensor means2D, Tensor sh, Tensor colors_precomp, Tensor opacities,
                                                                      Tensor scales, Tensor rotations, Tensor cov3Ds_precomp, Tensor theta, Tensor rho, Tensor bg,
                                                                      Tensor viewmatrix, Tensor projmatrix, Tensor projmatrix_raw, Tensor campos, int64_t H,
                                                                      int64_t W, double tanfovx, double tanfovy, double scale_modifier, int64_t sh_degree,
                                                                      bool prefiltered, int64_t hint) {
    auto out = RasterizeFn::apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho, bg, viewmatrix, projmatrix,
                                  projmatrix_raw, campos, H, W, tanfovx, tanfovy, scale_modifier, sh_degree, prefiltered, hint);
    return {out[0], out[1], out[2], out[3], out[4], t_last_num_rendered};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "C++ host side of diff_gaussian_rasterization's standard op (calls the C ABI of libg4r.so)";
    // the GIL is released for the call: the forward blocks on a CUDA event once (num_rendered), like the ctypes call of the Python path
    m.def("rasterize", &rasterize, pybind11::call_guard<pybind11::gil_scoped_release>(),
          "forward of the rasterizer op with autograd (returns the 5 outputs and num_rendered)");
    m.def("abi_version", []() { return g4r_version(); });
    m.def("struct_sizes", []() {     // of the g4r.h this file was compiled against; the importer compares them with the library's
        return std::vector<int64_t>{(int64_t)sizeof(G4RFrame), (int64_t)sizeof(G4RGaussians), (int64_t)sizeof(G4RForwardOut), (int64_t)sizeof(G4RBackwardIO)};
    });
}
