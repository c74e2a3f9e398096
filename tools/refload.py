"""Load the UNMODIFIED reference rasterizer from baseline/_ref under the module name ``dgr_ref`` (so that it can live
next to this repo's drop-in package of the same import name) and decode its opaque scratch buffers.

Test / bench infrastructure only."""
from __future__ import annotations

import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, "baseline", "_ref", "diff_gaussian_rasterization")


def available() -> bool:
    return os.path.exists(os.path.join(REF_PKG, "__init__.py"))


def load():
    if "dgr_ref" in sys.modules:
        return sys.modules["dgr_ref"]
    if not available():
        raise ImportError("baseline/_ref is absent: run `python -c 'import __graft_entry__ as g; g.build_reference()'` "
                          "in a container that has /root/reference")
    spec = importlib.util.spec_from_file_location("dgr_ref", os.path.join(REF_PKG, "__init__.py"),
                                                  submodule_search_locations=[REF_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["dgr_ref"] = mod
    spec.loader.exec_module(mod)
    return mod


def _al(o: int, a: int = 128) -> int:
    return (o + a - 1) // a * a


def decode_buffers(P: int, W: int, H: int, N: int, geom: torch.Tensor, binning: torch.Tensor, img: torch.Tensor) -> dict:
    """Views into the reference's geomBuffer / binningBuffer / imgBuffer.  Layout = the obtain() sequence of
    GeometryState / BinningState / ImageState::fromChunk (DGR/cuda_rasterizer/rasterizer_impl.cu:155-194) with 128-byte
    alignment (rasterizer_impl.h:22-27); torch allocations are >= 512-byte aligned so offsets are relative."""
    out = {}
    o = _al(geom.data_ptr()) - geom.data_ptr()
    def take(buf, o, nbytes, dtype, shape):
        o = _al(buf.data_ptr() + o) - buf.data_ptr()
        return buf[o:o + nbytes].view(dtype).view(*shape), o + nbytes
    out["depths"], o = take(geom, 0, 4 * P, torch.float32, (P,))
    out["clamped"], o = take(geom, o, 3 * P, torch.uint8, (P, 3))
    out["internal_radii"], o = take(geom, o, 4 * P, torch.int32, (P,))
    out["means2D"], o = take(geom, o, 8 * P, torch.float32, (P, 2))
    out["cov3D"], o = take(geom, o, 24 * P, torch.float32, (P, 6))
    out["conic_opacity"], o = take(geom, o, 16 * P, torch.float32, (P, 4))
    out["rgb"], o = take(geom, o, 12 * P, torch.float32, (P, 3))
    out["tiles_touched"], o = take(geom, o, 4 * P, torch.int32, (P,))
    X = W * H
    out["final_T"], o = take(img, 0, 4 * X, torch.float32, (H, W))
    out["n_contrib"], o = take(img, o, 4 * X, torch.int32, (H, W))
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    rng, o = take(img, o, 8 * X, torch.int32, (X, 2))
    out["ranges"] = rng[:tiles]
    if N > 0:
        out["point_list"], _ = take(binning, 0, 4 * N, torch.int32, (N,))
    else:
        out["point_list"] = torch.zeros(0, dtype=torch.int32, device=geom.device)
    return out


def run_reference(sc, want_grads: bool = True) -> dict:
    """Run the reference fwd (+bwd) on a tools.scenes.Scene that already lives on the GPU.  Returns outputs, saved
    integer state and gradients as GPU tensors."""
    ref = load()
    e = torch.Tensor([])
    dev = sc.means3D.device
    shs = sc.shs if sc.shs is not None else e
    colors = sc.colors_precomp if sc.colors_precomp is not None else e
    scales = sc.scales if sc.scales is not None else e
    rots = sc.rotations if sc.rotations is not None else e
    cov = sc.cov3D_precomp if sc.cov3D_precomp is not None else e
    args = (sc.bg, sc.means3D, colors, sc.opacities, scales, rots, sc.scale_modifier, cov, sc.viewmatrix, sc.projmatrix,
            sc.projmatrix_raw, sc.tanfovx, sc.tanfovy, sc.H, sc.W, shs, sc.sh_degree, sc.campos, False, False)
    N, color, radii, geom, binning, img, depth, opacity, n_touched = ref._C.rasterize_gaussians(*args)
    out = dict(num_rendered=int(N), color=color, radii=radii, depth=depth, opacity=opacity, n_touched=n_touched)
    out.update(decode_buffers(sc.P, sc.W, sc.H, int(N), geom, binning, img))
    if want_grads:
        bargs = (sc.bg, sc.means3D, radii, colors, scales, rots, sc.scale_modifier, cov, sc.viewmatrix, sc.projmatrix,
                 sc.projmatrix_raw, sc.tanfovx, sc.tanfovy, sc.grad_color, sc.grad_depth, shs, sc.sh_degree, sc.campos, geom,
                 int(N), binning, img, False)
        (g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot, g_tau) = ref._C.rasterize_gaussians_backward(*bargs)
        tau = g_tau.view(-1, 6).sum(0)
        out.update(dL_dmeans2D=g_means2D, dL_dcolors=g_colors, dL_dopacity=g_opac, dL_dmeans3D=g_means3D, dL_dcov3D=g_cov3D,
                   dL_dshs=g_sh, dL_dscales=g_scales, dL_drots=g_rot, dL_dtau=tau, dL_dtau_rows=g_tau)
    return out
