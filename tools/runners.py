"""Run this repo's rasterizer / the reference / the oracle on a tools.scenes.Scene and compare results.
Test + bench infrastructure (the product package never imports this)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "4dgs-slam_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def settings_for(sc, dgr):
    return dgr.GaussianRasterizationSettings(
        image_height=sc.H, image_width=sc.W, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, bg=sc.bg,
        scale_modifier=sc.scale_modifier, viewmatrix=sc.viewmatrix, projmatrix=sc.projmatrix,
        projmatrix_raw=sc.projmatrix_raw, sh_degree=sc.sh_degree, campos=sc.campos, prefiltered=False, debug=False)


def run_public_api(sc, dgr, want_grads: bool = True) -> dict:
    """Forward (+backward) through the public GaussianRasterizer API of module `dgr` (ours or the reference)."""
    rs = settings_for(sc, dgr)
    names = ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp")
    leaf = {}
    for k in names:
        v = getattr(sc, k)
        leaf[k] = None if v is None else v.detach().clone().requires_grad_(want_grads)
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=want_grads)
    theta = torch.zeros(3, device=sc.means3D.device, requires_grad=want_grads)
    rho = torch.zeros(3, device=sc.means3D.device, requires_grad=want_grads)
    color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
        means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], shs=leaf["shs"],
        colors_precomp=leaf["colors_precomp"], scales=leaf["scales"], rotations=leaf["rotations"],
        cov3D_precomp=leaf["cov3D_precomp"], theta=theta, rho=rho)
    out = dict(color=color.detach(), radii=radii, depth=depth.detach(), opacity=opacity.detach(), n_touched=n_touched)
    if want_grads:
        loss = (color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()
        loss.backward()
        g = lambda t: None if t is None or t.grad is None else t.grad
        out.update(dL_dmeans3D=g(leaf["means3D"]), dL_dmeans2D=g(means2D), dL_dopacity=g(leaf["opacities"]),
                   dL_dshs=g(leaf["shs"]), dL_dcolors=g(leaf["colors_precomp"]), dL_dscales=g(leaf["scales"]),
                   dL_drots=g(leaf["rotations"]), dL_dcov3D=g(leaf["cov3D_precomp"]),
                   dL_dtau=torch.cat([rho.grad.reshape(-1), theta.grad.reshape(-1)]))
    return out


def run_g4r(sc, want_grads: bool = True) -> dict:
    """This repo's rasterizer: public API for outputs/grads + the inspection hook for the integer state."""
    import diff_gaussian_rasterization as dgr
    out = run_public_api(sc, dgr, want_grads)
    rs = settings_for(sc, dgr)
    _, info = dgr.rasterize_gaussians_with_state(rs, sc.means3D, sc.opacities, shs=sc.shs, colors_precomp=sc.colors_precomp,
                                                 scales=sc.scales, rotations=sc.rotations, cov3D_precomp=sc.cov3D_precomp)
    out.update(info)
    return out


def run_oracle(sc_cpu, precision: str = "f32", want_grads: bool = True) -> dict:
    from oracle.g4r_oracle import Oracle, scene_dict
    ora = Oracle(precision)
    f = ora.forward(scene_dict(sc_cpu))
    out = {k: v for k, v in f.items() if k != "_state"}
    if want_grads:
        g = ora.backward(f, sc_cpu.grad_color, sc_cpu.grad_depth)
        out.update({k: g[k] for k in ("dL_dmeans3D", "dL_dmeans2D", "dL_dopacity", "dL_dshs", "dL_dscales", "dL_drots", "dL_dcov3D",
                                      "dL_dcolors", "dL_dtau")})
        if sc_cpu.colors_precomp is None:
            out["dL_dcolors"] = None
        if sc_cpu.cov3D_precomp is None:
            out["dL_dcov3D"] = None
    return out


def _np(v):
    if v is None:
        return None
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return np.asarray(v)


INT_KEYS = ("radii", "n_touched", "point_list", "ranges", "n_contrib")
IMG_KEYS = ("color", "depth", "opacity", "final_T")
GRAD_KEYS = ("dL_dmeans3D", "dL_dmeans2D", "dL_dopacity", "dL_dshs", "dL_dcolors", "dL_dscales", "dL_drots", "dL_dcov3D", "dL_dtau")


def compare(a: dict, b: dict) -> dict:
    """Metrics of a (candidate) against b (truth).  ints: number of mismatching entries; images: max abs / max rel-to-peak
    error and the count of bitwise-different floats; grads: max abs error relative to the largest |truth| entry and the
    relative L2 error."""
    rep = {}
    if "num_rendered" in a and "num_rendered" in b:
        rep["num_rendered"] = (int(a["num_rendered"]), int(b["num_rendered"]))
    for k in INT_KEYS:
        x, y = _np(a.get(k)), _np(b.get(k))
        if x is None or y is None:
            continue
        x, y = x.astype(np.int64).reshape(-1), y.astype(np.int64).reshape(-1)
        rep[k] = dict(mismatch=int((x != y).sum()) if x.shape == y.shape else -1, size=int(y.size))
    for k in IMG_KEYS:
        x, y = _np(a.get(k)), _np(b.get(k))
        if x is None or y is None:
            continue
        x, y = x.reshape(-1).astype(np.float32), y.reshape(-1).astype(np.float32)
        d = np.abs(x.astype(np.float64) - y.astype(np.float64))
        rep[k] = dict(max_abs=float(d.max()) if d.size else 0.0, rel_to_peak=float(d.max() / (np.abs(y).max() + 1e-30)) if d.size else 0.0,
                      bit_diff=int((x.view(np.uint32) != y.view(np.uint32)).sum()), size=int(y.size))
    for k in GRAD_KEYS:
        x, y = _np(a.get(k)), _np(b.get(k))
        if x is None or y is None or y.size == 0:
            continue
        x, y = x.reshape(-1).astype(np.float64), y.reshape(-1).astype(np.float64)
        if x.shape != y.shape:
            rep[k] = dict(shape_mismatch=(x.shape, y.shape))
            continue
        d = np.abs(x - y)
        rep[k] = dict(max_abs_rel=float(d.max() / (np.abs(y).max() + 1e-30)), l2_rel=float(np.linalg.norm(d) / (np.linalg.norm(y) + 1e-30)),
                      peak=float(np.abs(y).max()))
    return rep


def fmt_report(rep: dict) -> str:
    lines = []
    for k, v in rep.items():
        lines.append(f"    {k:14s} {v}")
    return "\n".join(lines)


# ---------------------------------------------------------------------------------------------------------------------
# raw-parameter mode (FusedGaussianRasterizer) helpers
# ---------------------------------------------------------------------------------------------------------------------
def raw_parameters(sc, scale_dim: int = 3, seed: int = 0) -> dict:
    """GaussianModel-style raw parameters whose activations reproduce the scene: _opacity = logit, _scaling = log (the mean
    log-scale when isotropic), _rotation = the unit quaternion times a random length, SH split into dc / rest."""
    g = torch.Generator().manual_seed(seed)
    P = sc.P
    dev = sc.means3D.device
    o = sc.opacities.double().clamp(1e-6, 1 - 1e-6)
    length = (0.25 + 3.75 * torch.rand(P, 1, generator=g, dtype=torch.float64)).to(dev)
    scaling = torch.log(sc.scales.double())
    if scale_dim == 1:
        scaling = scaling.mean(1, keepdim=True)
    return dict(xyz=sc.means3D.clone(), opacity=torch.log(o / (1 - o)).float(), scaling=scaling.float(),
                rotation=(sc.rotations.double() * length).float(), dc=sc.shs[:, :1].contiguous().clone(),
                rest=sc.shs[:, 1:].contiguous().clone())


def run_raw(sc, dgr, raw: dict, fused: bool) -> dict:
    """fwd+bwd on the raw parameters: `fused` -> FusedGaussianRasterizer; else the reference's torch prelude
    (gaussian_model.py:100-128) followed by the standard GaussianRasterizer."""
    rs = settings_for(sc, dgr)
    dev = sc.means3D.device
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in raw.items()}
    m2d = torch.zeros_like(leaf["xyz"], requires_grad=True)
    theta = torch.zeros(3, device=dev, requires_grad=True)
    rho = torch.zeros(3, device=dev, requires_grad=True)
    if fused:
        out = dgr.FusedGaussianRasterizer(rs)(xyz=leaf["xyz"], means2D=m2d, features_dc=leaf["dc"], features_rest=leaf["rest"],
                                              opacity_raw=leaf["opacity"], scaling_raw=leaf["scaling"], rotation_raw=leaf["rotation"],
                                              theta=theta, rho=rho)
    else:
        scal = torch.exp(leaf["scaling"])
        if scal.shape[-1] == 1:
            scal = scal.repeat(1, 3)
        out = dgr.GaussianRasterizer(rs)(means3D=leaf["xyz"], means2D=m2d, opacities=torch.sigmoid(leaf["opacity"]),
                                         shs=torch.cat((leaf["dc"], leaf["rest"]), dim=1), scales=scal,
                                         rotations=torch.nn.functional.normalize(leaf["rotation"]), theta=theta, rho=rho)
    color, radii, depth, opacity, n_touched = out
    ((color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()).backward()
    res = dict(color=color.detach(), depth=depth.detach(), opacity=opacity.detach(), radii=radii, n_touched=n_touched,
               dL_dmeans2D=m2d.grad, dL_dtau=torch.cat([rho.grad.reshape(-1), theta.grad.reshape(-1)]))
    for k in leaf:
        res["g_" + k] = leaf[k].grad
    return res
