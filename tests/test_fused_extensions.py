"""Opt-in extensions beyond the reference surface: the static mask + dynamic offsets of render()
inside the kernels, and the fused RGB-D tracking / mapping losses.

CPU: the numpy restatement of the losses (oracle/g4r_oracle.py slam_loss_ref) against vectors produced by the reference's own
get_loss_tracking / get_loss_mapping under torch autograd (tests/golden/slam_loss.npz, tools/make_loss_golden.py).
GPU: the kernels against those vectors, against the restatement on full-size random images, and against the reference's torch
prelude feeding the standard rasterizer."""
import os

import numpy as np
import pytest
import torch

from oracle.g4r_oracle import prelude_ref, slam_loss_ref
from tools import runners
from tools.scenes import make_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "slam_loss.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files})


def _case(name):
    g = lambda k: GOLD[f"{name}/{k}"] if f"{name}/{k}" in GOLD.files else None
    mode = "tracking" if name.startswith("track") else "mapping"
    uid = int(g("uid"))
    motion = g("motion_mask")
    if mode == "tracking" and uid == 0:
        motion = None                                   # slam_utils.py:76: the motion mask only applies for uid > 0
    return mode, dict(image=g("image"), depth=g("depth"), opacity=g("opacity"), gt_image=g("gt_image"), gt_depth=g("gt_depth"),
                      exposure=g("exposure"), motion_mask=motion, grad_mask=g("grad_mask") if mode == "tracking" else None), g


@pytest.mark.parametrize("name", CASES)
def test_loss_restatement_matches_the_reference_functions(name):
    mode, kw, g = _case(name)
    r = slam_loss_ref(mode, kw["image"], kw["depth"], kw["gt_image"], kw["gt_depth"], opacity=kw["opacity"], exposure=kw["exposure"],
                      motion_mask=kw["motion_mask"], grad_mask=kw["grad_mask"])
    assert abs(r["loss"] - float(g("loss"))) < 1e-6 * abs(float(g("loss")))
    assert np.abs(r["d_image"].reshape(-1) - g("d_image").reshape(-1)).max() < 1e-9
    assert np.abs(r["d_depth"].reshape(-1) - g("d_depth").reshape(-1)).max() < 1e-9
    assert np.abs(r["d_exposure"] - g("d_exposure")).max() < 1e-6 * np.abs(g("d_exposure")).max() + 1e-9


def test_prelude_restatement_matches_torch_indexing():
    """prelude_ref == the reference's torch statements (gaussian_renderer/__init__.py:159-191) on CPU tensors."""
    g = torch.Generator().manual_seed(3)
    P = 200
    xyz, sc, rot = torch.randn(P, 3, generator=g), torch.rand(P, 3, generator=g), torch.randn(P, 4, generator=g)
    dygs = torch.rand(P, generator=g) < 0.3
    nd = int(dygs.sum())
    dx, ds, dr = torch.randn(nd, 3, generator=g), torch.rand(nd, 3, generator=g), torch.randn(nd, 4, generator=g)
    dxyz = torch.zeros_like(xyz); dxyz[dygs] = dx
    dscale = torch.zeros_like(sc); dscale[dygs] = ds
    drot = torch.zeros_like(rot); drot[dygs] = dr
    mask = dygs == False          # noqa: E712
    r = prelude_ref(xyz.numpy(), sc.numpy(), rot.numpy(), dygs.numpy(), dx.numpy(), ds.numpy(), dr.numpy(), mask.numpy())
    assert np.array_equal(r["means3D"], (xyz + dxyz)[mask].numpy())
    assert np.array_equal(r["scales"], (sc + dscale)[mask].numpy()) and np.array_equal(r["rotations"], (rot + drot)[mask].numpy())


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_fused_loss_kernel_matches_the_reference_vectors(device, name):
    from diff_gaussian_rasterization.losses import slam_loss
    mode, kw, g = _case(name)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device)
    image, depth = t(kw["image"]).requires_grad_(True), t(kw["depth"]).requires_grad_(True)
    ea = torch.tensor([float(kw["exposure"][0])], device=device, requires_grad=True)
    eb = torch.tensor([float(kw["exposure"][1])], device=device, requires_grad=True)
    loss = slam_loss(mode, image, depth, t(kw["gt_image"]), t(kw["gt_depth"]), opacity=t(kw["opacity"]), exposure_a=ea, exposure_b=eb,
                     motion_mask=t(kw["motion_mask"]), grad_mask=t(kw["grad_mask"]))
    (2.0 * loss).backward()                              # a non-trivial upstream gradient
    assert abs(float(loss) - float(g("loss"))) < 2e-6 * abs(float(g("loss")))
    assert float((image.grad.cpu() - 2 * torch.from_numpy(g("d_image"))).abs().max()) < 1e-9
    assert float((depth.grad.cpu() - 2 * torch.from_numpy(g("d_depth")).view(depth.shape)).abs().max()) < 1e-9
    d_exp = np.array([float(ea.grad), float(eb.grad)])
    assert np.abs(d_exp - 2 * g("d_exposure")).max() < 1e-5 * np.abs(2 * g("d_exposure")).max() + 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["tracking", "mapping"])
def test_fused_loss_kernel_full_size_against_restatement(device, mode):
    from diff_gaussian_rasterization.losses import slam_loss
    g = torch.Generator().manual_seed(9)
    H, W = 480, 640
    image, depth = torch.rand(3, H, W, generator=g), 0.2 + 4 * torch.rand(1, H, W, generator=g)
    opacity = 0.8 + 0.2 * torch.rand(1, H, W, generator=g)
    gt, gd = torch.rand(3, H, W, generator=g), 4 * torch.rand(1, H, W, generator=g)
    mm = torch.rand(H, W, generator=g) > 0.2
    gm = torch.rand(1, H, W, generator=g) > 0.1 if mode == "tracking" else None
    ref = slam_loss_ref(mode, image.numpy(), depth.numpy(), gt.numpy(), gd.numpy(), opacity=opacity.numpy(), exposure=(0.1, 0.03),
                        motion_mask=mm.numpy(), grad_mask=None if gm is None else gm.numpy())
    d = lambda x: None if x is None else x.to(device)
    im, dp = d(image).requires_grad_(True), d(depth).requires_grad_(True)
    ea, eb = torch.tensor([0.1], device=device, requires_grad=True), torch.tensor([0.03], device=device, requires_grad=True)
    loss = slam_loss(mode, im, dp, d(gt), d(gd), opacity=d(opacity), exposure_a=ea, exposure_b=eb, motion_mask=d(mm), grad_mask=d(gm))
    loss.backward()
    assert abs(float(loss) - ref["loss"]) < 1e-5 * abs(ref["loss"])
    assert float((im.grad.cpu().double() - torch.from_numpy(ref["d_image"]).view(3, H, W)).abs().max()) < 1e-10
    assert float((dp.grad.cpu().double().view(-1) - torch.from_numpy(ref["d_depth"])).abs().max()) < 1e-10
    assert abs(float(ea.grad) - ref["d_exposure"][0]) < 1e-4 * abs(ref["d_exposure"][0]) + 1e-8
    assert abs(float(eb.grad) - ref["d_exposure"][1]) < 1e-4 * abs(ref["d_exposure"][1]) + 1e-8


def _l2rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("use_mask,use_offsets", [(True, False), (False, True), (True, True)])
def test_mask_and_dynamic_offsets_in_kernel_match_the_torch_prelude(device, use_mask, use_offsets):
    """FusedGaussianRasterizer(mask=, dx=, ds=, dr=, dyn_slot=) against the reference's prelude: torch activations, offsets
    scattered into zero tensors and added, boolean-mask gather, standard rasterizer (gaussian_renderer/__init__.py:108-191).
    Images 1e-4 of peak, integer outputs equal on the kept rows, every gradient 1e-3 (raw parameters, offsets, pose)."""
    import diff_gaussian_rasterization as dgr
    sc = make_scene(20000, 320, 240, sh_degree=1, seed=411).to(device)
    raw = runners.raw_parameters(sc, scale_dim=3, seed=8)
    g = torch.Generator().manual_seed(5)
    dygs = (torch.rand(sc.P, generator=g) < 0.3).to(device)
    nd = int(dygs.sum())
    off = dict(dx=(0.01 * torch.randn(nd, 3, generator=g)).to(device), ds=(0.002 * torch.rand(nd, 3, generator=g)).to(device),
               dr=(0.02 * torch.randn(nd, 4, generator=g)).to(device))
    mask = (dygs == False) if use_mask else None          # noqa: E712
    rs = runners.settings_for(sc, dgr)
    res = {}
    for fused in (True, False):
        leaf = {k: v.detach().clone().requires_grad_(True) for k, v in raw.items()}
        o = {k: v.detach().clone().requires_grad_(True) for k, v in off.items()} if use_offsets else {}
        m2d = torch.zeros_like(leaf["xyz"], requires_grad=True)
        theta, rho = torch.zeros(3, device=device, requires_grad=True), torch.zeros(3, device=device, requires_grad=True)
        if fused:
            out = dgr.FusedGaussianRasterizer(rs)(xyz=leaf["xyz"], means2D=m2d, features_dc=leaf["dc"], features_rest=leaf["rest"],
                                                  opacity_raw=leaf["opacity"], scaling_raw=leaf["scaling"], rotation_raw=leaf["rotation"],
                                                  theta=theta, rho=rho, mask=mask, dyn_slot=dgr.dynamic_slots(dygs) if use_offsets else None, **o)
        else:
            means3D, scales = leaf["xyz"], torch.exp(leaf["scaling"])
            rots = torch.nn.functional.normalize(leaf["rotation"])
            if use_offsets:
                dxyz = torch.zeros_like(means3D); dxyz[dygs] = o["dx"]
                means3D = leaf["xyz"] + dxyz
                dscale = torch.zeros_like(scales); dscale[dygs] = o["ds"]
                scales = scales + dscale
                drot = torch.zeros_like(rots); drot[dygs] = o["dr"]
                rots = rots + drot
            shs, op, m2 = torch.cat((leaf["dc"], leaf["rest"]), dim=1), torch.sigmoid(leaf["opacity"]), m2d
            if mask is not None:
                means3D, scales, rots, shs, op, m2 = means3D[mask], scales[mask], rots[mask], shs[mask], op[mask], m2d[mask]
            out = dgr.GaussianRasterizer(rs)(means3D=means3D, means2D=m2, opacities=op, shs=shs, scales=scales, rotations=rots, theta=theta, rho=rho)
        color, radii, depth, opacity, n_touched = out
        ((color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()).backward()
        res[fused] = dict(color=color.detach(), depth=depth.detach(), radii=radii, n_touched=n_touched, m2d=m2d.grad,
                          tau=torch.cat([rho.grad, theta.grad]), **{"g_" + k: v.grad for k, v in leaf.items()}, **{"g_" + k: v.grad for k, v in o.items()})
    a, b = res[True], res[False]
    for k in ("color", "depth"):
        assert float(((a[k] - b[k]).abs() > 1e-4 * max(1.0, float(b[k].abs().max()))).float().mean()) < 1e-4, k
    keep = mask if mask is not None else torch.ones(sc.P, dtype=torch.bool, device=device)
    assert a["radii"].shape[0] == sc.P                                   # fused outputs keep the full length
    assert int((a["radii"][keep] != b["radii"]).sum()) <= 2
    assert int(a["radii"][~keep].abs().sum()) == 0 and int(a["n_touched"][~keep].abs().sum()) == 0
    assert int((a["n_touched"][keep] != b["n_touched"]).sum()) <= max(4, sc.P // 2000)
    for k in a:
        if not (k.startswith("g_") or k in ("m2d", "tau")):
            continue
        if b[k] is None:
            assert a[k] is None or float(a[k].abs().max()) == 0.0, k
            continue
        assert a[k] is not None and a[k].shape == b[k].shape, k
        assert _l2rel(a[k], b[k]) < 1e-3, (k, _l2rel(a[k], b[k]))
