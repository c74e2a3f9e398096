"""The reference's OWN caller runs unchanged on the drop-in.

`gaussian_splatting/gaussian_renderer/__init__.py` of the reference -- byte-for-byte the upstream file, installed into the
git-ignored baseline/_ref/caller by `__graft_entry__.build()` -- is executed twice: once with `diff_gaussian_rasterization`
resolved to this repo's package and once resolved to the unmodified reference build (baseline/_ref).  `render()` (plain, with
the static `mask=` of the tracking loop, with the dynamic offsets dx/ds/dr of the mapping loop) and `render_flow()` are called
on duck-typed `pc` / `viewpoint_camera` objects; the returned dicts (`gaussian_renderer/__init__.py:218-226,352-361`) must
agree bit for bit for images and integer outputs and within 1e-3 for every gradient that reaches the model parameters, the
pose deltas and `viewspace_points`."""
import pytest
import torch

from tools import refcaller, refload
from tools.scenes import make_scene
from tools.slam_shapes import PIPE, DuckCamera, DuckGaussians
from tools import slam_shapes

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (refcaller.available() and refload.available()), reason="baseline/_ref (reference build + caller) not present")]


def _l2rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _ducks(device, sh_degree, M):
    sc_cpu = make_scene(20000, 320, 240, sh_degree=sh_degree, sh_coeffs=M, seed=300 + M)
    sc = sc_cpu.to(device)
    g = torch.Generator().manual_seed(7)
    dygs = (torch.rand(sc.P, generator=g) < 0.3).to(device)
    gt = torch.rand(3, sc.H, sc.W, generator=g).to(device)
    gd = (0.5 + 5.0 * torch.rand(1, sc.H, sc.W, generator=g)).to(device)
    return sc, dygs, gt, gd


def _run(renderer, sc, dygs, gt, gd, mode):
    """One render through `renderer` (a loaded gaussian_renderer module, or slam_shapes with a bound dgr) + a loss that uses
    every differentiable output + backward.  Returns the dict and the gradients."""
    pc = DuckGaussians.from_scene(sc, dygs=dygs)
    cam = DuckCamera.from_scene(sc, image=gt, depth=gd)
    bg = torch.tensor([0.0, 0.0, 0.0], device=sc.means3D.device)            # slam.py:97-98 uses black or white
    kw = {}
    extra = []
    nd = int(dygs.sum())
    if mode == "mask":
        kw["mask"] = (pc.dygs == False)          # noqa: E712  (utils/slam_frontend.py:413)
    if mode == "offsets":
        g = torch.Generator().manual_seed(11)
        dx = (0.01 * torch.randn(nd, 3, generator=g)).to(dygs.device).requires_grad_(True)
        ds = (0.001 * torch.rand(nd, 3, generator=g)).to(dygs.device).requires_grad_(True)
        dr = (0.01 * torch.randn(nd, 4, generator=g)).to(dygs.device).requires_grad_(True)
        kw.update(dx=dx, ds=ds, dr=dr)
        extra = [dx, ds, dr]
    pkg = renderer.render(cam, pc, PIPE, bg, **kw)
    loss = slam_shapes.loss_tracking_rgbd(pkg["render"], pkg["depth"], pkg["opacity"], cam) + 0.1 * (pkg["depth"] * gd).mean()
    loss.backward()
    grads = {n: p.grad for n, p in zip(("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"), pc.parameters())}
    grads.update(theta=cam.cam_rot_delta.grad, rho=cam.cam_trans_delta.grad, exposure_a=cam.exposure_a.grad,
                 viewspace_points=pkg["viewspace_points"].grad)
    for n, t in zip(("dx", "ds", "dr"), extra):
        grads[n] = t.grad
    return pkg, grads


@pytest.mark.parametrize("mode", ["plain", "mask", "offsets"])
@pytest.mark.parametrize("deg,M", [(0, 1), (1, 4)])
def test_reference_render_runs_unchanged_on_the_drop_in(device, mode, deg, M):
    import diff_gaussian_rasterization as ours
    ref = refload.load()
    sc, dygs, gt, gd = _ducks(device, deg, M)
    r_ours = refcaller.load_renderer(ours, "ours")
    r_ref = refcaller.load_renderer(ref, "ref")
    a, ga = _run(r_ours, sc, dygs, gt, gd, mode)
    b, gb = _run(r_ref, sc, dygs, gt, gd, mode)
    assert set(a.keys()) == set(b.keys()) == {"render", "viewspace_points", "visibility_filter", "radii", "depth", "opacity", "n_touched"}
    for k in ("radii", "n_touched", "visibility_filter"):
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    for k in ("render", "depth", "opacity"):
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), (k, float((a[k] - b[k]).abs().max()))     # identical bits
    assert int(a["radii"].numel()) == (int((~dygs).sum()) if mode == "mask" else sc.P)
    for k in gb:
        if gb[k] is None:
            assert ga[k] is None or float(ga[k].abs().max()) == 0.0, k
            continue
        assert ga[k] is not None and ga[k].shape == gb[k].shape, k
        assert _l2rel(ga[k], gb[k]) < 1e-3, (k, _l2rel(ga[k], gb[k]))
    # the restated caller used by bench.py (tools/slam_shapes.render) is the same computation as the reference's file
    class Bound:
        @staticmethod
        def render(*args, **kw):
            return slam_shapes.render(ours, *args, **kw)
    c, gc = _run(Bound, sc, dygs, gt, gd, mode)
    for k in ("render", "depth", "opacity", "radii", "n_touched"):
        assert torch.equal(a[k], c[k]), k


def test_reference_render_flow_runs_unchanged_on_the_drop_in(device):
    """render_flow (gaussian_renderer/__init__.py:229-361): colors_precomp = per-Gaussian NDC flow + dynamic flag, no theta/rho,
    detached opacity/scales, gradients reach d_xyz1 / d_xyz2 / d_rotation1 / d_scaling1."""
    import diff_gaussian_rasterization as ours
    ref = refload.load()
    sc, dygs, gt, gd = _ducks(device, 0, 1)
    nd = int(dygs.sum())
    out = {}
    for tag, dgr in (("ours", ours), ("ref", ref)):
        renderer = refcaller.load_renderer(dgr, "flow_" + tag)
        pc = DuckGaussians.from_scene(sc, dygs=dygs)
        cam1 = DuckCamera.from_scene(sc)
        cam2 = slam_shapes.perturbed(cam1, 1)
        g = torch.Generator().manual_seed(13)
        d1 = (0.01 * torch.randn(nd, 3, generator=g)).to(device).requires_grad_(True)
        d2 = (0.01 * torch.randn(nd, 3, generator=g)).to(device).requires_grad_(True)
        dr = (0.01 * torch.randn(nd, 4, generator=g)).to(device).requires_grad_(True)
        ds = (0.001 * torch.rand(nd, 3, generator=g)).to(device).requires_grad_(True)
        pkg = renderer.render_flow(pc, cam1, cam2, d1, d2, dr, ds)
        ((pkg["render"] * gt).sum() + (pkg["depth"] * gd).sum() * 1e-3).backward()
        out[tag] = (pkg, dict(d1=d1.grad, d2=d2.grad, dr=dr.grad, ds=ds.grad, xyz=pc._xyz.grad, vsp=pkg["viewspace_points"].grad))
    a, ga = out["ours"]
    b, gb = out["ref"]
    assert set(a.keys()) == set(b.keys()) == {"render", "depth", "alpha", "viewspace_points", "visibility_filter", "radii"}
    for k in ("render", "depth", "alpha", "radii", "visibility_filter"):
        assert torch.equal(a[k], b[k]), k
    for k in gb:
        if gb[k] is None:
            assert ga[k] is None or float(ga[k].abs().max()) == 0.0, k
        else:
            assert _l2rel(ga[k], gb[k]) < 1e-3, (k, _l2rel(ga[k], gb[k]))


def test_tracking_and_mapping_loop_bodies_match_the_reference_build(device):
    """The two hot loops as bench.py runs them (tools/slam_shapes.tracking_iteration / mapping_iteration, restating
    utils/slam_frontend.py:411-448 and utils/slam_backend.py:357-771): three iterations on this repo's rasterizer and on the
    reference build end in the same poses and parameters (1e-3: Adam's g / sqrt(v) amplifies the atomics-order noise of small
    gradient entries to a full learning-rate step)."""
    import diff_gaussian_rasterization as ours
    ref = refload.load()
    sc, dygs, gt, gd = _ducks(device, 0, 1)
    bg = torch.zeros(3, device=device)
    res = {}
    for tag, dgr in (("ours", ours), ("ref", ref)):
        render_fn = lambda *a, _d=dgr, **k: slam_shapes.render(_d, *a, **k)
        pc = DuckGaussians.from_scene(sc, dygs=dygs)
        cam = slam_shapes.perturbed(DuckCamera.from_scene(sc, image=gt, depth=gd), 3, rot=0.004, trans=0.01)
        opt = slam_shapes.pose_optimizer(cam)
        for _ in range(3):
            slam_shapes.tracking_iteration(dgr, render_fn, cam, pc, bg, opt)
        views = [slam_shapes.perturbed(DuckCamera.from_scene(sc, image=gt, depth=gd), 10 + i) for i in range(3)]
        gopt = pc.optimizer()
        popts = [slam_shapes.pose_optimizer(v) for v in views]
        for _ in range(2):
            slam_shapes.mapping_iteration(dgr, render_fn, views, pc, bg, gopt, popts)
        res[tag] = (cam.R.clone(), cam.T.clone(), [p.detach().clone() for p in pc.parameters()], [v.T.clone() for v in views])
    a, b = res["ours"], res["ref"]
    assert _l2rel(a[0], b[0]) < 1e-3 and _l2rel(a[1], b[1]) < 1e-3
    for x, y in zip(a[2], b[2]):
        assert _l2rel(x, y) < 1e-3
    for x, y in zip(a[3], b[3]):
        assert _l2rel(x, y) < 1e-3
