"""Edge cases of the drop-in surface (SURVEY.md section 4 list + ADVICE.md round 1), on the B200 box."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tools import runners
from tools.scenes import Scene, make_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _leafs(sc, device, grad=True):
    leaf = {k: getattr(sc, k).clone().requires_grad_(grad) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
    theta = torch.zeros(3, device=device, requires_grad=True)
    rho = torch.zeros(3, device=device, requires_grad=True)
    return leaf, m2d, theta, rho


def test_graph_replay_that_overflows_the_capacity_is_detected_and_harmless(device):
    """a replay whose instance count outgrows the capacity fixed at capture must not walk an
    unwritten point_list in the backward.  The kernels compare N with the capacity on the device: the forward outputs stay
    as they were, every gradient comes back as zero, captured_overflow() reports it, and a later in-capacity replay is right."""
    import diff_gaussian_rasterization as dgr
    sc = make_scene(20000, 320, 240, sh_degree=0, seed=95).to(device)
    rs = runners.settings_for(sc, dgr)
    leaf, _, _, _ = _leafs(sc, device)
    small = leaf["scales"].detach().clone()

    def step():
        m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta = torch.zeros(3, device=device, requires_grad=True)
        rho = torch.zeros(3, device=device, requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
            means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf["shs"], scales=leaf["scales"],
            rotations=leaf["rotations"], theta=theta, rho=rho)
        loss = (color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()
        grads = torch.autograd.grad(loss, [leaf["means3D"], leaf["scales"], leaf["rotations"], leaf["opacities"], leaf["shs"], m2d, theta, rho])
        return color, depth, grads

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            eager = step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    dgr.reset_captured()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    graph.replay()
    torch.cuda.synchronize()
    assert not dgr.captured_overflow()
    good_color = out[0].clone()
    assert torch.equal(good_color, eager[0])
    # 12x larger splats: ~100x more (tile, Gaussian) instances than the captured capacity (2 x max(hint, 4 P) + 4096)
    with torch.no_grad():
        leaf["scales"].copy_(small * 12.0)
    graph.replay()
    torch.cuda.synchronize()
    assert dgr.captured_overflow()                     # reported ...
    assert not dgr.captured_overflow()                 # ... and cleared by the read
    assert torch.equal(out[0], good_color)             # forward outputs untouched (stale), nothing was written out of bounds
    for g in out[2]:
        assert float(g.abs().max()) == 0.0             # the backward saw N > capacity: zero gradients, not garbage
    with torch.no_grad():
        leaf["scales"].copy_(small)
    graph.replay()
    torch.cuda.synchronize()
    assert not dgr.captured_overflow()
    assert torch.equal(out[0], eager[0])
    for a, b in zip(out[2], eager[2]):
        assert float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)) < 1e-4


def test_pose_only_backward_skips_the_gaussian_gradients(device):
    """Tracking consumes only grad_theta / grad_rho (utils/slam_frontend.py:441-448).  When the Gaussian inputs do not require
    grad, autograd's needs_input_grad reaches the C ABI as NULL output pointers: nothing per-Gaussian is allocated or written,
    and the pose gradient equals the one of the full backward."""
    import diff_gaussian_rasterization as dgr
    sc = make_scene(30000, 320, 240, sh_degree=0, seed=96).to(device)
    rs = runners.settings_for(sc, dgr)
    res = {}
    for grad in (True, False):
        leaf, m2d, theta, rho = _leafs(sc, device, grad=grad)
        m2d = m2d if grad else torch.zeros_like(leaf["means3D"])
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
            means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf["shs"], scales=leaf["scales"],
            rotations=leaf["rotations"], theta=theta, rho=rho)
        ((color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()).backward()
        res[grad] = (theta.grad.clone(), rho.grad.clone(), color.detach().clone())
        if not grad:
            assert all(v.grad is None for v in leaf.values()) and m2d.grad is None
    assert torch.equal(res[True][2], res[False][2])
    for a, b in zip(res[False][:2], res[True][:2]):
        assert float((a - b).norm() / b.norm()) < 1e-4


def test_saved_state_holds_only_what_backward_reads(device):
    """Per live autograd context: 49 B / Gaussian (records + clamp flags), 4 B / instance (sorted ids), 8 B / pixel; the
    16 B / instance of unsorted pairs die with the forward (BackEnd.map keeps ~10 contexts alive, utils/slam_backend.py:657)."""
    import diff_gaussian_rasterization as dgr
    sc = make_scene(20000, 320, 240, sh_degree=0, seed=97).to(device)
    e = torch.Tensor([])
    out = dgr._forward_impl(sc.means3D, sc.shs, e, sc.opacities, sc.scales, sc.rotations, e, runners.settings_for(sc, dgr))
    st = out[-1]
    cap = st["capacity"]
    assert st["binning"].numel() <= 4 * cap + 1024
    assert st["sort_scratch"].numel() >= 16 * cap
    assert st["binning"].untyped_storage().nbytes() <= 4 * cap + 1024            # not a view into a larger allocation


def test_low_opacity_nan_and_inf_inputs_match_the_oracle(device):
    """opacity < 1/255 (the composite kernels' cull threshold is -1 for those), a NaN mean, an infinite mean and a zero scale:
    the reference neither crashes nor culls them specially (the NaN depth test `z <= 0.2` is false, the tile rectangle of a
    NaN centre is empty); integer outputs must equal the oracle's and the image must stay finite."""
    sc_cpu = make_scene(4000, 160, 120, sh_degree=1, seed=98)
    sc_cpu.opacities[:400] = torch.linspace(1e-6, 1.0 / 200.0, 400).view(-1, 1)
    sc_cpu.means3D[500] = float("nan")
    sc_cpu.means3D[501, 0] = float("inf")
    sc_cpu.scales[502] = 0.0
    mine = runners.run_g4r(sc_cpu.to(device))
    ora = runners.run_oracle(sc_cpu)
    rep = runners.compare(mine, ora)
    for k in ("radii", "point_list", "ranges"):
        assert rep[k]["mismatch"] == 0, (k, rep[k])
    assert int(mine["radii"][500]) == 0 and int(mine["radii"][501]) == 0
    assert bool(torch.isfinite(mine["color"]).all()) and bool(torch.isfinite(mine["depth"]).all())
    a, b = mine["color"].cpu().numpy().reshape(-1), ora["color"].reshape(-1)
    assert (np.abs(a - b) > 1e-4).mean() < 2e-3
    for k in ("dL_dmeans3D", "dL_dopacity", "dL_dscales"):
        g = mine[k]
        keep = torch.ones(sc_cpu.P, dtype=torch.bool, device=device)
        keep[500:503] = False
        assert bool(torch.isfinite(g[keep]).all()), k
        assert rep[k]["l2_rel"] < 2e-3 or not np.isfinite(rep[k]["l2_rel"]), (k, rep[k])


def test_prefiltered_flag_is_accepted(device):
    """prefiltered=True makes the reference __trap() when a culled point shows up (auxiliary.h:156-160); no caller sets it
    (gaussian_renderer/__init__.py:99).  Here the flag is accepted and culled points are simply skipped."""
    import diff_gaussian_rasterization as dgr
    sc = make_scene(3000, 128, 96, sh_degree=0, seed=99).to(device)
    rs = runners.settings_for(sc, dgr)
    a = dgr.GaussianRasterizer(rs)(means3D=sc.means3D, means2D=torch.zeros_like(sc.means3D), opacities=sc.opacities, shs=sc.shs,
                                   scales=sc.scales, rotations=sc.rotations)
    b = dgr.GaussianRasterizer(rs._replace(prefiltered=True))(means3D=sc.means3D, means2D=torch.zeros_like(sc.means3D),
                                                              opacities=sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_forward_and_backward_on_a_side_stream(device):
    """Everything is enqueued on torch's CURRENT stream (the reference uses the legacy default stream); autograd replays the
    backward on the forward's stream."""
    import diff_gaussian_rasterization as dgr
    sc = make_scene(8000, 160, 120, sh_degree=1, seed=101).to(device)
    base = runners.run_public_api(sc, dgr)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        out = runners.run_public_api(sc, dgr)
    side.synchronize()
    for k in ("color", "depth", "opacity", "radii", "n_touched"):
        assert torch.equal(base[k], out[k]), k
    for k in ("dL_dmeans3D", "dL_dscales", "dL_drots", "dL_dopacity", "dL_dshs", "dL_dtau"):
        assert float((out[k].double() - base[k].double()).norm() / base[k].double().norm()) < 1e-4, k


_WORKER = r"""
import hashlib, json, sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r})
import diff_gaussian_rasterization as dgr
from tools import runners
sc = torch.load({scene!r}, weights_only=False).to(torch.device("cuda:0"))
h = hashlib.sha256()
for _ in range(40):
    out = runners.run_public_api(sc, dgr)
    torch.cuda.synchronize()
    for k in ("color", "depth", "opacity", "radii", "n_touched"):
        h.update(out[k].detach().cpu().numpy().tobytes())
print("HASH " + h.hexdigest() + " " + json.dumps(float(out["dL_dmeans3D"].double().norm())))
"""


def test_two_processes_share_one_gpu(device, tmp_path):
    """Frontend and backend are two OS processes rendering on the same GPU (slam.py:143-150): two concurrent processes give
    the bits a single one gives."""
    sc = make_scene(20000, 320, 240, sh_degree=0, seed=102)
    path = str(tmp_path / "scene.pt")
    torch.save(sc, path)
    code = _WORKER.format(root=ROOT, pkg=os.path.join(ROOT, "4dgs-slam_b200"), scene=path)
    procs = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for _ in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    hashes = []
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-2000:]
        hashes.append([l for l in o.splitlines() if l.startswith("HASH ")][0].split()[1])
    import diff_gaussian_rasterization as dgr
    h = hashlib.sha256()
    scd = sc.to(device)
    for _ in range(40):
        out = runners.run_public_api(scd, dgr)
        for k in ("color", "depth", "opacity", "radii", "n_touched"):
            h.update(out[k].detach().cpu().numpy().tobytes())
    assert hashes[0] == hashes[1] == h.hexdigest()


def test_cpp_host_path_equals_the_python_host_path(device):
    """The standard op dispatches to the C++ host extension (csrc/host/g4r_torch.cpp) when it is built; the Python Function is the
    specification.  Same kernels through the same C ABI: outputs and integers are bit-identical, gradients agree to the run-to-run
    noise of the backward's floating-point atomics (1e-5), for SH and precomputed inputs, a pose-only backward (nothing but
    theta / rho asks for a gradient) and an empty model."""
    import diff_gaussian_rasterization as dgr
    if dgr.host_backend() != "cpp":
        pytest.skip("_g4r_host.so not built")

    def both(fn):
        a = fn()
        saved, dgr._host = dgr._host, None
        try:
            assert dgr.host_backend() == "python"
            b = fn()
        finally:
            dgr._host = saved
        return a, b

    for kw in (dict(P=6000, W=160, H=128, sh_degree=2, seed=21), dict(P=3000, W=96, H=64, sh_degree=0, seed=22, colors_precomp=True, cov3D_precomp=True)):
        sc = make_scene(**kw).to(device)
        a, b = both(lambda: runners.run_public_api(sc, dgr))
        for k, v in a.items():
            if not isinstance(v, torch.Tensor):
                assert v is None and b[k] is None, k
            elif k.startswith("dL_"):
                assert v.shape == b[k].shape and float((v - b[k]).norm() / b[k].norm()) < 1e-5, k
            else:
                assert torch.equal(v, b[k]), k

    sc = make_scene(4000, 128, 96, sh_degree=1, seed=23).to(device)

    def pose_only():
        leaf, m2d, theta, rho = _leafs(sc, device, grad=False)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(runners.settings_for(sc, dgr))(
            means3D=leaf["means3D"], means2D=m2d.detach(), opacities=leaf["opacities"], shs=leaf["shs"], scales=leaf["scales"],
            rotations=leaf["rotations"], theta=theta, rho=rho)
        ((color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()).backward()
        return torch.cat([rho.grad.reshape(-1), theta.grad.reshape(-1)]), color.detach(), n_touched

    a, b = both(pose_only)
    assert float((a[0] - b[0]).norm() / b[0].norm()) < 1e-5 and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])

    def empty():
        e3 = torch.zeros(0, 3, device=device, requires_grad=True)
        out = dgr.GaussianRasterizer(runners.settings_for(sc, dgr))(
            means3D=e3, means2D=torch.zeros(0, 3, device=device), opacities=torch.zeros(0, 1, device=device), shs=torch.zeros(0, 4, 3, device=device),
            scales=torch.zeros(0, 3, device=device), rotations=torch.zeros(0, 4, device=device), theta=torch.zeros(3, device=device), rho=torch.zeros(3, device=device))
        out[0].sum().backward()
        return out[0].detach(), out[1], e3.grad

    a, b = both(empty)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
