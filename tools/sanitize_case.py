"""Small forward+backward cases for compute-sanitizer (memcheck / racecheck / initcheck): the standard path on three scene
types (SH 3; precomputed colour + covariance; oversized tiles -> global-memory radix sort), the raw-parameter path with the
in-kernel mask and dynamic offsets, the fused loss kernel, distCUDA2 (simple_knn drop-in) and the control-node warp."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools import runners
from tools.scenes import make_scene

dev = torch.device("cuda:0")
for kw in (dict(P=700, W=100, H=75, sh_degree=3, seed=1), dict(P=500, W=64, H=48, sh_degree=0, seed=2, colors_precomp=True, cov3D_precomp=True),
           dict(P=6000, W=48, H=32, sh_degree=0, seed=3, px_min=4.0, px_max=30.0)):
    out = runners.run_public_api(make_scene(**kw).to(dev), __import__("diff_gaussian_rasterization"))
    torch.cuda.synchronize()
    print(kw, "visible", int((out["radii"] > 0).sum()))

import diff_gaussian_rasterization as dgr
from diff_gaussian_rasterization.losses import slam_loss
sc = make_scene(900, 96, 64, sh_degree=1, seed=4).to(dev)
raw = runners.raw_parameters(sc, seed=2)
g = torch.Generator().manual_seed(1)
dygs = (torch.rand(sc.P, generator=g) < 0.3).to(dev)
nd = int(dygs.sum())
leaf = {k: v.detach().clone().requires_grad_(True) for k, v in raw.items()}
off = [t.to(dev).requires_grad_(True) for t in (0.01 * torch.randn(nd, 3, generator=g), 0.001 * torch.rand(nd, 3, generator=g), 0.01 * torch.randn(nd, 4, generator=g))]
m2d = torch.zeros_like(leaf["xyz"], requires_grad=True)
theta, rho = torch.zeros(3, device=dev, requires_grad=True), torch.zeros(3, device=dev, requires_grad=True)
color, radii, depth, opacity, n_touched = dgr.FusedGaussianRasterizer(runners.settings_for(sc, dgr))(
    xyz=leaf["xyz"], means2D=m2d, features_dc=leaf["dc"], features_rest=leaf["rest"], opacity_raw=leaf["opacity"], scaling_raw=leaf["scaling"],
    rotation_raw=leaf["rotation"], theta=theta, rho=rho, mask=torch.rand(sc.P, generator=g).to(dev) > 0.2, dx=off[0], ds=off[1], dr=off[2],
    dyn_slot=dgr.dynamic_slots(dygs))
ea, eb = torch.tensor([0.05], device=dev, requires_grad=True), torch.tensor([0.01], device=dev, requires_grad=True)
loss = slam_loss("tracking", color, depth, torch.rand(3, sc.H, sc.W, generator=g).to(dev), (3 * torch.rand(1, sc.H, sc.W, generator=g)).to(dev),
                 opacity=opacity, exposure_a=ea, exposure_b=eb, motion_mask=(torch.rand(sc.H, sc.W, generator=g) > 0.3).to(dev),
                 grad_mask=(torch.rand(1, sc.H, sc.W, generator=g) > 0.1).to(dev))
loss.backward()
torch.cuda.synchronize()
print("fused path loss", float(loss.detach()))

from simple_knn._C import distCUDA2
from tools import knn_cases
for name, pts in knn_cases.random_cases(seed=5, n=3000):
    d = distCUDA2(torch.from_numpy(pts).to(dev))
torch.cuda.synchronize()
print("distCUDA2 cases done", float(d[0]))

from diff_gaussian_rasterization.deform import control_node_warp
for N, M, K, local in ((3000, 512, 3, True), (1500, 2300, 5, False)):
    x, nodes = torch.randn(N, 3, generator=g).to(dev), torch.randn(M, 5, generator=g).to(dev)
    lr, wl = (torch.randn(M, generator=g) * 0.3 - 0.5).to(dev).requires_grad_(), torch.randn(M, 1, generator=g).to(dev).requires_grad_()
    attrs = {k: (torch.randn(M, c, generator=g) * 0.3).to(dev).requires_grad_() for k, c in (("d_xyz", 3), ("d_rotation", 4), ("d_scaling", 3), ("local_rotation", 4))}
    o = control_node_warp(x, nodes, lr, wl, attrs, (torch.rand(N, 1, generator=g) > 0.3).float().to(dev), K=K, local_frame=local)
    (o["d_xyz"].sum() + o["d_rotation"].sum() + o["d_scaling"].sum()).backward()
torch.cuda.synchronize()
print("control_node_warp cases done", float(lr.grad.abs().sum()))
print("sanitize cases done")
