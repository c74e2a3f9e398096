"""fwd+bwd time of this repo vs the reference build over a matrix of scene types (run under gpurun, 1 GPU)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import refload, runners  # noqa: E402
from tools.scenes import Scene, config_scene, make_scene  # noqa: E402


def clustered(P, W, H, seed):
    """Depths concentrated on three thin 'surfaces' (what a SLAM map looks like): stresses the per-tile sort."""
    sc = make_scene(P, W, H, sh_degree=0, seed=seed, posed=False)
    g = torch.Generator().manual_seed(seed + 1)
    z_old = sc.means3D[:, 2].clone()
    layer = torch.randint(0, 3, (P,), generator=g)
    z_new = torch.tensor([1.0, 2.5, 4.0])[layer] + 0.01 * torch.randn(P, generator=g)
    z_new = torch.where(z_old > 0.2, z_new, z_old)
    scale = (z_new / z_old.clamp_min(1e-3)).unsqueeze(1)
    sc.means3D = torch.cat([sc.means3D[:, :2] * scale, z_new.unsqueeze(1)], 1).contiguous()
    sc.scales = (sc.scales * scale).contiguous()
    sc.name = "clustered"
    return sc


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def graph_ms(sc, dgr, iters=10):
    """fwd + loss + bwd of the public API captured in a CUDA graph; per-replay milliseconds."""
    dev = sc.means3D.device
    rs = runners.settings_for(sc, dgr)
    keys = [k for k in ("means3D", "opacities", "shs", "scales", "rotations") if getattr(sc, k) is not None]
    leaf = {k: getattr(sc, k).detach().clone().requires_grad_(True) for k in keys}

    def step():
        m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta = torch.zeros(3, device=dev, requires_grad=True)
        rho = torch.zeros(3, device=dev, requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
            means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf.get("shs"), colors_precomp=sc.colors_precomp,
            scales=leaf.get("scales"), rotations=leaf.get("rotations"), cov3D_precomp=sc.cov3D_precomp, theta=theta, rho=rho)
        loss = (color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()
        return torch.autograd.grad(loss, list(leaf.values()) + [m2d, theta, rho])

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    dgr.reset_captured()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    ms = timeit(g.replay, warmup=3, iters=iters)
    ok = not dgr.captured_overflow()
    dgr.reset_captured()
    del out
    return ms, ok


def main():
    dev = torch.device("cuda:0")
    import diff_gaussian_rasterization as dgr
    ref = refload.load() if refload.available() else None
    cases = {
        "C2 100k SH3 640x480": config_scene("C2"),
        "C3 500k SH0 640x480": config_scene("C3"),
        "C4 2M SH0 1280x960": config_scene("C4"),
        "100k big splats (3-25 px) 640x480": make_scene(100_000, 640, 480, sh_degree=0, seed=3, px_min=3.0, px_max=25.0),
        "200k clustered depths 640x480": clustered(200_000, 640, 480, 4),
        "30k tracking-size SH0 640x480": make_scene(30_000, 640, 480, sh_degree=0, seed=5, px_min=1.0, px_max=8.0),
    }
    rows = {}
    for name, sc_cpu in cases.items():
        sc = sc_cpu.to(dev)
        mine = runners.run_g4r(sc, want_grads=False)
        row = {"P": sc.P, "N": int(mine["num_rendered"]), "max_tile": int((mine["ranges"][:, 1] - mine["ranges"][:, 0]).max())}
        row["ours_ms"] = timeit(lambda: runners.run_public_api(sc, dgr))
        try:
            row["ours_cuda_graph_ms"], row["graph_capacity_ok"] = graph_ms(sc, dgr)
        except Exception as exc:
            row["ours_cuda_graph_ms"] = f"failed: {exc!r}"[:200]
        if ref is not None:
            r = refload.run_reference(sc, want_grads=False)
            row["ints_equal"] = bool(torch.equal(mine["point_list"], r["point_list"]) and torch.equal(mine["radii"], r["radii"])
                                     and torch.equal(mine["n_contrib"], r["n_contrib"]))
            row["image_bits_equal"] = bool(torch.equal(mine["color"], r["color"]) and torch.equal(mine["depth"], r["depth"]))
            row["ref_ms"] = timeit(lambda: runners.run_public_api(sc, ref))
            row["speedup"] = row["ref_ms"] / row["ours_ms"]
        rows[name] = row
        print(name, json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "perf_matrix.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
