"""What folding the activation prelude into the kernels buys (run under gpurun, 1 GPU).

Times one mapping-style iteration on GaussianModel-style RAW parameters, fwd + loss + bwd:
  * prelude + rasterizer : torch.sigmoid / exp / normalize / cat (gaussian_model.py:100-128) feeding GaussianRasterizer
  * fused                : FusedGaussianRasterizer on the raw parameters
each eager and captured in a CUDA graph, and -- when baseline/_ref is present -- the same prelude feeding the reference build."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import refload, runners  # noqa: E402
from tools.perf_matrix import timeit  # noqa: E402
from tools.scenes import config_scene  # noqa: E402


def make_step(sc, dgr, raw, fused):
    rs = runners.settings_for(sc, dgr)
    dev = sc.means3D.device
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in raw.items()}

    def step():
        m2d = torch.zeros_like(leaf["xyz"], requires_grad=True)
        theta = torch.zeros(3, device=dev, requires_grad=True)
        rho = torch.zeros(3, device=dev, requires_grad=True)
        if fused:
            out = dgr.FusedGaussianRasterizer(rs)(xyz=leaf["xyz"], means2D=m2d, features_dc=leaf["dc"], features_rest=leaf["rest"],
                                                  opacity_raw=leaf["opacity"], scaling_raw=leaf["scaling"],
                                                  rotation_raw=leaf["rotation"], theta=theta, rho=rho)
        else:
            out = dgr.GaussianRasterizer(rs)(means3D=leaf["xyz"], means2D=m2d, opacities=torch.sigmoid(leaf["opacity"]),
                                             shs=torch.cat((leaf["dc"], leaf["rest"]), dim=1), scales=torch.exp(leaf["scaling"]),
                                             rotations=torch.nn.functional.normalize(leaf["rotation"]), theta=theta, rho=rho)
        loss = (out[0] * sc.grad_color).sum() + (out[2] * sc.grad_depth).sum()
        return torch.autograd.grad(loss, list(leaf.values()) + [m2d, theta, rho], allow_unused=True)
    return step


def graph_time(step, dgr, dev):
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
    ms = timeit(g.replay, warmup=3, iters=30)
    ok = not dgr.captured_overflow()
    dgr.reset_captured()
    del out
    return ms, ok


def main():
    dev = torch.device("cuda:0")
    import diff_gaussian_rasterization as dgr
    ref = refload.load() if refload.available() else None
    rows = {}
    for name in ("C3", "C2"):
        sc = config_scene(name).to(dev)
        raw = runners.raw_parameters(sc, seed=1)
        row = {"P": sc.P, "M": int(sc.shs.shape[1])}
        for tag, fused in (("prelude_plus_rasterizer", False), ("fused", True)):
            step = make_step(sc, dgr, raw, fused)
            row[tag + "_eager_ms"] = timeit(step, warmup=5, iters=30)
            try:
                row[tag + "_graph_ms"], row[tag + "_graph_ok"] = graph_time(step, dgr, dev)
            except Exception as exc:      # noqa: BLE001
                row[tag + "_graph_ms"] = f"failed: {exc!r}"[:200]
        if ref is not None:
            row["reference_prelude_plus_rasterizer_eager_ms"] = timeit(make_step(sc, ref, raw, False), warmup=5, iters=30)
        rows[name] = row
        print(name, json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "prelude_bench.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
